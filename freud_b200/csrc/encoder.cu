// Host launchers for the tcgen05 encoder GEMM kernels (freud_topk_encode, freud_gemm_nt).
#include "gemm_sm100.cuh"
#include "host_common.h"
#include "../../include/freud_b200.h"

#include <cstdlib>

namespace freud {

template <int BN, int STAGES, int EPI, bool TF32, int SETS, int CL, int NBUF = 2>
static int launch_gemm(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, GemmParams p,
                       int passes, cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES, EPI, SETS, NBUF>;
  const int eb = TF32 ? 4 : 2;
  CUtensorMap mA0, mA1, mB0, mB1;
  if (make_tensor_map_2d(&mA0, a_hi, p.M, p.K, p.K, eb, kBM)) return 3;
  if (make_tensor_map_2d(&mB0, b_hi, p.N, p.K, p.K, eb, BN / CL)) return 3;
  if (passes > 1) {
    if (make_tensor_map_2d(&mA1, a_lo, p.M, p.K, p.K, eb, kBM)) return 3;
    if (make_tensor_map_2d(&mB1, b_lo, p.N, p.K, p.K, eb, BN / CL)) return 3;
  } else {
    mA1 = mA0;
    mB1 = mB0;
  }
  p.passes = passes;
  auto kern = sm100_gemm_kernel<BN, STAGES, EPI, TF32, SETS, CL, NBUF>;
  static bool attr_set = false;
  if (!attr_set) {
    FREUD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  int grid = (p.M + kBM - 1) / kBM;
  grid = (grid + CL - 1) / CL * CL;  // whole clusters; surplus CTAs run the protocol on out-of-range rows
  cudaLaunchConfig_t cfg{};
  int nsplit = 1;
  if (p.kb_per_split > 0) {
    const int kbe = TF32 ? 32 : 64;
    const int total_kb = (p.K + kbe - 1) / kbe;
    nsplit = (total_kb + p.kb_per_split - 1) / p.kb_per_split;
  }
  cfg.gridDim = dim3(grid, nsplit);
  cfg.blockDim = dim3(L::kThreads);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FREUD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, mA0, mA1, mB0, mB1, p));
  return 0;
}

// Encoder variant (epilogue sets / smem stages / multicast cluster); FREUD_ENC_VARIANT overrides for experiments.
static int encoder_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FREUD_ENC_VARIANT");
    v = e ? atoi(e) : 0;
  }
  return v;
}

}  // namespace freud

using namespace freud;

extern "C" int freud_topk_encode(const void* xc_hi, const void* xc_lo, const void* w_hi, const void* w_lo,
                                 const float* b_enc, float* top_vals, int32_t* top_idx, int64_t N, int64_t d,
                                 int64_t n, int precision, void* stream) {
  FREUD_REQUIRE(N > 0 && d > 0 && n >= 64, "freud_topk_encode needs N > 0 and n >= 64");
  FREUD_REQUIRE(d % 8 == 0, "activation size must be a multiple of 8");
  FREUD_REQUIRE(N < (1ll << 31) && n < (1ll << 31), "sizes exceed int32");
  GemmParams p{};
  p.M = static_cast<int>(N);
  p.N = static_cast<int>(n);
  p.K = static_cast<int>(d);
  p.bias = b_enc;
  p.relu = 1;
  p.top_vals = top_vals;
  p.top_idx = top_idx;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (precision == FREUD_BF16) {
    switch (encoder_variant()) {
      case 2: return launch_gemm<256, 3, EPI_TOPK, false, 2, 2>(xc_hi, nullptr, w_hi, nullptr, p, 1, s);
      // (a triple-buffered BN = 160 variant, launch_gemm<160, 4, EPI_TOPK, false, 2, 1, 3>, measured slower:
      //  2.02 vs 1.78 ms on C3 -- per-tile fixed costs outweigh the extra MMA/scan overlap)
      case 3: return launch_gemm<256, 3, EPI_TOPK, false, 1, 1>(xc_hi, nullptr, w_hi, nullptr, p, 1, s);
      case 6: p.out = top_vals; return launch_gemm<256, 3, EPI_NONE, false, 1, 1>(xc_hi, nullptr, w_hi, nullptr, p, 1, s);
      case 7: p.out = top_vals; return launch_gemm<256, 4, EPI_NONE, false, 1, 1>(xc_hi, nullptr, w_hi, nullptr, p, 1, s);
      case 8: p.out = top_vals; return launch_gemm<256, 3, EPI_NONE, false, 2, 1>(xc_hi, nullptr, w_hi, nullptr, p, 1, s);
      default: return launch_gemm<256, 3, EPI_TOPK, false, 2, 1>(xc_hi, nullptr, w_hi, nullptr, p, 1, s);
    }
  }
  if (precision == FREUD_FP32) return launch_gemm<256, 3, EPI_TOPK, true, 2, 1>(xc_hi, xc_lo, w_hi, w_lo, p, 3, s);
  FREUD_REQUIRE(false, "unknown precision");
}

extern "C" int freud_gemm_nt(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo,
                             const float* bias, float* out, int64_t M, int64_t N, int64_t K, int64_t ldo, int relu,
                             int precision, void* stream) {
  FREUD_REQUIRE(M > 0 && N > 0 && K > 0, "empty GEMM");
  FREUD_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "sizes exceed int32");
  FREUD_REQUIRE(ldo >= N, "ldo < N");
  GemmParams p{};
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(N);
  p.K = static_cast<int>(K);
  p.bias = bias;
  p.relu = relu;
  p.out = out;
  p.ldo = ldo;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (precision == FREUD_BF16) {
    FREUD_REQUIRE(K % 8 == 0, "K must be a multiple of 8 for bf16 operands");
    return launch_gemm<256, 4, EPI_STORE, false, 2, 1>(a_hi, nullptr, b_hi, nullptr, p, 1, s);
  }
  if (precision == FREUD_FP32) {
    FREUD_REQUIRE(K % 4 == 0, "K must be a multiple of 4 for fp32 operands");
    return launch_gemm<256, 4, EPI_STORE, true, 2, 1>(a_hi, a_lo, b_hi, b_lo, p, 3, s);
  }
  FREUD_REQUIRE(false, "unknown precision");
}

// Split-K variant for "weight-gradient shaped" products (few output rows, very long K): `splits` partial results are
// written to workspace [splits, M, N] and summed into out by freud_sum_splits.
extern "C" int freud_gemm_nt_splitk(const void* a_hi, const void* b_hi, float* workspace, int64_t M, int64_t N, int64_t K,
                                    int64_t splits, void* stream) {
  FREUD_REQUIRE(M > 0 && N > 0 && K > 0 && splits >= 1, "empty split-K GEMM");
  FREUD_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "sizes exceed int32");
  FREUD_REQUIRE(K % 8 == 0, "K must be a multiple of 8 for bf16 operands");
  GemmParams p{};
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(N);
  p.K = static_cast<int>(K);
  p.out = workspace;
  p.ldo = N;
  const int total_kb = static_cast<int>((K + 63) / 64);
  p.kb_per_split = static_cast<int>((total_kb + splits - 1) / splits);
  FREUD_REQUIRE((total_kb + p.kb_per_split - 1) / p.kb_per_split == splits,
                "splits must equal ceil(k_blocks / ceil(k_blocks / splits)) so that no partial is left unwritten");
  p.split_stride = M * N;
  return launch_gemm<256, 4, EPI_STORE, false, 2, 1>(a_hi, nullptr, b_hi, nullptr, p, 1, static_cast<cudaStream_t>(stream));
}
