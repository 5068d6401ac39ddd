// Host launchers for the tcgen05 encoder GEMM kernels (freud_topk_encode, freud_gemm_nt).
#include "gemm_sm100.cuh"
#include "host_common.h"
#include "../../include/freud_b200.h"

namespace freud {

template <int BN, int STAGES, int EPI, bool TF32>
static int launch_gemm(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, GemmParams p,
                       int passes, cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES>;
  const int eb = TF32 ? 4 : 2;
  CUtensorMap mA0, mA1, mB0, mB1;
  if (make_tensor_map_2d(&mA0, a_hi, p.M, p.K, p.K, eb, kBM)) return 3;
  if (make_tensor_map_2d(&mB0, b_hi, p.N, p.K, p.K, eb, BN)) return 3;
  if (passes > 1) {
    if (make_tensor_map_2d(&mA1, a_lo, p.M, p.K, p.K, eb, kBM)) return 3;
    if (make_tensor_map_2d(&mB1, b_lo, p.N, p.K, p.K, eb, BN)) return 3;
  } else {
    mA1 = mA0;
    mB1 = mB0;
  }
  p.passes = passes;
  auto kern = sm100_gemm_kernel<BN, STAGES, EPI, TF32>;
  static bool attr_set = false;
  if (!attr_set) {
    FREUD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  const int grid = (p.M + kBM - 1) / kBM;
  kern<<<grid, kGemmThreads, L::kTotal, stream>>>(mA0, mA1, mB0, mB1, p);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
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
  if (precision == FREUD_BF16) return launch_gemm<256, 3, EPI_TOPK, false>(xc_hi, nullptr, w_hi, nullptr, p, 1, s);
  if (precision == FREUD_FP32) return launch_gemm<256, 3, EPI_TOPK, true>(xc_hi, xc_lo, w_hi, w_lo, p, 3, s);
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
    return launch_gemm<256, 3, EPI_STORE, false>(a_hi, nullptr, b_hi, nullptr, p, 1, s);
  }
  if (precision == FREUD_FP32) {
    FREUD_REQUIRE(K % 4 == 0, "K must be a multiple of 4 for fp32 operands");
    return launch_gemm<256, 3, EPI_STORE, true>(a_hi, a_lo, b_hi, b_lo, p, 3, s);
  }
  FREUD_REQUIRE(false, "unknown precision");
}
