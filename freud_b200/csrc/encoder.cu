// Host launchers for the tcgen05 encoder GEMM kernels (freud_topk_encode, freud_gemm_nt).
#include "gemm_sm100.cuh"
#include "topk_sm100.cuh"
#include "host_common.h"
#include "../../include/freud_b200.h"

#include <cstdlib>

namespace freud {

// Tail of the fused encoder: rows of a row block whose column tiles were scanned by several CTAs have one sorted
// top-32 list per piece (indices already global); one warp per row folds them together exactly like
// shard_merge_kernel does for feature shards (max with the reversed other list + a 5-stage bitonic merge).
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  const uint32_t lo = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v), src);
  const uint32_t hi = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), src);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
__global__ void __launch_bounds__(256) topk_merge_pieces_kernel(const float* __restrict__ part_vals,
                                                                const int32_t* __restrict__ part_idx,
                                                                int64_t part_stride, float* __restrict__ top_vals,
                                                                int32_t* __restrict__ top_idx, int M, int row0,
                                                                int pieces, int32_t* __restrict__ hist) {
  const int lane = threadIdx.x & 31;
  const int64_t row = row0 + static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  uint64_t cur = 0;
  for (int piece = 0; piece < pieces; ++piece) {
    const int64_t o = piece * part_stride + row * 32 + lane;
    const uint64_t key = (static_cast<uint64_t>(__float_as_uint(part_vals[o])) << 32) |
                         static_cast<uint32_t>(~static_cast<uint32_t>(part_idx[o]));
    if (piece == 0) {
      cur = key;
    } else {
      const uint64_t rev = shfl_u64(key, 31 - lane);
      cur = cur > rev ? cur : rev;
#pragma unroll
      for (int j = 16; j > 0; j >>= 1) {
        const uint64_t other = shfl_u64(cur, lane ^ j);
        const bool lower = (lane & j) == 0;
        const uint64_t mx = cur > other ? cur : other, mn = cur > other ? other : cur;
        cur = lower ? mx : mn;
      }
    }
  }
  top_vals[row * 32 + lane] = __uint_as_float(static_cast<uint32_t>(cur >> 32));
  top_idx[row * 32 + lane] = static_cast<int32_t>(~static_cast<uint32_t>(cur));
  if (hist != nullptr) atomicAdd(hist + ~static_cast<uint32_t>(cur), 1);
}

// Per-feature counts of an index matrix: the encoder variants that do not count while they emit.
__global__ void __launch_bounds__(256) hist_rows_kernel(const int32_t* __restrict__ top_idx, int64_t total,
                                                        int32_t* __restrict__ hist) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int32_t f = top_idx[i];
    if (f >= 0) atomicAdd(hist + f, 1);
  }
}

// A batch that is not a multiple of sm_count row blocks would leave SMs idle in its last wave, so the row blocks
// of that wave are cut into S column ranges scanned by different CTAs.  S minimises the wave count of the tail,
// charging five tiles per piece (measured on C3: pipeline fill, threshold warm-up and less L2 sharing of the weight
// tiles; the pieces of a row do share their selection threshold).  Returns S (1 = no split)
// and the number of row blocks in whole waves.
static int plan_tail_split(int num_mb, int num_nt, int* full_count) {
  const int sms = sm_count();
  const int full = num_mb / sms * sms, tail = num_mb - full;
  *full_count = full;
  if (tail == 0 || getenv("FREUD_NO_TAIL_SPLIT")) return 1;
  int best = 1;
  double best_cost = 1.0;
  for (int S = 2; S <= 16 && S * 2 <= num_nt; ++S) {
    const double waves = static_cast<double>((static_cast<int64_t>(tail) * S + sms - 1) / sms) / S;
    const double cost = waves * (1.0 + 5.0 * S / num_nt);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = S;
    }
  }
  if (const char* e = getenv("FREUD_TAIL_SPLIT")) {  // experiments: force the split factor
    if (atoi(e) >= 1 && atoi(e) * 2 <= num_nt) best = atoi(e);
  }
  return best;
}
// partial lists [S][rows][32] (values + indices) and the shared per-row thresholds
static int64_t tail_split_bytes(int num_mb, int S) {
  const int64_t rows = static_cast<int64_t>(num_mb) * kBM;
  return S > 1 ? S * rows * 32 * 8 + rows * 4 : 0;
}

template <int BN, int STAGES, int EPI, bool TF32, int SETS, int CL, int NBUF = 2, int CEV = 0, bool AMN = false,
          bool BMN = false>
static int launch_gemm(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, GemmParams p,
                       int passes, cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES, EPI, SETS, NBUF>;
  const int eb = TF32 ? 4 : 2;
  CUtensorMap mA0, mA1, mB0, mB1;
  // K-major operand: stored [M or N, K], boxes of {128 B of K, rows}; MN-major: stored [K, M or N], boxes {64, 64}
  auto map_a = [&](CUtensorMap* m, const void* base) {
    if (AMN) return make_tensor_map_2d(m, base, p.K, p.M, p.lda ? p.lda : p.M, eb, 64);
    return make_tensor_map_2d(m, base, p.M, p.K, p.lda ? p.lda : p.K, eb, kBM);
  };
  auto map_b = [&](CUtensorMap* m, const void* base) {
    if (BMN) return make_tensor_map_2d(m, base, p.K, p.N, p.ldb ? p.ldb : p.N, eb, 64);
    return make_tensor_map_2d(m, base, p.N, p.K, p.ldb ? p.ldb : p.K, eb, BN / CL);
  };
  if (map_a(&mA0, a_hi)) return 3;
  if (map_b(&mB0, b_hi)) return 3;
  if (passes > 1) {
    if (map_a(&mA1, a_lo)) return 3;
    if (map_b(&mB1, b_lo)) return 3;
  } else {
    mA1 = mA0;
    mB1 = mB0;
  }
  p.passes = passes;
  // output through TMA boxes when the output matrix qualifies (16-byte aligned rows; no split-K partial planes)
  CUtensorMap mO = mA0;
  p.tma_out = 0;
  if (L::kHasOut && p.kb_per_split == 0 && !getenv("FREUD_NO_TMA_OUT")) {
    if (EPI == EPI_STORE) {
      if (p.out && (p.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0)
        p.tma_out = make_tensor_map_2d(&mO, p.out, p.M, p.N, p.ldo, 4, 32) == 0;
    } else if (p.out16 && (p.ld16 & 7) == 0 && (reinterpret_cast<uintptr_t>(p.out16) & 15) == 0) {
      p.tma_out = make_tensor_map_2d(&mO, p.out16, p.M, p.ld16, p.ld16, 2, 32) == 0;
    }
  }
  // EPI_RESID / EPI_MASK: the matrix the epilogue reads row-per-thread arrives as TMA boxes too
  CUtensorMap mI = mA0;
  p.tma_in = 0;
  if (L::kHasIn && p.tma_out && !getenv("FREUD_NO_TMA_IN")) {
    if (EPI == EPI_RESID) {
      if (p.target && (p.ldt & 3) == 0 && (reinterpret_cast<uintptr_t>(p.target) & 15) == 0)
        p.tma_in = make_tensor_map_2d(&mI, p.target, p.M, p.N, p.ldt, 4, 32) == 0;
    } else if (p.mask_src && (reinterpret_cast<uintptr_t>(p.mask_src) & 15) == 0) {
      p.tma_in = make_tensor_map_2d(&mI, p.mask_src, p.M, p.ld16, p.ld16, 2, 32) == 0;
    }
  }
  auto kern = sm100_gemm_kernel<BN, STAGES, EPI, TF32, SETS, CL, NBUF, CEV, AMN, BMN>;
  if (const char* e = getenv("FREUD_ENC_FLAGS")) p.flags = atoi(e);
  static bool attr_set = false;
  if (!attr_set) {
    FREUD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  int grid = (p.M + kBM - 1) / kBM;
  grid = (grid + CL - 1) / CL * CL;  // whole clusters; surplus CTAs run the protocol on out-of-range rows
  if ((EPI == EPI_STORE || EPI == EPI_MASK || EPI == EPI_RELU16 || EPI == EPI_RESID) && CL == 1 &&
      p.kb_per_split == 0 && grid > sm_count() &&
      !getenv("FREUD_NO_PERSISTENT")) {
    p.persistent = 1;  // one CTA per SM, each a contiguous run of row blocks
    grid = sm_count();
  }
  // Fused top-k encoder: the row blocks of a last, partial wave are cut into column ranges (plan_tail_split) when
  // the caller supplied the workspace for the partial lists.
  const int num_mb = (p.M + kBM - 1) / kBM;
  const int num_nt = (p.N + BN - 1) / BN;
  if (EPI == EPI_TOPK && CL == 1 && p.part_vals != nullptr) {
    int full = 0;
    const int S = plan_tail_split(num_mb, num_nt, &full);
    const int64_t stride = static_cast<int64_t>(num_mb) * kBM * 32;
    if (S > 1 && tail_split_bytes(num_mb, S) <= p.part_stride) {  // part_stride carries the workspace size in
      p.full_count = full;
      p.tail_split = S;
      p.part_stride = stride;
      p.part_idx = reinterpret_cast<int32_t*>(p.part_vals + S * stride);
      p.part_thr = reinterpret_cast<float*>(p.part_idx + S * stride);
      FREUD_CHECK_CUDA(cudaMemsetAsync(p.part_thr, 0, static_cast<size_t>(num_mb) * kBM * sizeof(float), stream));
      grid = full + (num_mb - full) * S;
    } else {
      p.part_vals = nullptr;
    }
  }
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
  FREUD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, mA0, mA1, mB0, mB1, mO, mI, p));
  if (p.tail_split > 1) {
    const int row0 = p.full_count * kBM;
    topk_merge_pieces_kernel<<<(p.M - row0 + 7) / 8, 256, 0, stream>>>(p.part_vals, p.part_idx, p.part_stride, p.top_vals,
                                                                       p.top_idx, p.M, row0, p.tail_split, p.hist);
    FREUD_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

// Fused top-k encoder with specialised scanner / compactor epilogue warps (topk_sm100.cuh): one CTA per row block,
// the row blocks of a partial last wave cut into column pieces exactly as in launch_gemm.
// Cycle counters of the diagnostic build (FREUD_ENC_STATS=1 selects it): [0] scanner warp lifetime, [1] scanner waiting
// for an accumulator, [2] scanner waiting for a free candidate buffer, [3] hand-overs, [4] compactor warp lifetime,
// [5] compactor waiting for a hand-over, [6] compactions, [7] compactor warps, [8] tiles scanned (summed over warps).
static unsigned long long* g_enc_stats = nullptr;
static bool enc_stats_enabled() {
  static const bool on = [] { const char* e = getenv("FREUD_ENC_STATS"); return e && atoi(e) != 0; }();
  return on;
}

template <int BN, int STAGES, bool TF32, int CEV = 0, bool STATS = false>
static int launch_topk(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, GemmParams p, int passes,
                       cudaStream_t stream) {
  using L = TopkSmem<BN, STAGES>;
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
  p.stats = nullptr;
  if (STATS) {
    if (g_enc_stats == nullptr) {
      FREUD_CHECK_CUDA(cudaMalloc(&g_enc_stats, 16 * sizeof(unsigned long long)));
      FREUD_CHECK_CUDA(cudaMemset(g_enc_stats, 0, 16 * sizeof(unsigned long long)));
    }
    p.stats = g_enc_stats;
  }
  auto kern = sm100_topk_kernel<BN, STAGES, TF32, CEV, STATS>;
  static bool attr_set = false;
  if (!attr_set) {
    FREUD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  const int num_mb = (p.M + kBM - 1) / kBM;
  const int num_nt = (p.N + BN - 1) / BN;
  int grid = num_mb;
  if (p.part_vals != nullptr) {
    int full = 0;
    const int S = plan_tail_split(num_mb, num_nt, &full);
    const int64_t stride = static_cast<int64_t>(num_mb) * kBM * 32;
    if (S > 1 && tail_split_bytes(num_mb, S) <= p.part_stride) {  // part_stride carries the workspace size in
      p.full_count = full;
      p.tail_split = S;
      p.part_stride = stride;
      p.part_idx = reinterpret_cast<int32_t*>(p.part_vals + S * stride);
      p.part_thr = reinterpret_cast<float*>(p.part_idx + S * stride);
      FREUD_CHECK_CUDA(cudaMemsetAsync(p.part_thr, 0, static_cast<size_t>(num_mb) * kBM * sizeof(float), stream));
      grid = full + (num_mb - full) * S;
    } else {
      p.part_vals = nullptr;
    }
  }
  kern<<<grid, L::kThreads, L::kTotal, stream>>>(mA0, mA1, mB0, mB1, p);
  FREUD_CHECK_CUDA(cudaGetLastError());
  if (p.tail_split > 1) {
    const int row0 = p.full_count * kBM;
    topk_merge_pieces_kernel<<<(p.M - row0 + 7) / 8, 256, 0, stream>>>(p.part_vals, p.part_idx, p.part_stride, p.top_vals,
                                                                       p.top_idx, p.M, row0, p.tail_split, p.hist);
    FREUD_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

// Encoder variant (FREUD_ENC_VARIANT, experiments): 0 = specialised scanner / compactor warps, compare-exchange on the
// keys as doubles (default; C2 0.49 ms); 4 / 5 = same with the borrow-mask / integer-predicate compare-exchange (0.58 /
// 0.63 ms); 1 = generic kernel, two column-split epilogue sets compacting inline; 3 = generic kernel, one inline set
// (0.68 ms); 6-8 = mainloop probes.
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

extern "C" int freud_topk_encode_workspace(int64_t N, int64_t n, int64_t* bytes) {
  FREUD_REQUIRE(N > 0 && n >= 64 && bytes != nullptr, "freud_topk_encode_workspace: bad arguments");
  FREUD_REQUIRE(N < (1ll << 31) && n < (1ll << 31), "sizes exceed int32");
  const int num_mb = static_cast<int>((N + kBM - 1) / kBM), num_nt = static_cast<int>((n + 255) / 256);
  int full = 0;
  *bytes = tail_split_bytes(num_mb, plan_tail_split(num_mb, num_nt, &full));
  return 0;
}

extern "C" int freud_topk_encode(const void* xc_hi, const void* xc_lo, const void* w_hi, const void* w_lo,
                                 const float* b_enc, float* top_vals, int32_t* top_idx, int64_t N, int64_t d,
                                 int64_t n, int precision, void* workspace, int64_t workspace_bytes, int32_t* hist,
                                 void* stream) {
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
  p.part_vals = static_cast<float*>(workspace);  // launch_gemm turns these two into the tail-split plan
  p.part_stride = workspace ? workspace_bytes : 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // the specialised kernel counts while it emits; the experiment variants get a separate counting pass
  const int variant = encoder_variant();
  const bool counts_inline = variant == 0 || variant == 4 || variant == 5 || variant > 8 || precision == FREUD_FP32;
  p.hist = counts_inline && !(precision == FREUD_FP32 && variant == 1) ? hist : nullptr;
  if (hist != nullptr && p.hist == nullptr) {
    const int rc = freud_topk_encode(xc_hi, xc_lo, w_hi, w_lo, b_enc, top_vals, top_idx, N, d, n, precision, workspace,
                                     workspace_bytes, nullptr, stream);
    if (rc) return rc;
    hist_rows_kernel<<<sm_count() * 4, 256, 0, s>>>(top_idx, N * 32, hist);
    FREUD_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  if (precision == FREUD_BF16) {
    switch (encoder_variant()) {
      case 1: return launch_gemm<256, 3, EPI_TOPK, false, 2, 1>(xc_hi, nullptr, w_hi, nullptr, p, 1, s);
      case 2: return launch_gemm<256, 3, EPI_TOPK, false, 2, 2>(xc_hi, nullptr, w_hi, nullptr, p, 1, s);
      case 3: return launch_gemm<256, 3, EPI_TOPK, false, 1, 1>(xc_hi, nullptr, w_hi, nullptr, p, 1, s);
      case 4: return launch_topk<256, 3, false, 0>(xc_hi, nullptr, w_hi, nullptr, p, 1, s);
      case 5: return launch_topk<256, 3, false, 1>(xc_hi, nullptr, w_hi, nullptr, p, 1, s);
      case 6: p.out = top_vals; return launch_gemm<256, 3, EPI_NONE, false, 1, 1>(xc_hi, nullptr, w_hi, nullptr, p, 1, s);
      case 7: p.out = top_vals; return launch_gemm<256, 4, EPI_NONE, false, 1, 1>(xc_hi, nullptr, w_hi, nullptr, p, 1, s);
      case 8: p.out = top_vals; return launch_gemm<256, 3, EPI_NONE, false, 2, 1>(xc_hi, nullptr, w_hi, nullptr, p, 1, s);
      default:
        if (enc_stats_enabled()) return launch_topk<256, 3, false, 2, true>(xc_hi, nullptr, w_hi, nullptr, p, 1, s);
        return launch_topk<256, 3, false, 2>(xc_hi, nullptr, w_hi, nullptr, p, 1, s);
    }
  }
  if (precision == FREUD_FP32) {
    if (encoder_variant() == 1) return launch_gemm<256, 3, EPI_TOPK, true, 2, 1>(xc_hi, xc_lo, w_hi, w_lo, p, 3, s);
    return launch_topk<256, 3, true, 2>(xc_hi, xc_lo, w_hi, w_lo, p, 3, s);
  }
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
    return launch_gemm<256, 3, EPI_STORE, false, 2, 1>(a_hi, nullptr, b_hi, nullptr, p, 1, s);
  }
  if (precision == FREUD_FP32) {
    FREUD_REQUIRE(K % 4 == 0, "K must be a multiple of 4 for fp32 operands");
    return launch_gemm<256, 3, EPI_STORE, true, 2, 1>(a_hi, a_lo, b_hi, b_lo, p, 3, s);
  }
  FREUD_REQUIRE(false, "unknown precision");
}

// Products whose operands are stored "the other way round" (MN-major operands, see sm100_gemm_kernel):
//   freud_gemm_tn_splitk : out[M, N] = A^T B with A stored [K, lda >= M] and B stored [K, ldb >= N], both bf16 and
//                          row-major -- the weight-gradient shape (K = tokens); `splits` fp32 partials go to
//                          workspace [splits, M, N] for freud_sum_splits.  No operand is transposed in memory.
//   freud_gemm_nn        : out[M, ldo] = act(A B + bias) with A stored [M, K] and B stored [K, ldb >= N], bf16.
extern "C" int freud_gemm_tn_splitk(const void* a, const void* b, float* workspace, int64_t M, int64_t N, int64_t K,
                                    int64_t lda, int64_t ldb, int64_t splits, void* stream) {
  FREUD_REQUIRE(M > 0 && N > 0 && K > 0 && splits >= 1, "empty split-K GEMM");
  FREUD_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "sizes exceed int32");
  FREUD_REQUIRE(lda >= M && ldb >= N && lda % 8 == 0 && ldb % 8 == 0, "row pitches must be multiples of 8 elements");
  GemmParams p{};
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(N);
  p.K = static_cast<int>(K);
  p.lda = lda;
  p.ldb = ldb;
  p.out = workspace;
  p.ldo = N;
  const int total_kb = static_cast<int>((K + 63) / 64);
  p.kb_per_split = static_cast<int>((total_kb + splits - 1) / splits);
  FREUD_REQUIRE((total_kb + p.kb_per_split - 1) / p.kb_per_split == splits,
                "splits must equal ceil(k_blocks / ceil(k_blocks / splits)) so that no partial is left unwritten");
  p.split_stride = M * N;
  return launch_gemm<256, 3, EPI_STORE, false, 2, 1, 2, 0, true, true>(a, nullptr, b, nullptr, p, 1,
                                                                       static_cast<cudaStream_t>(stream));
}

extern "C" int freud_gemm_nn(const void* a, const void* b, const float* bias, float* out, int64_t M, int64_t N,
                             int64_t K, int64_t lda, int64_t ldb, int64_t ldo, int relu, void* stream) {
  FREUD_REQUIRE(M > 0 && N > 0 && K > 0, "empty GEMM");
  FREUD_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "sizes exceed int32");
  FREUD_REQUIRE(lda >= K && ldb >= N && lda % 8 == 0 && ldb % 8 == 0 && ldo >= N, "bad row pitches");
  GemmParams p{};
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(N);
  p.K = static_cast<int>(K);
  p.lda = lda;
  p.ldb = ldb;
  p.bias = bias;
  p.relu = relu;
  p.out = out;
  p.ldo = ldo;
  return launch_gemm<256, 3, EPI_STORE, false, 2, 1, 2, 0, false, true>(a, nullptr, b, nullptr, p, 1,
                                                                        static_cast<cudaStream_t>(stream));
}

// dpre[M, ld16] (bf16) = (act > 0) ? A B^T : 0 -- activation gradient of the dense AuxK branch with the ReLU / top-k
// mask applied in the GEMM epilogue (autograd of relu + topk on the dead subset, topkautoencoder.py:118-123).
extern "C" int freud_gemm_nt_mask(const void* a, const void* b, const void* act_bf16, void* out_bf16,
                                  const float* affine, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ld16,
                                  void* stream) {
  FREUD_REQUIRE(M > 0 && N > 0 && K > 0 && K % 8 == 0, "gemm_nt_mask: bad sizes");
  FREUD_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "sizes exceed int32");
  FREUD_REQUIRE(ld16 >= N && ld16 % 8 == 0, "ld16 must be a multiple of 8 and >= N");
  GemmParams p{};
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(N);
  p.K = static_cast<int>(K);
  p.mask_src = static_cast<const __nv_bfloat16*>(act_bf16);
  p.out16 = static_cast<__nv_bfloat16*>(out_bf16);
  p.ld16 = ld16;
  p.affine = affine;
  p.lda = lda;
  return launch_gemm<256, 3, EPI_MASK, false, 2, 1>(a, nullptr, b, nullptr, p, 1, static_cast<cudaStream_t>(stream));
}

// L1 SAE forward on bf16 operands (l1autoencoder.py:69-95):
//   freud_l1_encode_fused: c = relu(x W + b) -> bf16 [M, ld16] (the decode GEMM's operand), sums[0] += sum(c); optional
//                          fp32 copy `latent` [M, N].   a = x bf16 [M, K = d], b = W^T bf16 [N = n, K = d].
//   freud_l1_decode_fused: x_hat = c W^T; against target x: sums[0] += masked sse, sums[1] += count(x != -1),
//                          sums[2] += sse; e * [x != -1] -> bf16 [M, ld16] (unscaled gradient seed); optional fp32 copy
//                          `x_hat` [M, N].   a = c bf16 [M, lda] (K = n valid columns), b = W bf16 [N = d, K = n].
extern "C" int freud_l1_encode_fused(const void* x_bf16, const void* wt_bf16, const float* bias, void* c_bf16,
                                     float* latent, double* sums, int64_t M, int64_t N, int64_t K, int64_t ld16,
                                     void* stream) {
  FREUD_REQUIRE(M > 0 && N > 0 && K > 0 && K % 8 == 0, "l1_encode_fused: bad sizes");
  FREUD_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "sizes exceed int32");
  FREUD_REQUIRE(ld16 >= N && ld16 % 8 == 0, "ld16 must be a multiple of 8 and >= N");
  GemmParams p{};
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(N);
  p.K = static_cast<int>(K);
  p.bias = bias;
  p.out16 = static_cast<__nv_bfloat16*>(c_bf16);
  p.ld16 = ld16;
  p.out = latent;
  p.ldo = N;
  p.sums = sums;
  return launch_gemm<256, 3, EPI_RELU16, false, 2, 1>(x_bf16, nullptr, wt_bf16, nullptr, p, 1,
                                                      static_cast<cudaStream_t>(stream));
}

extern "C" int freud_l1_decode_fused(const void* c_bf16, const void* w_bf16, const float* target, void* resid_bf16,
                                     float* x_hat, double* sums, int64_t M, int64_t N, int64_t K, int64_t lda,
                                     int64_t ldb, int64_t ld16, void* stream) {
  FREUD_REQUIRE(M > 0 && N > 0 && K > 0, "l1_decode_fused: bad sizes");
  FREUD_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "sizes exceed int32");
  FREUD_REQUIRE(lda >= K && lda % 8 == 0 && ldb >= K && ldb % 8 == 0, "operand pitches must be multiples of 8");
  FREUD_REQUIRE(ld16 >= N && ld16 % 8 == 0, "ld16 must be a multiple of 8 and >= N");
  GemmParams p{};
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(N);
  p.K = static_cast<int>(K);
  p.lda = lda;
  p.ldb = ldb;
  p.target = target;
  p.ldt = N;
  p.out16 = static_cast<__nv_bfloat16*>(resid_bf16);
  p.ld16 = ld16;
  p.out = x_hat;
  p.ldo = N;
  p.sums = sums;
  return launch_gemm<256, 2, EPI_RESID, false, 2, 1>(c_bf16, nullptr, w_bf16, nullptr, p, 1,
                                                     static_cast<cudaStream_t>(stream));
}

/* Diagnostic: cycle counters of the instrumented top-k encoder build (selected by FREUD_ENC_STATS=1), summed over all
 * launches since the last reset.  out[0..8], see g_enc_stats.  Synchronises the device. */
extern "C" int freud_topk_encode_stats(unsigned long long* out, int reset) {
  FREUD_REQUIRE(out != nullptr, "stats: out is NULL");
  for (int i = 0; i < 9; ++i) out[i] = 0;
  if (g_enc_stats == nullptr) return 0;
  FREUD_CHECK_CUDA(cudaDeviceSynchronize());
  FREUD_CHECK_CUDA(cudaMemcpy(out, g_enc_stats, 9 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (reset) FREUD_CHECK_CUDA(cudaMemset(g_enc_stats, 0, 16 * sizeof(unsigned long long)));
  return 0;
}
