// Warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   D[M, N] = A[M, K] * B[N, K]^T + bias[N]   (both operands K-major, i.e. row-major with K contiguous)
//
//   * A is the (centred) activation block  x - b_dec      [tokens, d]
//   * B is the encoder weight              W_enc          [features, d]
//
// One CTA owns a 128-token row block and walks over ALL feature tiles (BN columns each), so each epilogue
// thread (TMEM lane == token row) sees every pre-activation of its token exactly once:
//   EPI_TOPK : ReLU + streaming exact top-32 selection per token, never materialising [M, N]
//              (reference: TopKAutoEncoder.pre_acts + select_topk, topkautoencoder.py:72-85)
//   EPI_STORE: (ReLU and) store fp32 [M, N]  (reference: pre_acts / L1 encode + decode GEMMs)
//   EPI_NONE : discard the tile (mainloop-ceiling probe used by scripts/enc_variants.py)
//   EPI_MASK : store bf16 (act[row, col] > 0 ? scale * acc + shift : 0) -- the ReLU / top-k mask of a dense backward
//              applied to the activation gradient without materialising the fp32 product (dense AuxK; L1 SAE dz)
//   EPI_RELU16: c = relu(acc + bias) stored as bf16 (the next GEMM's operand) with sum(c) accumulated -- the L1 SAE's
//              encode + L1 penalty (l1autoencoder.py:74-75,85); optional fp32 copy
//   EPI_RESID: e = acc - target; masked / unmasked squared error and count accumulated, e * [target != -1] stored as
//              bf16 -- the L1 SAE's decode + masked MSE (l1autoencoder.py:78,86,29-36) and its gradient seed; optional
//              fp32 copy of acc
//
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warp 3 spare, then SETS x 4
// epilogue warps (TMEM lane quarter == warp_idx % 4).  With SETS == 2 both sets drain EVERY tile, set s taking
// columns [s*BN/2, (s+1)*BN/2) of it, so every SM sub-partition hosts two epilogue warps that hide each other's
// latency and neither ever waits for "its" buffer while the MMA issuer fills it: the tile period is
// max(MMA, scan / 2).  (Round 1 gave set s the tiles s, s+2, ... with one accumulator buffer each; a set then idled
// for a whole MMA time per tile -- period (MMA + scan) / 2 -- and a single set, period max(MMA, scan), was faster.)
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM accumulator double buffer full/empty (MMA <-> epilogue).
//
// Bias: each set stages the bias row of its next tile in shared memory two tiles ahead; the scan reads it with
// broadcast loads and forms fl(accumulator + bias) per value -- the reference's own order of operations -- so the
// accumulator buffer goes back to the MMA issuer as soon as its last chunk has been read out of TMEM.  (Pre-storing
// the bias row into the drained buffer with tcgen05.st and letting every MMA accumulate was built first: it saves
// one FADD per value but holds the buffer for ~4 k cycles per tile; measured 3 % slower on C3.)  Out-of-range
// feature columns get -inf so they can never be selected.
//
// Selection (EPI_TOPK): each thread keeps its token's 32 best (value, ~index) keys SORTED IN REGISTERS and
// appends candidates above its threshold to a small shared-memory column (predicated store, no divergence).
// When any lane's column is nearly full the whole warp compacts in lock-step, every lane on its own row: a
// register-resident bitonic network sorts the new keys, merges them into the survivors and raises the
// threshold.  No cross-lane traffic, static register indexing only.
//
// Precision: kind::f16 with bf16 operands (1 pass), or kind::tf32 with the 3-pass split
//   A*B ~= A_hi*B_lo + A_lo*B_hi + A_hi*B_hi   (hi/lo are exact tf32 values prepared by prep kernels)
// which recovers ~fp32 accuracy (the reference's fp32 path) on the tensor cores.
#pragma once
#include "ptx.cuh"

namespace freud {

constexpr int kBM = 128;        // token rows per CTA (UMMA M)
constexpr int kBKBytes = 128;   // one swizzle-128B row per k-block
constexpr int kTopK = 32;       // fused selection width
constexpr int kNewSlots = 32;   // candidate slots per token between compactions
constexpr int kChunk = 16;      // accumulator columns per TMEM load
constexpr int kCheck = 8;       // columns between candidate-column occupancy checks
constexpr int kSlotStride = 32 * 8;  // bytes between slots of one lane: [slot][32 lanes] x 8 B.  A lane's 8-byte word
                                     // sits in banks {2*lane, 2*lane+1} whatever its slot (256 B == 0 mod 128 B), so the
                                     // predicated appends of lanes at DIFFERENT fill levels never collide

enum { EPI_TOPK = 0, EPI_STORE = 1, EPI_NONE = 2, EPI_MASK = 3, EPI_RELU16 = 4, EPI_RESID = 5 };

struct GemmParams {
  int M, N, K;          // K in elements
  int passes;           // 1 (bf16 / plain tf32) or 3 (split tf32)
  const float* bias;    // [N] or nullptr
  int relu;
  // EPI_TOPK
  float* top_vals;      // [M, 32]
  int32_t* top_idx;     // [M, 32]
  // EPI_STORE
  float* out;           // [M, ldo]
  int64_t ldo;
  // split-K (EPI_STORE): blockIdx.y = split, each handling kb_per_split k-blocks and writing its own partial
  // [M, ldo] at out + split * split_stride (bias only in split 0); 0 = no split
  int kb_per_split;
  int64_t split_stride;
  // Tail split (EPI_TOPK): row blocks [0, full_count) -- whole waves of one CTA per SM -- are scanned by one CTA
  // each; every row block of the last, partial wave is cut into tail_split column ranges scanned by different CTAs
  // (CTA full_count + s * tail + i takes range s of tail row block i, so CTAs running together walk the same weight
  // tiles), each writing a partial top-32 list at part_* + s * part_stride (+ row * 32); topk_merge_pieces_kernel
  // folds them.  tail_split <= 1: one CTA per row block throughout.
  int full_count;
  int tail_split;
  float* part_vals;
  int32_t* part_idx;
  int64_t part_stride;
  // per-row selection threshold shared by the pieces of a row block (zeroed by the launcher): any piece's 32nd
  // largest value so far bounds the row's final 32nd largest from below, so later / concurrent pieces start
  // filtering at once instead of rebuilding a threshold from zero
  float* part_thr;
  // Persistent row-block runs (EPI_STORE): CTA b owns row blocks [b*R/grid, (b+1)*R/grid) and walks them back to
  // back, so the TMA / MMA of the next row block overlaps the store epilogue of the previous one (a one-tile-wide
  // product such as the L1 SAE's n = 200 otherwise pays the pipeline fill and drain once per 128 rows).
  int persistent;
  // EPI_MASK
  const __nv_bfloat16* mask_src;  // [M, ld16] activations whose positivity gates the output
  __nv_bfloat16* out16;           // [M, ld16]
  int64_t ld16;                   // multiple of 8
  const float* affine;            // EPI_MASK: device (scale, shift) or nullptr for (1, 0)
  const float* target;            // EPI_RESID: fp32 [M, ldt]
  int64_t ldt;
  double* sums;                   // EPI_RELU16: [sum c]; EPI_RESID: [masked sse, count, sse]
  // Output through TMA (set by the launcher when the output pitch allows it and there is no split-K): the epilogue
  // warps stage 32-row x 128-byte boxes in swizzled shared memory and one lane issues a bulk tensor store per box --
  // whole 128-byte lines instead of 32 half-written sectors per store instruction (fp32 out for EPI_STORE, the bf16
  // out16 matrix for EPI_MASK / EPI_RELU16 / EPI_RESID)
  int tma_out;
  int tma_in;   // EPI_RESID target / EPI_MASK mask_src arrive as TMA boxes (needs tma_out: the smem budget assumes both)
  int64_t lda, ldb;  // host side only: row pitch (elements) of the A / B matrix as stored; 0 = dense
  // fused top-k encoder: per-feature counts of the emitted indices (the histogram the CSC index of the backward starts
  // from), accumulated with one atomic per emitted entry as rows are written; nullptr = not wanted
  int32_t* hist;
  unsigned long long* stats;  // diagnostic build of the top-k encoder (freud_topk_encode_stats), else nullptr
  int flags;  // experiments (FREUD_ENC_FLAGS): bit 0 = do not share 16th-largest values between the epilogue sets;
              // bits 1-4 = bare spin (no sleep between polls) in the producer / MMA-empty / MMA-full / epilogue waits
};

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int BN, int STAGES, int EPI, int SETS, int NBUF = 2>
struct GemmSmem {
  static constexpr int kEpiWarps = 4 * SETS;
  static constexpr int kThreads = 128 + kEpiWarps * 32;
  static constexpr int kABytes = kBM * kBKBytes;
  static constexpr int kBBytes = BN * kBKBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kRing = STAGES * kStageBytes;
  static constexpr bool kHasOut = EPI == EPI_STORE || EPI == EPI_MASK || EPI == EPI_RELU16 || EPI == EPI_RESID;
  static constexpr int kOutBox = 32 * 128;                              // 32 rows x 128 bytes
  // EPI_RESID also READS a matrix row-per-thread (the fp32 target): those boxes come in by TMA as well (two per
  // epilogue warp, fetched one box ahead), the output then gets by with one box per warp and the ring with two stages
  // (its products have K <= a few hundred).  EPI_MASK keeps three ring stages and reads its bf16 gate rows through
  // registers one chunk ahead behind the L2 prefetch warp: with K = d = 768 the third stage is worth more than the
  // input boxes (measured at the dense-AuxK shape: 0.23 ms vs 0.30 ms).
  static constexpr bool kHasIn = EPI == EPI_RESID;
  static constexpr int kOutBufs = kHasIn ? 1 : 2;
  static constexpr int kStageOut = kHasOut ? kEpiWarps * kOutBufs * kOutBox : 0;  // boxes are 1024-byte aligned
  static constexpr int kStageIn = kHasIn ? kEpiWarps * 2 * kOutBox : 0;
  static constexpr int kBufPerWarp = kNewSlots * kSlotStride;
  static constexpr int kBuf = EPI == EPI_TOPK ? kEpiWarps * kBufPerWarp : 0;
  static constexpr int kThr = 2 * kBM * 4 + 16;  // per-set published 16th-largest values
  static constexpr int kCols = BN / SETS;                   // columns of a tile one epilogue warp scans
  static constexpr int kBiasS = kEpiWarps * 2 * kCols * 4;  // staged bias rows: [epilogue warp][use parity][kCols]
  static constexpr int kBars = (2 * STAGES + 2 * NBUF) * 8 + 16 + (kHasIn ? kEpiWarps * 2 * 8 : 0);
  static constexpr int kTotal = kRing + kStageOut + kStageIn + kBuf + kThr + kBiasS + kBars;
};

// ---------------------------------------------------------------------------------------------------------------
// Register-resident sorting networks on 64-bit keys: high word = fp32 bits of a strictly positive value,
// low word = ~index, so unsigned order == (value descending, index ascending).  0 == empty.
// Compare-exchange without predicates: hi <- max(hi, lo), lo <- min(hi, lo).  A 64-bit compare-and-select costs the
// compiler 4 ISETP + 4 SEL and, worse, two of the thread's seven predicate registers per exchange, which caps the
// sixteen independent exchanges of a network stage at ~3 in flight (measured: 12-14 k cycles per compaction, IPC
// 0.2).  Keys are < 2^63 (value bits < 2^31), so the sign of the 64-bit difference is the comparison: one
// subtract-with-borrow pair, one arithmetic shift to a mask, four LOP3 selects -- 7 instructions, carry flag only.
template <int CEV = 0>
__device__ __forceinline__ void cmp_exchange(uint64_t& hi, uint64_t& lo) {
  if constexpr (CEV == 2) {  // experiment: the keys as positive doubles (same order as their bit patterns: value
                             // bits < 0x7f800000 never form a NaN / Inf exponent): one DSETP and four selects
    const bool sw = __longlong_as_double(static_cast<long long>(hi)) < __longlong_as_double(static_cast<long long>(lo));
    const uint64_t mx = sw ? lo : hi, mn = sw ? hi : lo;
    hi = mx;
    lo = mn;
    return;
  }
  if constexpr (CEV == 1) {  // experiment: one 64-bit compare (ISETP + ISETP.EX) and four predicated selects
    const bool sw = hi < lo;
    const uint64_t mx = sw ? lo : hi, mn = sw ? hi : lo;
    hi = mx;
    lo = mn;
    return;
  }
  const uint32_t al = static_cast<uint32_t>(hi), ah = static_cast<uint32_t>(hi >> 32);
  const uint32_t bl = static_cast<uint32_t>(lo), bh = static_cast<uint32_t>(lo >> 32);
  uint32_t m;  // all ones iff hi < lo
  asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %1, %2;\n\tsubc.u32 %0, %3, %4;\n\tshr.s32 %0, %0, 31;\n\t}"
      : "=r"(m)
      : "r"(al), "r"(bl), "r"(ah), "r"(bh));
  const uint32_t xl = (al & ~m) | (bl & m), xh = (ah & ~m) | (bh & m);  // max
  const uint32_t nl = (bl & ~m) | (al & m), nh = (bh & ~m) | (ah & m);  // min
  hi = (static_cast<uint64_t>(xh) << 32) | xl;
  lo = (static_cast<uint64_t>(nh) << 32) | nl;
}

template <int N, int CEV = 0>
__device__ __forceinline__ void bitonic_sort_desc(uint64_t (&a)[N]) {
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const int l = i ^ j;
        if (l > i) {
          if ((i & k) == 0)
            cmp_exchange<CEV>(a[i], a[l]);  // descending run: larger key first
          else
            cmp_exchange<CEV>(a[l], a[i]);
        }
      }
    }
  }
}
// Sort a bitonic sequence descending.
template <int N, int CEV = 0>
__device__ __forceinline__ void bitonic_merge_desc(uint64_t (&a)[N]) {
#pragma unroll
  for (int j = N >> 1; j > 0; j >>= 1) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const int l = i ^ j;
      if (l > i) cmp_exchange<CEV>(a[i], a[l]);
    }
  }
}

__device__ __forceinline__ uint64_t lds64(uint32_t addr) {
  uint32_t lo, hi;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr));
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint64_t v) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(static_cast<uint32_t>(v)),
               "r"(static_cast<uint32_t>(v >> 32))
               : "memory");
}

// Lock-step compaction: every lane merges the candidates of ITS OWN column (slots below `ptr`) into its sorted
// survivors.  Returns the lane's new 32nd-best value (0 while fewer than 32 positives have been seen) and, in t16,
// its 16th-best.
template <int CEV>
__device__ __noinline__ float compact_rows(uint64_t (&surv)[kTopK], uint32_t my_base, uint32_t ptr, float& t16) {
  uint64_t fresh[kNewSlots];
#pragma unroll
  for (int j = 0; j < kNewSlots; ++j) {
    const uint32_t addr = my_base + j * kSlotStride;
    fresh[j] = addr < ptr ? lds64(addr) : 0ull;
  }
  bitonic_sort_desc<kNewSlots, CEV>(fresh);
  // max(descending, reversed descending) = the 32 largest of the union, as a bitonic sequence
#pragma unroll
  for (int i = 0; i < kTopK; ++i) {
    uint64_t y = fresh[kNewSlots - 1 - i];
    cmp_exchange<CEV>(surv[i], y);  // keeps the larger; the smaller is dropped
  }
  bitonic_merge_desc<kTopK, CEV>(surv);
  t16 = __uint_as_float(static_cast<uint32_t>(surv[kTopK / 2 - 1] >> 32));
  return __uint_as_float(static_cast<uint32_t>(surv[kTopK - 1] >> 32));
}

// NBUF accumulator buffers of BN columns live in TMEM (NBUF * BN <= 512).  With two epilogue sets scanning two
// buffers, a THIRD buffer (BN = 160) lets the MMA issuer fill the next tile meanwhile: the tile rate becomes
// min(1/M, 2/E) instead of 2/(M + E)  (M = MMA time, E = scan + bias pre-store time of one tile).
// AMN / BMN: the operand is MN-major -- the global matrix is stored [K, M] (resp. [K, N]) row-major, e.g. a
// [tokens, features] activation matrix used as the A^T of a weight-gradient product whose K axis is the token axis.
// The TMA producer then fetches 64-column x 64-row boxes and the MMA reads them through MN-major descriptors, so no
// physical transpose of the operand is ever made (bf16 only).
template <int BN, int STAGES, int EPI, bool TF32, int SETS, int CL, int NBUF = 2, int CEV = 0, bool AMN = false,
          bool BMN = false>
__global__ void __launch_bounds__(128 + SETS * 128, 1)
sm100_gemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                  const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB1,
                  const __grid_constant__ CUtensorMap mapO, const __grid_constant__ CUtensorMap mapI,
                  const GemmParams p) {
  using L = GemmSmem<BN, STAGES, EPI, SETS, NBUF>;
  constexpr int kEpiWarps = L::kEpiWarps;
  constexpr uint16_t kMcMask = static_cast<uint16_t>((1u << CL) - 1u);
  constexpr int kBKe = TF32 ? 32 : 64;  // elements per 128-byte k-block
  constexpr uint32_t kTmemCols = NBUF * BN <= 32 ? 32 : NBUF * BN <= 64 ? 64 : NBUF * BN <= 128 ? 128
                                 : NBUF * BN <= 256 ? 256 : 512;  // allocation granularity: power of two
  static_assert(NBUF * BN <= 512, "accumulator buffers exceed the 512 TMEM columns");
  static_assert(kNewSlots == kTopK, "the merge step pairs survivor i with candidate 31-i");
  static_assert(!(AMN || BMN) || (!TF32 && CL == 1 && BN % 64 == 0), "MN-major operands: bf16, no multicast");
  const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0u;

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* ring = smem;
  uint8_t* stage_out = smem + L::kRing;
  uint8_t* stage_in = smem + L::kRing + L::kStageOut;
  constexpr int kIo = L::kStageOut + L::kStageIn;
  uint8_t* cand = smem + L::kRing + kIo;
  float* thr_s = reinterpret_cast<float*>(smem + L::kRing + kIo + L::kBuf);
  float* bias_s = reinterpret_cast<float*>(smem + L::kRing + kIo + L::kBuf + L::kThr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kRing + kIo + L::kBuf + L::kThr + L::kBiasS);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + NBUF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * NBUF);
  int* tile_ctr = reinterpret_cast<int*>(tmem_slot + 1);  // tiles the MMA issuer has started (read by the prefetch warp)
  uint64_t* in_bar = bars + 2 * STAGES + 2 * NBUF + 2;    // [epilogue warp][2]: input boxes landed

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_nt = (p.N + BN - 1) / BN;
  // this CTA's run of flattened tiles g = row_block * num_nt + column_tile: a whole row block, or one column range
  // of a row block of the split tail wave
  int64_t g_begin = static_cast<int64_t>(blockIdx.x) * num_nt, g_end = g_begin + num_nt;
  int piece = 0;
  if (p.tail_split > 1 && static_cast<int>(blockIdx.x) >= p.full_count) {
    const int tail = (p.M + kBM - 1) / kBM - p.full_count;
    const int t = static_cast<int>(blockIdx.x) - p.full_count;
    piece = t / tail;
    const int64_t base = static_cast<int64_t>(p.full_count + t - piece * tail) * num_nt;
    g_begin = base + static_cast<int64_t>(piece) * num_nt / p.tail_split;
    g_end = base + static_cast<int64_t>(piece + 1) * num_nt / p.tail_split;
  }
  if (p.persistent) {
    const int64_t num_mb = (p.M + kBM - 1) / kBM;
    g_begin = blockIdx.x * num_mb / gridDim.x * num_nt;
    g_end = (blockIdx.x + 1) * num_mb / gridDim.x * num_nt;
  }
  const int num_lt = static_cast<int>(g_end - g_begin);
  const bool is_piece = p.tail_split > 1 && static_cast<int>(blockIdx.x) >= p.full_count;
  const int total_kb = (p.K + kBKe - 1) / kBKe;
  const int split = p.kb_per_split > 0 ? static_cast<int>(blockIdx.y) : 0;
  const int kb0 = p.kb_per_split > 0 ? split * p.kb_per_split : 0;
  const int num_kb = p.kb_per_split > 0 ? max(0, min(total_kb - kb0, p.kb_per_split)) : total_kb;
  const int num_vk = num_kb * p.passes;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapB0);
    if (p.passes > 1) {
      tma_prefetch_desc(&mapA1);
      tma_prefetch_desc(&mapB1);
    }
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CL);  // every CTA of the cluster releases a slot its peers multicast into
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 4 * SETS);  // every epilogue warp drains its column range of buffer b
    }
    if constexpr (L::kHasIn) {
      for (int i = 0; i < kEpiWarps * 2; ++i) mbar_init(&in_bar[i], 1);
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  if (threadIdx.x < 2 * kBM) thr_s[threadIdx.x] = 0.f;  // published per-set 16th-largest values start at the ReLU floor
  if (threadIdx.x == 0) *tile_ctr = -1;
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();  // peers' barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int lt = 0; lt < num_lt; ++lt) {
        const int mb = static_cast<int>((g_begin + lt) / num_nt);
        const int nt = static_cast<int>(g_begin + lt - static_cast<int64_t>(mb) * num_nt);
        const int m0 = mb * kBM;
        for (int vk = 0; vk < num_vk; ++vk) {
          const int pass = vk / num_kb;
          const int kb = vk - pass * num_kb;
          // pass order (3-pass split): hi*lo, lo*hi, then the dominant hi*hi term
          const CUtensorMap* ma = (p.passes == 1 || pass != 1) ? &mapA0 : &mapA1;
          const CUtensorMap* mb = (p.passes == 1 || pass != 0) ? &mapB0 : &mapB1;
          if (p.flags & 2) mbar_wait(&empty_bar[stage], phase ^ 1);
          else mbar_wait_relaxed(&empty_bar[stage], phase ^ 1, 64);  // STAGES deep: nothing waits for this warp
          uint8_t* sa = ring + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
          if constexpr (AMN) {  // two 64-wide MN atoms of 64 K-rows each
#pragma unroll
            for (int h = 0; h < kBM / 64; ++h)
              tma_load_2d(sa + h * 8192, ma, &full_bar[stage], m0 + h * 64, (kb0 + kb) * kBKe, kEvictNormal);
          } else {
            tma_load_2d(sa, ma, &full_bar[stage], (kb0 + kb) * kBKe, m0, kEvictNormal);
          }
          if constexpr (BMN) {
#pragma unroll
            for (int h = 0; h < BN / 64; ++h)
              tma_load_2d(sb + h * 8192, mb, &full_bar[stage], nt * BN + h * 64, (kb0 + kb) * kBKe, kEvictLast);
          } else if constexpr (CL == 1) {
            tma_load_2d(sb, mb, &full_bar[stage], (kb0 + kb) * kBKe, nt * BN, kEvictLast);
          } else {
            // every CTA of the cluster walks the same weight tiles: fetch 1/CL of the tile, multicast it to all
            constexpr int kSlice = BN / CL;
            tma_load_2d_mc(sb + cta_rank * kSlice * kBKBytes, mb, &full_bar[stage], (kb0 + kb) * kBKe,
                           nt * BN + static_cast<int>(cta_rank) * kSlice, kMcMask, kEvictLast);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TF32 ? 2u : 1u, kBM, BN, AMN, BMN);
      int stage = 0;
      uint32_t phase = 0;
      for (int lt = 0; lt < num_lt; ++lt) {
        const int buf = lt % NBUF;
        // the epilogue arrives once it has read the buffer's previous tile out of TMEM (fresh buffers pass at once)
        if (p.flags & 4) mbar_wait(&tempty_bar[buf], ((lt / NBUF) & 1) ^ 1);
        else mbar_wait_relaxed(&tempty_bar[buf], ((lt / NBUF) & 1) ^ 1, 32);
        tc_fence_after();
        if constexpr (EPI == EPI_RESID || EPI == EPI_MASK) *reinterpret_cast<volatile int*>(tile_ctr) = lt;
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int vk = 0; vk < num_vk; ++vk) {
          if (p.flags & 8) mbar_wait(&full_bar[stage], phase);
          else mbar_wait_relaxed(&full_bar[stage], phase, 20);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {  // 4 x 32-byte UMMA_K steps per 128-byte k-block
            // K-major: 32 bytes along the 128-byte row per K = 16 step; MN-major: 16 K-rows of 128 bytes = 2048 B
            const uint64_t adesc = AMN ? make_mnmajor_sw128_desc(sa + k4 * 2048, 8192) : make_kmajor_sw128_desc(sa + k4 * 32);
            const uint64_t bdesc = BMN ? make_mnmajor_sw128_desc(sb + k4 * 2048, 8192) : make_kmajor_sw128_desc(sb + k4 * 32);
            const uint32_t acc = (vk > 0 || k4 > 0) ? 1u : 0u;  // the tile's first MMA overwrites the buffer
            if constexpr (TF32)
              mma_tf32_ss(d_tmem, adesc, bdesc, idesc, acc);
            else
              mma_f16_ss(d_tmem, adesc, bdesc, idesc, acc);
          }
          if constexpr (CL == 1)
            tc_commit(&empty_bar[stage]);
          else
            tc_commit_mc(&empty_bar[stage], kMcMask);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(&tfull_bar[buf]);
      }
    }
  } else if (warp_idx == 3) {
    // ===================== epilogue-input prefetcher (EPI_RESID / EPI_MASK) =====================
    // The epilogue threads read their OWN rows of the target / mask matrix (16-byte pieces, 32 rows per request):
    // latency-bound at DRAM distance (ncu: long_scoreboard, 18 % warps active).  This otherwise idle warp pulls the
    // NEXT tile's rows into L2 with bulk prefetches while the current tile is computed and drained.
    if constexpr (EPI == EPI_RESID || EPI == EPI_MASK) {
      const uint8_t* base = EPI == EPI_RESID ? reinterpret_cast<const uint8_t*>(p.target)
                                             : reinterpret_cast<const uint8_t*>(p.mask_src);
      const int64_t pitch = EPI == EPI_RESID ? p.ldt * 4 : p.ld16 * 2;  // bytes
      const int esz = EPI == EPI_RESID ? 4 : 2;
      const int64_t row_cols = EPI == EPI_RESID ? p.N : p.ld16;          // readable columns of a row
      const bool ok = (pitch & 15) == 0 && ((row_cols * esz) & 15) == 0 && p.tma_in == 0;
      for (int lt = 0; ok && lt < num_lt; ++lt) {
        // tile lt's rows are requested once the MMA issuer has reached tile lt - 1.  Paced by a monotone counter, not
        // by the accumulator barriers: a waiter that falls two phases behind a parity barrier would wait for a
        // completion that never comes at the end of the run.
        if (lt > 0) {
          uint32_t spins = 0;
          while (*reinterpret_cast<volatile int*>(tile_ctr) < lt - 1) {
            asm volatile("nanosleep.u32 %0;" ::"r"(200u));
            if (++spins > (1u << 22)) __trap();
          }
        }
        const int mb = static_cast<int>((g_begin + lt) / num_nt);
        const int nt = static_cast<int>(g_begin + lt - static_cast<int64_t>(mb) * num_nt);
        const int64_t c0 = static_cast<int64_t>(nt) * BN;
        const int64_t cols = min(static_cast<int64_t>(BN), row_cols - c0);
        if (cols <= 0) continue;
        const uint32_t bytes = static_cast<uint32_t>(cols * esz) & ~15u;
        for (int r = lane; r < kBM; r += 32) {
          const int64_t row = static_cast<int64_t>(mb) * kBM + r;
          if (row < p.M && bytes > 0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + row * pitch + c0 * esz), "r"(bytes)
                         : "memory");
        }
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== epilogue =====================
    const int q = warp_idx & 3;  // TMEM lane quarter
    const int ew = warp_idx - 4;
    const int set = SETS == 2 ? (ew >> 2) : 0;
    const int stid = q * 32 + lane;  // thread index within the set == token row within the row block
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    // Bias rows are staged PER WARP (two slots, filled one tile of this warp ahead): lane l fetches columns
    // [8l, 8l+8) of the row its warp scans next while the current tile is scanned and stores them behind a
    // __syncwarp.  No barrier ties the four warps of a set together, so a warp that has to compact does not hold
    // its three peers at the next tile boundary; they only meet through the accumulator hand-off, which has a
    // whole tile of slack.  The scan reads the row with broadcast loads and forms fl(accumulator + bias) per value.
    constexpr int kCols = L::kCols;            // this warp scans columns [cb, cb + kCols) of every tile
    constexpr int kPerLane = kCols / 32;       // bias columns a lane stages
    static_assert(kCols % 32 == 0 && kPerLane <= 8 && kPerLane % 4 == 0, "bias staging layout");
    const int cb = set * kCols;
    float* bias_w = bias_s + ew * (2 * kCols);
    float nb[kPerLane];
    auto load_bias = [&](int lt) {  // lt: CTA-local tile number; out-of-range columns get -inf
      const int nt = static_cast<int>((g_begin + lt) % num_nt);
#pragma unroll
      for (int j = 0; j < kPerLane; ++j) {
        const int gc = nt * BN + cb + lane * kPerLane + j;
        nb[j] = gc < p.N ? ((p.bias && split == 0) ? __ldg(p.bias + gc) : 0.f) : -INFINITY;
      }
    };
    auto store_bias = [&](int slot) {
      float4* dst = reinterpret_cast<float4*>(bias_w + slot * kCols + lane * kPerLane);
#pragma unroll
      for (int j = 0; j < kPerLane; j += 4) dst[j / 4] = make_float4(nb[j], nb[j + 1], nb[j + 2], nb[j + 3]);
    };

    // TMA output staging of this warp: two 32 x 128-byte boxes, SWIZZLE_128B (16-byte piece pc of row r sits at
    // r * 128 + ((pc ^ (r & 7)) * 16): conflict-free for the row-per-lane writes, undone by the tensor map)
    const bool tma_out = L::kHasOut && p.tma_out != 0;
    uint8_t* my_stage = stage_out + ew * (L::kOutBufs * L::kOutBox);
    int sbuf = 0;
    auto stage_piece = [&](int pc, uint4 v4) {
      *reinterpret_cast<uint4*>(my_stage + sbuf * L::kOutBox + lane * 128 + ((pc ^ (lane & 7)) << 4)) = v4;
    };
    auto stage_begin = [&]() {  // the box about to be written must have been read by the store that last used it
      if (lane == 0) bulk_wait_group_read<L::kOutBufs - 1>();
      __syncwarp();
    };
    // input boxes (EPI_RESID: 32 fp32 target columns; EPI_MASK: 64 bf16 gate columns; 32 rows x 128 bytes either way):
    // box g of this warp's sequence lives in buffer g & 1 and is requested while box g - 1 is consumed
    const bool tma_in = L::kHasIn && p.tma_in != 0;
    uint8_t* my_in = stage_in + ew * (2 * L::kOutBox);
    uint64_t* my_in_bar = in_bar + ew * 2;
    constexpr int kInCols = EPI == EPI_RESID ? 32 : 64;         // columns per input box
    constexpr int kInPerTile = L::kHasIn ? kCols / kInCols : 1;  // boxes per tile and warp
    uint32_t in_phase = 0u;  // bit b: phase parity of input buffer b
    auto in_issue = [&](int64_t g) {  // g: box ordinal over this CTA's tiles
      const int lt_g = static_cast<int>(g / kInPerTile);
      if (lt_g >= num_lt) return;
      const int bx = static_cast<int>(g - static_cast<int64_t>(lt_g) * kInPerTile);
      const int mb_g = static_cast<int>((g_begin + lt_g) / num_nt);
      const int nt_g = static_cast<int>(g_begin + lt_g - static_cast<int64_t>(mb_g) * num_nt);
      if (lane == 0) {
        uint64_t* b = my_in_bar + (g & 1);
        mbar_arrive_expect_tx(b, L::kOutBox);
        tma_load_2d(my_in + (g & 1) * L::kOutBox, &mapI, b, nt_g * BN + cb + bx * kInCols, mb_g * kBM + q * 32,
                    kEvictNormal);
      }
    };
    auto in_piece = [&](int64_t g, int pc) {
      return *reinterpret_cast<const uint4*>(my_in + (g & 1) * L::kOutBox + lane * 128 + ((pc ^ (lane & 7)) << 4));
    };
    int64_t in_g = 0;  // ordinal of the box being consumed
    if (tma_in) in_issue(0);
    auto stage_flush = [&](int col, int row0) {
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&mapO, my_stage + sbuf * L::kOutBox, col, row0);
        bulk_commit_group();
      }
      sbuf = (sbuf + 1) % L::kOutBufs;
    };
    static_assert(!L::kHasOut || kCols % 64 == 0, "output boxes are 32 fp32 / 64 bf16 columns wide");
    double acc_sum[3] = {0.0, 0.0, 0.0};  // EPI_RELU16 / EPI_RESID running sums of this thread's rows
    float aff_scale = 1.f, aff_shift = 0.f;
    if constexpr (EPI == EPI_MASK) {
      if (p.affine != nullptr) {
        aff_scale = __ldg(p.affine);
        aff_shift = __ldg(p.affine + 1);
      }
    }
    uint64_t surv[kTopK];
    float thresh = 0.f;  // candidates must be > thresh
    float t16 = 0.f;     // this set's 16th largest value so far (0 while it holds fewer than 16)
    const uint32_t my_base = smem_u32(cand) + ew * L::kBufPerWarp + lane * 8;
    const uint32_t ptr_limit = my_base + (kNewSlots - kCheck) * kSlotStride;
    uint32_t ptr = my_base;
    int bslot = 0;
    // Segments: maximal runs of this CTA's local tiles inside one row block; per-row state is reset at the start,
    // merged and emitted at the end.
    for (int lt0 = 0; lt0 < num_lt;) {
      const int mb = static_cast<int>((g_begin + lt0) / num_nt);
      const int nt0 = static_cast<int>(g_begin + lt0 - static_cast<int64_t>(mb) * num_nt);
      const int seg_len = min(num_lt - lt0, num_nt - nt0);
      const int seg_end = lt0 + seg_len;
      const int row = mb * kBM + stid;
      if constexpr (EPI == EPI_TOPK) {
#pragma unroll
        for (int s = 0; s < kTopK; ++s) surv[s] = 0ull;
        thresh = 0.f;
        t16 = 0.f;
        ptr = my_base;
        if (lt0 > 0) {
          // the values published for the previous row block must not leak into this one, and set 0 must be done
          // reading set 1's hand-off column before set 1 stages new candidates in it
          if constexpr (SETS == 2) thr_s[set * kBM + stid] = 0.f;
          asm volatile("bar.sync 3, %0;" ::"n"(kEpiWarps * 32) : "memory");
        }
      }
      {  // bias row of the segment's first tile
        load_bias(lt0);
        __syncwarp();
        store_bias(bslot);
        __syncwarp();
      }
      for (int lt = lt0; lt < seg_end; ++lt) {
      const int nt = nt0 + (lt - lt0);
      const int buf = lt % NBUF;
      if constexpr (EPI == EPI_TOPK) {
        // Lower bounds of the row's final 32nd largest value, beyond this set's own 32nd largest:
        //  * both sets see alike halves of the row (alternate 128-column runs), so min(own 16th, other's 16th) -- 16 + 16
        //    distinct values at least that large -- sits near the 32nd largest of the UNION, far above either set's
        //    own 32nd: the candidate count per row drops from 2 x 32(1 + ln(n/64)) towards 32(1 + ln(n/32));
        //  * the 32nd largest another column piece of this row block has published (tail split).
        // Values EQUAL to such a foreign bound may still win on the index tie-break, so the bound enters one ulp low.
        float foreign = 0.f;
        if constexpr (SETS == 2) {
          if (!(p.flags & 1)) foreign = fminf(t16, thr_s[(set ^ 1) * kBM + stid]);
        }
        if (is_piece && row < p.M) foreign = fmaxf(foreign, __ldcg(p.part_thr + row));
        if (foreign > 0.f) thresh = fmaxf(thresh, __uint_as_float(__float_as_uint(foreign) - 1u));
      }
      if (lt + 1 < seg_end) load_bias(lt + 1);  // lands in registers while this tile is scanned
      const uint32_t bs_addr = smem_u32(bias_w + bslot * kCols);
      if (p.flags & 16) mbar_wait(&tfull_bar[buf], (lt / NBUF) & 1);
      else mbar_wait_relaxed(&tfull_bar[buf], (lt / NBUF) & 1, 32);
      tc_fence_after();
      const uint32_t t_addr = lane_taddr + buf * BN + cb;
      uint32_t r[2][kChunk];
      // epilogue inputs of a chunk (EPI_RESID: 16 fp32 targets; EPI_MASK: 16 bf16 gate activations) are requested one
      // chunk ahead, like the TMEM loads, so their latency overlaps the previous chunk's arithmetic and stores
      float xin[2][EPI == EPI_RESID ? kChunk : 1];
      uint4 gin[2][EPI == EPI_MASK ? 2 : 1];
      const int64_t tile_col0 = static_cast<int64_t>(nt) * BN + cb;
      auto load_inputs = [&](int cc, int slot) {
        if (tma_in) return;  // the chunk's inputs are read from this warp's TMA box instead
        if constexpr (EPI == EPI_RESID) {
          const int64_t col0 = tile_col0 + cc;
          if (row < p.M) {
            const float* trow = p.target + static_cast<int64_t>(row) * p.ldt + col0;
            if (col0 + kChunk <= p.N && (p.ldt & 3) == 0) {
#pragma unroll
              for (int j = 0; j < kChunk; j += 4) {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(trow + j));
                xin[slot][j] = t4.x; xin[slot][j + 1] = t4.y; xin[slot][j + 2] = t4.z; xin[slot][j + 3] = t4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < kChunk; ++j) xin[slot][j] = col0 + j < p.N ? __ldg(trow + j) : 0.f;
            }
          }
        } else if constexpr (EPI == EPI_MASK) {
          const int64_t col0 = tile_col0 + cc;
          if (row < p.M) {
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8)
              if (col0 + h8 * 8 < p.ld16)
                gin[slot][h8] = __ldg(reinterpret_cast<const uint4*>(p.mask_src + static_cast<int64_t>(row) * p.ld16 +
                                                                     col0 + h8 * 8));
          }
        }
      };
      load_inputs(0, 0);
      tmem_ld_32x32b_x16(t_addr, r[0]);
#pragma unroll 1
      for (int c0 = 0; c0 < kCols; c0 += 2 * kChunk) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int cc = c0 + h * kChunk;  // column within this warp's range
          tmem_ld_wait();
          if (cc + kChunk < kCols) {
            tmem_ld_32x32b_x16(t_addr + cc + kChunk, r[h ^ 1]);  // prefetch the next chunk
            load_inputs(cc + kChunk, h ^ 1);
          } else {
            // this warp's columns now sit in registers: hand the buffer back to the MMA issuer before scanning the rest
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
          }
          // bias of these 16 columns (shared-memory broadcast reads): pre-activation = fl(accumulator + bias), the
          // reference's own order of operations (nn.Linear: x @ W.T, then + b)
          float v[kChunk];
#pragma unroll
          for (int j = 0; j < kChunk; j += 4) {
            const float4 b4 = lds128(bs_addr + (cc + j) * 4);
            v[j] = __uint_as_float(r[h][j]) + b4.x;
            v[j + 1] = __uint_as_float(r[h][j + 1]) + b4.y;
            v[j + 2] = __uint_as_float(r[h][j + 2]) + b4.z;
            v[j + 3] = __uint_as_float(r[h][j + 3]) + b4.w;
          }
          if constexpr (EPI == EPI_TOPK) {
            const uint32_t nidx0 = ~static_cast<uint32_t>(nt * BN + cb + cc);  // ~(col) == nidx0 - j
#pragma unroll
            for (int g = 0; g < kChunk; g += kCheck) {
#pragma unroll
              for (int j = g; j < g + kCheck; ++j) {
                // predicated store + UNpredicated pointer bump (selp/add into a fresh register): an in-place
                // predicated add would have to wait for the store to read its address register (WAR on the LSU)
                uint32_t next;
                asm volatile(
                    "{\n\t.reg .pred p;\n\t.reg .b32 inc;\n\t"
                    "setp.gt.f32 p, %2, %3;\n\t"
                    "@p st.shared.v2.b32 [%1], {%4, %5};\n\t"
                    "selp.b32 inc, %6, 0, p;\n\t"
                    "add.u32 %0, %1, inc;\n\t}"
                    : "=r"(next)
                    : "r"(ptr), "f"(v[j]), "f"(thresh), "r"(nidx0 - j), "r"(__float_as_uint(v[j])),
                      "n"(kSlotStride)
                    : "memory");
                ptr = next;
              }
              // the next kCheck columns could overflow some lane's column: all lanes of the WARP compact their rows
              if (__any_sync(0xffffffffu, ptr > ptr_limit)) {
                const float t = compact_rows<CEV>(surv, my_base, ptr, t16);
                thresh = fmaxf(thresh, t);
                ptr = my_base;
                if constexpr (SETS == 2) thr_s[set * kBM + stid] = t16;
                if (is_piece && row < p.M)  // non-negative floats order like their bit patterns
                  atomicMax(reinterpret_cast<int*>(p.part_thr + row), __float_as_int(t));
              }
            }
          } else if constexpr (EPI == EPI_NONE) {
            if (v[0] == 1.2345e38f) p.out[0] = 1.f;  // keep the loads alive
          } else if constexpr (EPI == EPI_RELU16 || EPI == EPI_RESID || EPI == EPI_MASK) {
            // three epilogues with a bf16 output row [M, ld16]: o[j] = the chunk's 16 output values
            const bool rv = row < p.M;
            const int64_t col0 = static_cast<int64_t>(nt) * BN + cb + cc;
            float o[kChunk];
            constexpr int kChunksPerIn = kInCols / kChunk;
            const int cib = (cc / kChunk) % kChunksPerIn;  // chunk within the current input box
            if constexpr (L::kHasIn) {
              if (tma_in && cib == 0) {
                __syncwarp();          // every lane is done with the other buffer (the previous box)
                in_issue(in_g + 1);    // request the next box, then wait for this one
                mbar_wait(my_in_bar + (in_g & 1), (in_phase >> (in_g & 1)) & 1u);
                in_phase ^= 1u << (in_g & 1);
              }
              if (tma_in) {
                if constexpr (EPI == EPI_RESID) {
#pragma unroll
                  for (int j = 0; j < kChunk; j += 4) {
                    const uint4 t4 = in_piece(in_g, cib * 4 + j / 4);
                    xin[h][j] = __uint_as_float(t4.x); xin[h][j + 1] = __uint_as_float(t4.y);
                    xin[h][j + 2] = __uint_as_float(t4.z); xin[h][j + 3] = __uint_as_float(t4.w);
                  }
                } else {
                  gin[h][0] = in_piece(in_g, cib * 2);
                  gin[h][1] = in_piece(in_g, cib * 2 + 1);
                }
              }
            }
            // The running sums are fp64, but a chunk's 16 terms are first added in fp32 (four interleaved partial sums)
            // and enter the fp64 sum once: per element that is one FADD instead of a conversion + a DADD on a serial
            // chain, and no branch (the epilogue was bound by instruction issue and fixed-latency dependencies).
            if constexpr (EPI == EPI_RELU16) {
              float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int j = 0; j < kChunk; ++j) {
                v[j] = fmaxf(v[j], 0.f);  // columns past N carry a -inf bias: they come out as 0
                o[j] = v[j];
                s4[j & 3] += v[j];
              }
              if (rv) acc_sum[0] += static_cast<double>((s4[0] + s4[1]) + (s4[2] + s4[3]));
            } else if constexpr (EPI == EPI_RESID) {
              const float (&xt)[kChunk] = xin[h];
              // valid columns of this chunk for this thread's row: kChunk everywhere but at the matrix edges
              const int ncol = static_cast<int>(p.N) - static_cast<int>(col0);
              const int nval = rv ? min(max(ncol, 0), kChunk) : 0;
              bool ign = false;  // mse_loss(..., ignored_index=-1): exact comparison, as the reference (:31)
#pragma unroll
              for (int j = 0; j < kChunk; ++j) ign |= __float_as_uint(xt[j]) == 0xbf800000u;  // the one encoding of -1
              if (__all_sync(0xffffffffu, nval == kChunk && !ign)) {
                // the common chunk -- every column valid, no ignored target: a subtraction and an FMA per element
                float sa[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < kChunk; ++j) {
                  const float d0 = v[j] - xt[j];
                  sa[j & 3] = fmaf(d0, d0, sa[j & 3]);
                  o[j] = d0;
                }
                const double sq = static_cast<double>((sa[0] + sa[1]) + (sa[2] + sa[3]));
                acc_sum[2] += sq;
                acc_sum[0] += sq;
                acc_sum[1] += static_cast<double>(kChunk);
              } else if (__all_sync(0xffffffffu, nval == 0)) {  // column range past N / rows past M
#pragma unroll
                for (int j = 0; j < kChunk; ++j) o[j] = 0.f;
              } else {
                float sa[4] = {0.f, 0.f, 0.f, 0.f}, sm[4] = {0.f, 0.f, 0.f, 0.f};
                int cnt = 0;
#pragma unroll
                for (int j = 0; j < kChunk; ++j) {
                  float xv = xt[j];
                  asm volatile("" : "+f"(xv));  // keeps this rare path's compares out of the common path's schedule
                  const float d0 = v[j] - xv;
                  const float d2 = d0 * d0;
                  const bool ok = j < nval;
                  const bool keep = ok && xv != -1.0f;
                  sa[j & 3] += ok ? d2 : 0.f;
                  sm[j & 3] += keep ? d2 : 0.f;
                  cnt += keep ? 1 : 0;
                  o[j] = keep ? d0 : 0.f;
                }
                acc_sum[2] += static_cast<double>((sa[0] + sa[1]) + (sa[2] + sa[3]));
                acc_sum[0] += static_cast<double>((sm[0] + sm[1]) + (sm[2] + sm[3]));
                acc_sum[1] += static_cast<double>(cnt);
              }
            } else {
#pragma unroll
              for (int h8 = 0; h8 < kChunk; h8 += 8) {
                const uint4 a8 = (rv && col0 + h8 < p.ld16) ? gin[h][h8 / 8] : make_uint4(0, 0, 0, 0);
                const uint32_t aw[4] = {a8.x, a8.y, a8.z, a8.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  // bf16 > 0  <=>  sign clear and not zero: as signed integers, (low half << 16) > 0 and word >= 2^16
                  const bool p0 = static_cast<int32_t>(aw[j] << 16) > 0;
                  const bool p1 = static_cast<int32_t>(aw[j]) >= 0x10000;
                  o[h8 + 2 * j] = p0 ? fmaf(aff_scale, v[h8 + 2 * j], aff_shift) : 0.f;
                  o[h8 + 2 * j + 1] = p1 ? fmaf(aff_scale, v[h8 + 2 * j + 1], aff_shift) : 0.f;
                }
              }
            }
            const int grp = (cc / kChunk) & 3;  // chunk within the 64-column (128-byte) output box
#pragma unroll
            for (int h8 = 0; h8 < kChunk; h8 += 8) {
              uint4 q;
              uint32_t* qw = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const __nv_bfloat162 hh = __floats2bfloat162_rn(o[h8 + 2 * j], o[h8 + 2 * j + 1]);
                qw[j] = *reinterpret_cast<const uint32_t*>(&hh);
              }
              if (tma_out) {
                if (grp == 0 && h8 == 0) stage_begin();
                stage_piece(grp * 2 + h8 / 8, q);
              } else if (rv && col0 + h8 < p.ld16) {  // ld16 % 8 == 0: a group of 8 is inside the pitch or outside
                *reinterpret_cast<uint4*>(p.out16 + static_cast<int64_t>(row) * p.ld16 + col0 + h8) = q;
              }
            }
            if (tma_out && grp == 3)
              stage_flush(static_cast<int>(col0) - 3 * kChunk, mb * kBM + q * 32);
            if constexpr (L::kHasIn) {
              if (tma_in && cib == kChunksPerIn - 1) ++in_g;
            }
            if constexpr (EPI != EPI_MASK) {
              if (rv && p.out != nullptr) {  // optional fp32 copy (module API); the trainer does not ask for it
#pragma unroll
                for (int j = 0; j < kChunk; ++j)
                  if (col0 + j < p.N) p.out[static_cast<int64_t>(row) * p.ldo + col0 + j] = v[j];
              }
            }
          } else {
            const bool rv = row < p.M;
            const int64_t col0 = static_cast<int64_t>(nt) * BN + cb + cc;
            if (tma_out) {
              const int grp = (cc / kChunk) & 1;  // chunk within the 32-column (128-byte) output box
              if (grp == 0) stage_begin();
              if (col0 + kChunk > p.N) {
                // TMA clips a box in 16-byte granules: up to three padding columns of the pitch that share a granule
                // with the last valid column are written too -- as zeros, not as the -inf of an out-of-range column
#pragma unroll
                for (int j = 0; j < kChunk; ++j)
                  if (col0 + j >= p.N) v[j] = 0.f;
              }
#pragma unroll
              for (int j = 0; j < kChunk; j += 4) {
                float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                if (p.relu) {
                  o.x = fmaxf(o.x, 0.f);
                  o.y = fmaxf(o.y, 0.f);
                  o.z = fmaxf(o.z, 0.f);
                  o.w = fmaxf(o.w, 0.f);
                }
                stage_piece(grp * 4 + j / 4, make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z),
                                                        __float_as_uint(o.w)));
              }
              if (grp == 1) stage_flush(static_cast<int>(col0) - kChunk, mb * kBM + q * 32);
            } else if (rv) {
              float* orow = p.out + split * p.split_stride + static_cast<int64_t>(row) * p.ldo + col0;
              const bool full_chunk = (col0 + kChunk <= p.N) && ((p.ldo & 3) == 0);
              if (full_chunk) {
#pragma unroll
                for (int j = 0; j < kChunk; j += 4) {
                  float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                  if (p.relu) {
                    o.x = fmaxf(o.x, 0.f);
                    o.y = fmaxf(o.y, 0.f);
                    o.z = fmaxf(o.z, 0.f);
                    o.w = fmaxf(o.w, 0.f);
                  }
                  *reinterpret_cast<float4*>(orow + j) = o;
                }
              } else {
#pragma unroll
                for (int j = 0; j < kChunk; ++j) {
                  if (col0 + j < p.N) {
                    float o = v[j];
                    if (p.relu) o = fmaxf(o, 0.f);
                    orow[j] = o;
                  }
                }
              }
            }
          }
        }
      }
      // the bias row of the next tile goes into the other slot (its last reader was this warp, one tile ago)
      if (lt + 1 < seg_end) {
        store_bias(bslot ^ 1);
        __syncwarp();
        bslot ^= 1;
      }
      }
    if constexpr (EPI == EPI_TOPK) {
      // final compaction; with two sets, set 1 hands its survivors to set 0 through its (now idle) candidate
      // column and set 0 merges; then each thread emits its own row.  Short rows (fewer than 32 positive
      // pre-activations) are completed with zeros at the lowest indices not already chosen, which is the oracle's
      // (value desc, index asc) order for the all-zero tail after ReLU.
      compact_rows<CEV>(surv, my_base, ptr, t16);
      if constexpr (SETS == 2) {
        if (set == 1) {
#pragma unroll
          for (int s = 0; s < kTopK; ++s) sts64(my_base + s * kSlotStride, surv[s]);
        }
        asm volatile("bar.sync 3, %0;" ::"n"(kEpiWarps * 32) : "memory");
        if (set == 0) {
          const uint32_t peer = my_base + 4 * L::kBufPerWarp;
#pragma unroll
          for (int i = 0; i < kTopK; ++i) {
            uint64_t y = lds64(peer + (kTopK - 1 - i) * kSlotStride);
            cmp_exchange<CEV>(surv[i], y);
          }
          bitonic_merge_desc<kTopK, CEV>(surv);
        }
      }
      if (set == 0 && row < p.M) {
        const bool whole = nt0 == 0 && seg_len == num_nt;  // otherwise one of several pieces of this row block
        const uint32_t col_base = static_cast<uint32_t>(nt0) * BN;
        int nvalid = 0;
#pragma unroll
        for (int s = 0; s < kTopK; ++s) nvalid += surv[s] != 0ull ? 1 : 0;
        float* ov = p.top_vals + static_cast<int64_t>(row) * kTopK;
        int32_t* oi = p.top_idx + static_cast<int64_t>(row) * kTopK;
        if (!whole) {
          ov = p.part_vals + piece * p.part_stride + static_cast<int64_t>(row) * kTopK;
          oi = p.part_idx + piece * p.part_stride + static_cast<int64_t>(row) * kTopK;
        }
        if (nvalid == kTopK) {
#pragma unroll
          for (int s = 0; s < kTopK; s += 4) {
            *reinterpret_cast<float4*>(ov + s) =
                make_float4(__uint_as_float(static_cast<uint32_t>(surv[s] >> 32)),
                            __uint_as_float(static_cast<uint32_t>(surv[s + 1] >> 32)),
                            __uint_as_float(static_cast<uint32_t>(surv[s + 2] >> 32)),
                            __uint_as_float(static_cast<uint32_t>(surv[s + 3] >> 32)));
            *reinterpret_cast<int4*>(oi + s) =
                make_int4(static_cast<int>(~static_cast<uint32_t>(surv[s])),
                          static_cast<int>(~static_cast<uint32_t>(surv[s + 1])),
                          static_cast<int>(~static_cast<uint32_t>(surv[s + 2])),
                          static_cast<int>(~static_cast<uint32_t>(surv[s + 3])));
          }
        } else {
          // survivors are sorted, so the valid ones are surv[0 .. nvalid); the tail takes the free indices
          // (a piece fills from its own first column: across pieces the merge keeps the lowest indices)
          uint64_t taken = 0;  // membership of indices [col_base, col_base + 64) in the valid set
#pragma unroll
          for (int s = 0; s < kTopK; ++s) {
            const uint32_t si = ~static_cast<uint32_t>(surv[s]) - col_base;
            if (surv[s] != 0ull && si < 64) taken |= 1ull << si;
          }
          uint64_t free_mask = ~taken;
#pragma unroll
          for (int s = 0; s < kTopK; ++s) {
            float val = __uint_as_float(static_cast<uint32_t>(surv[s] >> 32));
            uint32_t idx = ~static_cast<uint32_t>(surv[s]);
            if (surv[s] == 0ull) {
              idx = col_base + __ffsll(static_cast<long long>(free_mask)) - 1;
              free_mask &= free_mask - 1;
              val = 0.f;
            }
            ov[s] = val;
            oi[s] = static_cast<int32_t>(idx);
          }
        }
      }
    }
      lt0 = seg_end;
    }
    if (tma_out && lane == 0) bulk_wait_group_read<0>();  // the boxes must outlive the stores that read them
    if constexpr (EPI == EPI_RELU16 || EPI == EPI_RESID) {
      constexpr int kSums = EPI == EPI_RELU16 ? 1 : 3;
#pragma unroll
      for (int i = 0; i < kSums; ++i) {
        const double t = warp_sum_f64(acc_sum[i]);
        if (lane == 0 && t != 0.0) atomicAdd(p.sums + i, t);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();  // no CTA exits while peers may still signal its barriers
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace freud
