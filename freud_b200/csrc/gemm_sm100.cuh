// Warp-specialised tcgen05 GEMM for sm_100a with two fused epilogues.
//
//   D[M, N] = A[M, K] * B[N, K]^T   (both operands K-major, i.e. row-major with K contiguous)
//
//   * A is the (centred) activation block  x - b_dec      [tokens, d]
//   * B is the encoder weight              W_enc          [features, d]
//
// One CTA owns a 128-token row block and walks over ALL feature tiles (BN columns each), so each
// epilogue thread (TMEM lane == token row) sees every pre-activation of its token exactly once:
//   EPI_TOPK : bias + ReLU + streaming exact top-32 selection per token, never materialising [M, N]
//              (reference: TopKAutoEncoder.pre_acts + select_topk, topkautoencoder.py:72-85)
//   EPI_STORE: bias (+ReLU) and store fp32 [M, N]  (reference: pre_acts / L1 encode + decode GEMMs)
//
// Roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warp 3 idle,
// warps 4-11 = epilogue in two sets of four (TMEM lane quarter == warp_idx % 4); set s drains accumulator
// buffer s (tiles s, s+2, ...), so two epilogue warps share every SM sub-partition and hide each other's latency.
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM accumulator double buffer full/empty (MMA <-> epilogue).
//
// Precision: kind::f16 with bf16 operands (1 pass), or kind::tf32 with the 3-pass split
//   A*B ~= A_hi*B_lo + A_lo*B_hi + A_hi*B_hi   (hi/lo are exact tf32 values prepared by prep kernels)
// which recovers ~fp32 accuracy (the reference's fp32 path) on the tensor cores.
#pragma once
#include "ptx.cuh"

namespace freud {

constexpr int kBM = 128;          // token rows per CTA (UMMA M)
constexpr int kBKBytes = 128;     // one swizzle-128B row per k-block
constexpr int kTopK = 32;         // fused selection width (one survivor per lane)
constexpr int kNewSlots = 28;     // unsorted candidate slots per token between compactions
constexpr int kCheckEvery = 8;    // columns between buffer-occupancy checks
// epilogue warps come in SETS (1 or 2) of 4 (one warp per TMEM lane quarter); with 2 sets, set s drains
// accumulator buffer s.  threads = 128 (producer / MMA / alloc / spare) + SETS * 128

enum { EPI_TOPK = 0, EPI_STORE = 1, EPI_NONE = 2 };  // EPI_NONE: mainloop-ceiling probe, discards the tile

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
};

template <int BN, int STAGES, int EPI, int SETS>
struct GemmSmem {
  static constexpr int kEpiWarps = 4 * SETS;
  static constexpr int kThreads = 128 + kEpiWarps * 32;
  static constexpr int kABytes = kBM * kBKBytes;
  static constexpr int kBBytes = BN * kBKBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kRing = STAGES * kStageBytes;
  // candidate buffers: per epilogue warp, (32 sorted + kNewSlots new) slots x 33 lanes (padded) x 8 B
  static constexpr int kSlots = kTopK + kNewSlots;
  static constexpr int kBufPerWarp = kSlots * 33 * 8;
  static constexpr int kBuf = EPI == 0 ? kEpiWarps * kBufPerWarp : 0;
  static constexpr int kBias = 4 * BN * 4;             // [set][parity][BN]
  static constexpr int kThr = 2 * kBM * 4;             // per-set published thresholds
  static constexpr int kBars = (2 * STAGES + 4) * 8 + 16;
  static constexpr int kTotal = kRing + kBuf + kBias + kThr + kBars;
};

// 64-bit candidate key: high word = fp32 bits of a strictly positive value, low word = ~index.
// Unsigned order == (value descending, index ascending) order.  0 == empty slot.
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
  uint32_t lo = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(v), m);
  uint32_t hi = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), m);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_idx_u64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v), src);
  uint32_t hi = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), src);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
// Sort 32 keys (one per lane) descending by lane index.
__device__ __forceinline__ uint64_t warp_sort_desc(uint64_t key, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      uint64_t other = shfl_xor_u64(key, j);
      bool desc_block = (lane & k) == 0;  // k == 32: always true
      bool lower = (lane & j) == 0;
      bool take_max = (lower == desc_block);
      uint64_t mx = key > other ? key : other;
      uint64_t mn = key > other ? other : key;
      key = take_max ? mx : mn;
    }
  }
  return key;
}
// Sort a bitonic sequence of 32 keys descending.
__device__ __forceinline__ uint64_t warp_bitonic_merge_desc(uint64_t key, int lane) {
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    uint64_t other = shfl_xor_u64(key, j);
    bool lower = (lane & j) == 0;
    uint64_t mx = key > other ? key : other;
    uint64_t mn = key > other ? other : key;
    key = lower ? mx : mn;
  }
  return key;
}

// Merge the `ncnt` new candidates of token-lane `owner` into its sorted survivors; returns the new
// threshold (value of the 32nd survivor, 0 if fewer than 32).  Warp-cooperative, conflict-free
// thanks to the 33-lane padding of the [slot][lane] layout.
__device__ __noinline__ float compact_one(uint64_t* wbuf, int owner, int ncnt, int lane) {
  uint64_t old_key = wbuf[lane * 33 + owner];
  uint64_t new_key = lane < ncnt ? wbuf[(kTopK + lane) * 33 + owner] : 0ull;
  new_key = warp_sort_desc(new_key, lane);
  uint64_t rev = shfl_idx_u64(new_key, 31 - lane);
  uint64_t merged = old_key > rev ? old_key : rev;
  merged = warp_bitonic_merge_desc(merged, lane);
  wbuf[lane * 33 + owner] = merged;
  uint32_t kth_hi = __shfl_sync(0xffffffffu, static_cast<uint32_t>(merged >> 32), 31);
  return __uint_as_float(kth_hi);
}

template <int BN, int STAGES, int EPI, bool TF32, int SETS, int CL>
__global__ void __launch_bounds__(128 + SETS * 128, 1)
sm100_gemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                  const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB1,
                  const GemmParams p) {
  using L = GemmSmem<BN, STAGES, EPI, SETS>;
  constexpr int kEpiWarps = L::kEpiWarps;
  constexpr uint16_t kMcMask = static_cast<uint16_t>((1u << CL) - 1u);
  const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0u;
  constexpr int kBKe = TF32 ? 32 : 64;  // elements per 128-byte k-block
  constexpr uint32_t kTmemCols = 2 * BN;
  static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM columns must be pow2 <= 512");

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* ring = smem;
  uint64_t* cand = reinterpret_cast<uint64_t*>(smem + L::kRing);
  float* bias_s = reinterpret_cast<float*>(smem + L::kRing + L::kBuf);
  float* thr_s = reinterpret_cast<float*>(smem + L::kRing + L::kBuf + L::kBias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kRing + L::kBuf + L::kBias + L::kThr);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBM;
  const int num_nt = (p.N + BN - 1) / BN;
  const int num_kb = (p.K + kBKe - 1) / kBKe;
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
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 4);
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  if (threadIdx.x < 2 * kBM) thr_s[threadIdx.x] = 0.f;  // published per-set thresholds start at the ReLU floor
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
      for (int nt = 0; nt < num_nt; ++nt) {
        for (int vk = 0; vk < num_vk; ++vk) {
          const int pass = vk / num_kb;
          const int kb = vk - pass * num_kb;
          // pass order (3-pass split): hi*lo, lo*hi, then the dominant hi*hi term
          const CUtensorMap* ma = (p.passes == 1 || pass != 1) ? &mapA0 : &mapA1;
          const CUtensorMap* mb = (p.passes == 1 || pass != 0) ? &mapB0 : &mapB1;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = ring + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
          tma_load_2d(sa, ma, &full_bar[stage], kb * kBKe, m0, kEvictNormal);
          if constexpr (CL == 1) {
            tma_load_2d(sb, mb, &full_bar[stage], kb * kBKe, nt * BN, kEvictLast);
          } else {
            // every CTA of the cluster walks the same weight tiles: fetch 1/CL of the tile, multicast it to all
            constexpr int kSlice = BN / CL;
            tma_load_2d_mc(sb + cta_rank * kSlice * kBKBytes, mb, &full_bar[stage], kb * kBKe,
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
      constexpr uint32_t idesc = make_idesc(TF32 ? 2u : 1u, kBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int nt = 0; nt < num_nt; ++nt) {
        const int buf = nt & 1;
        mbar_wait(&tempty_bar[buf], ((nt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int vk = 0; vk < num_vk; ++vk) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {  // 4 x 32-byte UMMA_K steps per 128-byte k-block
            const uint64_t adesc = make_kmajor_sw128_desc(sa + k4 * 32);
            const uint64_t bdesc = make_kmajor_sw128_desc(sb + k4 * 32);
            if constexpr (TF32)
              mma_tf32_ss(d_tmem, adesc, bdesc, idesc, (vk | k4) != 0);
            else
              mma_f16_ss(d_tmem, adesc, bdesc, idesc, (vk | k4) != 0);
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
  } else if (warp_idx >= 4) {
    // ===================== epilogue =====================
    // Two sets of four warps; set s owns accumulator buffer s, i.e. tiles nt = s, s+2, ...  Within a set, warp
    // quarter q reads TMEM lanes [32q, 32q+32): one token row per thread.
    const int q = warp_idx & 3;
    const int ew = warp_idx - 4;
    const int set = SETS == 2 ? (ew >> 2) : 0;
    const int stid = q * 32 + lane;  // thread index within the set
    const int row = m0 + q * 32 + lane;
    uint64_t* wbuf = cand + ew * (L::kSlots * 33);
    float thresh = 0.f;
    // shared-memory address of this token-lane's next free "new" slot (slot stride = 33 lanes * 8 B)
    const uint32_t warp_new_base = smem_u32(wbuf + kTopK * 33);
    const uint32_t my_new_base = warp_new_base + lane * 8;
    const uint32_t ptr_limit = my_new_base + (kNewSlots - kCheckEvery) * 33 * 8;
    uint32_t ptr = my_new_base;
    if constexpr (EPI == EPI_TOPK) {
#pragma unroll 4
      for (int s = 0; s < kTopK; ++s) wbuf[s * 33 + lane] = 0ull;
      __syncwarp();
    }
    for (int nt = set; nt < num_nt; nt += SETS) {
      const int buf = nt & 1;
      const int it = nt >> 1;  // per-buffer tile counter
      // stage this tile's bias (or -inf for out-of-range columns so they can never be selected)
      float* bs = bias_s + (buf * 2 + (it & 1)) * BN;
      for (int c = stid; c < BN; c += 128) {
        const int gc = nt * BN + c;
        float b = 0.f;
        if (gc < p.N) {
          if (p.bias) b = __ldg(p.bias + gc);
        } else {
          b = -INFINITY;
        }
        bs[c] = b;
      }
      if (set == 0)
        asm volatile("bar.sync 1, 128;" ::: "memory");
      else
        asm volatile("bar.sync 2, 128;" ::: "memory");
      if constexpr (EPI == EPI_TOPK && SETS == 2) {
        // any lower bound of the row's 32nd largest value is a valid filter: adopt the other set's if tighter
        thresh = fmaxf(thresh, thr_s[(set ^ 1) * kBM + stid]);
      }
      mbar_wait(&tfull_bar[buf], it & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN;
      const uint32_t bs_addr = smem_u32(bs);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_addr + c0, r);
        float bb[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {  // bias loads overlap the TMEM load
          const float4 b4 = lds128(bs_addr + (c0 + j) * 4);
          bb[j] = b4.x; bb[j + 1] = b4.y; bb[j + 2] = b4.z; bb[j + 3] = b4.w;
        }
        tmem_ld_wait();
        if constexpr (EPI == EPI_TOPK) {
          const uint32_t nidx0 = ~static_cast<uint32_t>(nt * BN + c0);  // ~(col) == nidx0 - j
#pragma unroll
          for (int g = 0; g < 32; g += kCheckEvery) {
            float v[kCheckEvery];
#pragma unroll
            for (int jj = 0; jj < kCheckEvery; ++jj) v[jj] = __uint_as_float(r[g + jj]) + bb[g + jj];
#pragma unroll
            for (int jj = 0; jj < kCheckEvery; ++jj) {
              if (v[jj] > thresh) {
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(ptr), "r"(nidx0 - (g + jj)),
                             "r"(__float_as_uint(v[jj]))
                             : "memory");
                ptr += 33 * 8;
              }
            }
            // compaction round for lanes that could overflow in the next kCheckEvery columns
            uint32_t need = __ballot_sync(0xffffffffu, ptr > ptr_limit);
            if (need) {
              __syncwarp();
              while (need) {
                const int owner = __ffs(need) - 1;
                need &= need - 1;
                const int ncnt = (__shfl_sync(0xffffffffu, ptr, owner) - (warp_new_base + owner * 8)) / (33 * 8);
                const float t = compact_one(wbuf, owner, ncnt, lane);
                if (lane == owner) {
                  thresh = fmaxf(thresh, t);
                  ptr = my_new_base;
                  thr_s[set * kBM + stid] = thresh;
                }
              }
              __syncwarp();
            }
          }
        } else if constexpr (EPI == EPI_NONE) {
          if (__uint_as_float(r[0]) + bb[0] == 1.2345e38f) p.out[0] = 1.f;  // keep the loads alive
        } else {
          if (row < p.M) {
            float* orow = p.out + static_cast<int64_t>(row) * p.ldo + nt * BN + c0;
            const bool full_chunk = (nt * BN + c0 + 32 <= p.N) && ((p.ldo & 3) == 0);
            if (full_chunk) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 o;
                o.x = __uint_as_float(r[j + 0]) + bb[j + 0];
                o.y = __uint_as_float(r[j + 1]) + bb[j + 1];
                o.z = __uint_as_float(r[j + 2]) + bb[j + 2];
                o.w = __uint_as_float(r[j + 3]) + bb[j + 3];
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
              for (int j = 0; j < 32; ++j) {
                if (nt * BN + c0 + j < p.N) {
                  float o = __uint_as_float(r[j]) + bb[j];
                  if (p.relu) o = fmaxf(o, 0.f);
                  orow[j] = o;
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
    }
    if constexpr (EPI == EPI_TOPK) {
      // final compaction of every token-lane in both sets, then merge the two sets' sorted survivors and emit
      // (value, index) rows; short rows (fewer than 32 positive pre-activations) are completed with zeros at the
      // lowest indices not already chosen, which is the oracle's (value desc, index asc) order for the all-zero
      // tail after ReLU.
      __syncwarp();
      for (int owner = 0; owner < 32; ++owner) {
        const int ncnt = (__shfl_sync(0xffffffffu, ptr, owner) - (warp_new_base + owner * 8)) / (33 * 8);
        if (ncnt > 0) compact_one(wbuf, owner, ncnt, lane);
      }
      asm volatile("bar.sync 3, %0;" ::"n"(kEpiWarps * 32) : "memory");
      const uint64_t* bufA = cand + q * (L::kSlots * 33);        // set 0, this lane quarter
      const uint64_t* bufB = cand + (4 + q) * (L::kSlots * 33);  // set 1, this lane quarter
      for (int o = 0; o < 32 / SETS; ++o) {
        const int owner = set * (32 / SETS) + o;
        const int orow = m0 + q * 32 + owner;
        if (orow >= p.M) break;  // warp-uniform
        uint64_t key = bufA[lane * 33 + owner];
        if constexpr (SETS == 2) {
          const uint64_t kb = bufB[(31 - lane) * 33 + owner];
          // max(desc, reversed desc) = the 32 largest of the union as a bitonic sequence -> sort it
          key = warp_bitonic_merge_desc(key > kb ? key : kb, lane);
        }
        float val = __uint_as_float(static_cast<uint32_t>(key >> 32));
        uint32_t idx = ~static_cast<uint32_t>(key);
        const uint32_t valid = __ballot_sync(0xffffffffu, key != 0ull);
        if (valid != 0xffffffffu) {
          // survivors are sorted, so valid lanes are [0, nvalid); lanes >= nvalid take free indices
          const int nvalid = __popc(valid);
          uint32_t taken_lo = 0, taken_hi = 0;  // membership of indices [0,32) and [32,64) in the valid set
          for (int s = 0; s < nvalid; ++s) {
            const uint32_t si = __shfl_sync(0xffffffffu, idx, s);
            if (si < 32) taken_lo |= 1u << si;
            else if (si < 64) taken_hi |= 1u << (si - 32);
          }
          if (lane >= nvalid) {
            int want = lane - nvalid;  // rank among free indices
            uint32_t free_lo = ~taken_lo;
            const int nfree_lo = __popc(free_lo);
            uint32_t pick;
            if (want < nfree_lo) {
              for (int t = 0; t < want; ++t) free_lo &= free_lo - 1;
              pick = __ffs(free_lo) - 1;
            } else {
              uint32_t free_hi = ~taken_hi;
              want -= nfree_lo;
              for (int t = 0; t < want; ++t) free_hi &= free_hi - 1;
              pick = 32 + __ffs(free_hi) - 1;
            }
            idx = pick;
            val = 0.f;
          }
        }
        p.top_vals[static_cast<int64_t>(orow) * kTopK + lane] = val;
        p.top_idx[static_cast<int64_t>(orow) * kTopK + lane] = static_cast<int32_t>(idx);
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
