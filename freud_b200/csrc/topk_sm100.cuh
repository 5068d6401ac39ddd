// Fused encoder GEMM + ReLU + exact top-32 with SPECIALISED epilogue warps (reference: TopKAutoEncoder.pre_acts +
// select_topk, topkautoencoder.py:72-85).  Same TMA / tcgen05 / TMEM mainloop as sm100_gemm_kernel; the epilogue
// is split into two roles per TMEM lane quarter (= SM sub-partition):
//
//   scanner   warp 4+q : reads the accumulator tile out of TMEM, forms fl(acc + bias), and appends every value above
//                        its row's threshold to a shared-memory candidate column (predicated store, no divergence).
//                        Its work per tile is the same whatever the data: ~7 instructions per value.
//   compactor warp 8+q : owns the 32 sorted survivors of each row IN REGISTERS, folds handed-over candidate columns
//                        into them with the register-resident bitonic networks, and publishes the new thresholds.
//
// The two meet through a double-buffered candidate column per quarter (mbarrier full / empty pair per buffer): when
// any lane's column is nearly full the scanner hands the buffer over and carries on in the other one.  Why:
// with compaction inline in the scanning warp (round 1, and the generic kernel's EPI_TOPK path) a tile lasts as long
// as its SLOWEST quarter, and some quarter compacts (+2600 instructions) in most tiles, so the accumulator hand-off
// ran at scan + compaction per tile although the average work is scan + ~0.5 compaction; two scanning sets per
// quarter did not help (measured: one set 0.68 ms, two sets 0.70-0.83 ms on C2) because both stalled on the same
// hand-off.  Here the scanner's tile time is constant, the compaction runs beside it in the issue slots the scanner
// leaves idle (a lone warp issues ~0.3 instructions per cycle), survivors never leave registers (the inline version
// kept them in local memory across the call: 11 % of all stall samples), and a row is one candidate stream instead
// of two (32(1 + ln(n+/32)) accepted candidates instead of twice 32(1 + ln(n+/64))).
#pragma once
#include "gemm_sm100.cuh"

namespace freud {

template <int BN, int STAGES>
struct TopkSmem {
  static constexpr int kThreads = 384;  // 4 control warps + 4 scanners + 4 compactors
  static constexpr int kABytes = kBM * kBKBytes;
  static constexpr int kBBytes = BN * kBKBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kRing = STAGES * kStageBytes;
  static constexpr int kBufBytes = kNewSlots * kSlotStride;  // one candidate buffer of one quarter
  static constexpr int kCand = 4 * 2 * kBufBytes;             // [quarter][buffer]
  static constexpr int kCnt = 4 * 2 * 64 * 4;                 // [quarter][buffer][32 fill pointers + flags]
  static constexpr int kThr = kBM * 4;                        // published per-row thresholds
  static constexpr int kBiasS = 4 * 2 * BN * 4;               // [scanner warp][use parity][BN]
  static constexpr int kBars = (2 * STAGES + 4 + 16) * 8 + 16;
  static constexpr int kTotal = kRing + kCand + kCnt + kThr + kBiasS + kBars;
};

// STATS (diagnostic build, freud_topk_encode_stats): lane 0 of every scanner / compactor warp accumulates clock64
// cycles spent in its waits and counts its hand-overs / compactions into p.stats.
template <int BN, int STAGES, bool TF32, int CEV = 0, bool STATS = false>
__global__ void __launch_bounds__(384, 1)
sm100_topk_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                  const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB1,
                  const GemmParams p) {
  using L = TopkSmem<BN, STAGES>;
  constexpr int NBUF = 2;
  constexpr int kBKe = TF32 ? 32 : 64;  // elements per 128-byte k-block
  constexpr uint32_t kTmemCols = 512;
  static_assert(NBUF * BN <= 512 && BN == 256, "two 256-column accumulator buffers");

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* ring = smem;
  uint8_t* cand = smem + L::kRing;
  uint32_t* cnt_s = reinterpret_cast<uint32_t*>(smem + L::kRing + L::kCand);
  float* thr_s = reinterpret_cast<float*>(smem + L::kRing + L::kCand + L::kCnt);
  float* bias_s = reinterpret_cast<float*>(smem + L::kRing + L::kCand + L::kCnt + L::kThr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kRing + L::kCand + L::kCnt + L::kThr + L::kBiasS);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + NBUF;
  uint64_t* cfull_bar = bars + 2 * STAGES + 2 * NBUF;       // [quarter][buffer]: scanner -> compactor
  uint64_t* cempty_bar = bars + 2 * STAGES + 2 * NBUF + 8;  // [quarter][buffer]: compactor -> scanner
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * NBUF + 16);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_nt = (p.N + BN - 1) / BN;
  // this CTA's run of column tiles: a whole row block, or one column range of a row block of the split tail wave
  int mb = static_cast<int>(blockIdx.x);
  int nt_begin = 0, nt_end = num_nt, piece = 0;
  const bool is_piece = p.tail_split > 1 && static_cast<int>(blockIdx.x) >= p.full_count;
  if (is_piece) {
    const int tail = (p.M + kBM - 1) / kBM - p.full_count;
    const int t = static_cast<int>(blockIdx.x) - p.full_count;
    piece = t / tail;
    mb = p.full_count + t - piece * tail;
    nt_begin = static_cast<int>(static_cast<int64_t>(piece) * num_nt / p.tail_split);
    nt_end = static_cast<int>(static_cast<int64_t>(piece + 1) * num_nt / p.tail_split);
  }
  const int num_lt = nt_end - nt_begin;
  const int m0 = mb * kBM;
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
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 4);  // the four scanner warps
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&cfull_bar[i], 1);
      mbar_init(&cempty_bar[i], 1);
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  if (threadIdx.x < kBM) thr_s[threadIdx.x] = 0.f;  // published thresholds start at the ReLU floor
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int lt = 0; lt < num_lt; ++lt) {
        const int nt = nt_begin + lt;
        for (int vk = 0; vk < num_vk; ++vk) {
          const int pass = vk / num_kb;
          const int kb = vk - pass * num_kb;
          // pass order (3-pass split): hi*lo, lo*hi, then the dominant hi*hi term
          const CUtensorMap* ma = (p.passes == 1 || pass != 1) ? &mapA0 : &mapA1;
          const CUtensorMap* mbp = (p.passes == 1 || pass != 0) ? &mapB0 : &mapB1;
          mbar_wait_relaxed(&empty_bar[stage], phase ^ 1, 64);
          uint8_t* sa = ring + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
          tma_load_2d(sa, ma, &full_bar[stage], kb * kBKe, m0, kEvictNormal);
          tma_load_2d(sb, mbp, &full_bar[stage], kb * kBKe, nt * BN, kEvictLast);
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
      for (int lt = 0; lt < num_lt; ++lt) {
        const int buf = lt % NBUF;
        // the scanners arrive once they have read the buffer's previous tile out of TMEM (fresh buffers pass at once)
        mbar_wait_relaxed(&tempty_bar[buf], ((lt / NBUF) & 1) ^ 1, 32);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int vk = 0; vk < num_vk; ++vk) {
          mbar_wait_relaxed(&full_bar[stage], phase, 20);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {  // 4 x 32-byte UMMA_K steps per 128-byte k-block
            const uint64_t adesc = make_kmajor_sw128_desc(sa + k4 * 32);
            const uint64_t bdesc = make_kmajor_sw128_desc(sb + k4 * 32);
            const uint32_t acc = (vk > 0 || k4 > 0) ? 1u : 0u;  // the tile's first MMA overwrites the buffer
            if constexpr (TF32)
              mma_tf32_ss(d_tmem, adesc, bdesc, idesc, acc);
            else
              mma_f16_ss(d_tmem, adesc, bdesc, idesc, acc);
          }
          tc_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(&tfull_bar[buf]);
      }
    }
  } else if (warp_idx >= 4 && warp_idx < 8) {
    // ===================== scanner =====================
    const int q = warp_idx & 3;        // TMEM lane quarter
    const int stid = q * 32 + lane;    // token row within the row block
    const int row = m0 + stid;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // bias rows staged per warp, one tile ahead (see sm100_gemm_kernel)
    float* bias_w = bias_s + q * (2 * BN);
    float nb[8];
    auto load_bias = [&](int lt) {
      const int nt = nt_begin + lt;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int gc = nt * BN + lane * 8 + j;
        nb[j] = gc < p.N ? (p.bias ? __ldg(p.bias + gc) : 0.f) : -INFINITY;  // out-of-range columns can never win
      }
    };
    auto store_bias = [&](int slot) {
      float4* dst = reinterpret_cast<float4*>(bias_w + slot * BN + lane * 8);
      dst[0] = make_float4(nb[0], nb[1], nb[2], nb[3]);
      dst[1] = make_float4(nb[4], nb[5], nb[6], nb[7]);
    };
    const uint32_t cand_q = smem_u32(cand) + q * 2 * L::kBufBytes + lane * 8;
    uint32_t* cnt_q = cnt_s + q * 2 * 64;
    int cur = 0;
    // parity the scanner waits for before (re)using a buffer.  Buffer 1 is first waited for before its FIRST use
    // (a fresh barrier passes a wait on parity 1); buffer 0 starts in use, so its first wait -- at the second
    // hand-over -- must see the compactor's first release (parity 0).  Getting this wrong lets the scanner lap the
    // compactor: two arrivals on a count-1 barrier flip its phase twice and the waiter never sees either.
    uint32_t ephase[2] = {0u, 1u};
    long long st_t0 = 0, st_tfull = 0, st_cempty = 0, st_ho = 0;
    if constexpr (STATS) st_t0 = clock64();
    uint32_t base = cand_q;
    uint32_t ptr = base;
    uint32_t ptr_limit = base + (kNewSlots - kCheck) * kSlotStride;
    float thresh = 0.f;
    int bslot = 0;
    // hand the current candidate buffer to the compactor and continue in the other one
    auto hand_over = [&](uint32_t last) {
      cnt_q[cur * 64 + lane] = ptr;
      if (lane == 0) cnt_q[cur * 64 + 32] = last;
      __syncwarp();
      if (lane == 0) mbar_arrive(&cfull_bar[q * 2 + cur]);
      cur ^= 1;
      long long c0 = 0;
      if constexpr (STATS) c0 = clock64();
      mbar_wait(&cempty_bar[q * 2 + cur], ephase[cur]);  // the compactor frees a buffer as soon as it has loaded it
      if constexpr (STATS) {
        st_cempty += clock64() - c0;
        ++st_ho;
      }
      ephase[cur] ^= 1u;
      base = cand_q + cur * L::kBufBytes;
      ptr = base;
      ptr_limit = base + (kNewSlots - kCheck) * kSlotStride;
    };
    if (num_lt > 0) {
      load_bias(0);
      store_bias(bslot);
      __syncwarp();
    }
    for (int lt = 0; lt < num_lt; ++lt) {
      const int nt = nt_begin + lt;
      const int buf = lt % NBUF;
      // lower bounds of the row's final 32nd largest value: the compactor's latest (this row's own stream: later
      // candidates have higher indices, so ties lose and the comparison is strict) and, in a split tail wave, the one
      // other column pieces of the row block have published -- a value EQUAL to a foreign bound may still win on the
      // index tie-break, so that one enters one ulp low
      thresh = fmaxf(thresh, thr_s[stid]);
      if (is_piece && row < p.M) {
        const float foreign = __ldcg(p.part_thr + row);
        if (foreign > 0.f) thresh = fmaxf(thresh, __uint_as_float(__float_as_uint(foreign) - 1u));
      }
      if (lt + 1 < num_lt) load_bias(lt + 1);  // lands in registers while this tile is scanned
      const uint32_t bs_addr = smem_u32(bias_w + bslot * BN);
      long long c1 = 0;
      if constexpr (STATS) c1 = clock64();
      mbar_wait_relaxed(&tfull_bar[buf], (lt / NBUF) & 1, 32);
      if constexpr (STATS) st_tfull += clock64() - c1;
      tc_fence_after();
      const uint32_t t_addr = lane_taddr + buf * BN;
      uint32_t r[2][kChunk];
      tmem_ld_32x32b_x16(t_addr, r[0]);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 2 * kChunk) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int cc = c0 + h * kChunk;
          tmem_ld_wait();
          if (cc + kChunk < BN) {
            tmem_ld_32x32b_x16(t_addr + cc + kChunk, r[h ^ 1]);  // prefetch the next chunk
          } else {
            // the whole tile now sits in registers: hand the accumulator back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
          }
          // pre-activation = fl(accumulator + bias), the reference's order of operations (nn.Linear: x @ W.T, then + b)
          float v[kChunk];
#pragma unroll
          for (int j = 0; j < kChunk; j += 4) {
            const float4 b4 = lds128(bs_addr + (cc + j) * 4);
            v[j] = __uint_as_float(r[h][j]) + b4.x;
            v[j + 1] = __uint_as_float(r[h][j + 1]) + b4.y;
            v[j + 2] = __uint_as_float(r[h][j + 2]) + b4.z;
            v[j + 3] = __uint_as_float(r[h][j + 3]) + b4.w;
          }
          thresh = fmaxf(thresh, thr_s[stid]);  // whatever the compactor has published meanwhile
          const uint32_t nidx0 = ~static_cast<uint32_t>(nt * BN + cc);  // ~(col) == nidx0 - j
#pragma unroll
          for (int g = 0; g < kChunk; g += kCheck) {
            // Append the values above the threshold: value j of the group goes to slot ptr + (number of accepted
            // values before it).  The slot addresses are formed as a shallow tree of 3-input adds over the eight
            // 0 / stride increments (depth 4 per group) instead of a bump-after-every-value chain (depth 16): a lone
            // scanning warp is latency-bound, and the chain was its critical path.
            uint32_t inc[kCheck], addr[kCheck + 1];
            bool take[kCheck];
            addr[0] = ptr;
#pragma unroll
            for (int j = 0; j < kCheck; ++j) {
              take[j] = v[g + j] > thresh;
              // (opaque to the optimiser on purpose: written as a C++ select, the compiler folds the adds below back
              //  into the serial  next = take ? addr + stride : addr  chain)
              asm("{\n\t.reg .pred p;\n\t"
                  "setp.gt.f32 p, %1, %2;\n\t"
                  "selp.b32 %0, %3, 0, p;\n\t}"
                  : "=r"(inc[j])
                  : "f"(v[g + j]), "f"(thresh), "n"(kSlotStride));
            }
            static_assert(kCheck == 8, "the address tree below is written for groups of eight");
            addr[1] = addr[0] + inc[0];
            addr[2] = addr[0] + inc[0] + inc[1];
            addr[3] = addr[2] + inc[2];
            addr[4] = addr[2] + inc[2] + inc[3];
            addr[5] = addr[4] + inc[4];
            addr[6] = addr[4] + inc[4] + inc[5];
            addr[7] = addr[6] + inc[6];
            addr[8] = addr[6] + inc[6] + inc[7];
#pragma unroll
            for (int j = 0; j < kCheck; ++j) {
              if (take[j])  // a single predicated st.shared.v2 (checked in the SASS: no branch)
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr[j]), "r"(nidx0 - (g + j)),
                             "r"(__float_as_uint(v[g + j]))
                             : "memory");
            }
            ptr = addr[8];
            // the next kCheck columns could overflow some lane's column.  (A warp-uniform fill bound that defers the
            // check -- one REDUX when the bound reaches the limit instead of a vote per group -- measured 4 % slower.)
            if (__any_sync(0xffffffffu, ptr > ptr_limit)) hand_over(0u);
          }
        }
      }
      if (lt + 1 < num_lt) {  // next tile's bias row goes into the other slot (last read by this warp one tile ago)
        store_bias(bslot ^ 1);
        __syncwarp();
        bslot ^= 1;
      }
    }
    {  // end of the row block: the last (possibly empty) buffer, flagged
      cnt_q[cur * 64 + lane] = ptr;
      if (lane == 0) cnt_q[cur * 64 + 32] = 1u;
      __syncwarp();
      if (lane == 0) mbar_arrive(&cfull_bar[q * 2 + cur]);
    }
    if constexpr (STATS) {
      if (lane == 0 && p.stats != nullptr) {
        atomicAdd(p.stats + 0, static_cast<unsigned long long>(clock64() - st_t0));
        atomicAdd(p.stats + 1, static_cast<unsigned long long>(st_tfull));
        atomicAdd(p.stats + 2, static_cast<unsigned long long>(st_cempty));
        atomicAdd(p.stats + 3, static_cast<unsigned long long>(st_ho));
        atomicAdd(p.stats + 8, static_cast<unsigned long long>(num_lt));
      }
    }
  } else if (warp_idx >= 8) {
    // ===================== compactor =====================
    const int q = warp_idx & 3;
    const int stid = q * 32 + lane;
    const int row = m0 + stid;
    const uint32_t cand_q = smem_u32(cand) + q * 2 * L::kBufBytes + lane * 8;
    const uint32_t* cnt_q = cnt_s + q * 2 * 64;
    uint64_t surv[kTopK];
#pragma unroll
    for (int s = 0; s < kTopK; ++s) surv[s] = 0ull;
    int cur = 0;
    uint32_t fphase[2] = {0u, 0u};
    long long ct_t0 = 0, ct_wait = 0, ct_n = 0;
    if constexpr (STATS) ct_t0 = clock64();
    for (;;) {
      long long c2 = 0;
      if constexpr (STATS) c2 = clock64();
      mbar_wait_suspended(&cfull_bar[q * 2 + cur], fphase[cur]);
      if constexpr (STATS) ct_wait += clock64() - c2;
      fphase[cur] ^= 1u;
      const uint32_t my_ptr = cnt_q[cur * 64 + lane];
      const uint32_t last = cnt_q[cur * 64 + 32];
      const uint32_t my_base = cand_q + cur * L::kBufBytes;
      uint64_t fresh[kNewSlots];
#pragma unroll
      for (int j = 0; j < kNewSlots; ++j) {
        const uint32_t addr = my_base + j * kSlotStride;
        fresh[j] = addr < my_ptr ? lds64(addr) : 0ull;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&cempty_bar[q * 2 + cur]);  // the scanner may refill it while we sort
      cur ^= 1;
      if (__any_sync(0xffffffffu, my_ptr != my_base)) {
        if constexpr (STATS) ++ct_n;
        bitonic_sort_desc<kNewSlots, CEV>(fresh);
        // max(descending, reversed descending) = the 32 largest of the union, as a bitonic sequence
#pragma unroll
        for (int i = 0; i < kTopK; ++i) {
          uint64_t y = fresh[kNewSlots - 1 - i];
          cmp_exchange<CEV>(surv[i], y);  // keeps the larger; the smaller is dropped
        }
        bitonic_merge_desc<kTopK, CEV>(surv);
        const float t = __uint_as_float(static_cast<uint32_t>(surv[kTopK - 1] >> 32));
        thr_s[stid] = t;
        if (is_piece && row < p.M && t > 0.f)  // non-negative floats order like their bit patterns
          atomicMax(reinterpret_cast<int*>(p.part_thr + row), __float_as_int(t));
      }
      if (last) break;
    }
    if constexpr (STATS) {
      if (lane == 0 && p.stats != nullptr) {
        atomicAdd(p.stats + 4, static_cast<unsigned long long>(clock64() - ct_t0));
        atomicAdd(p.stats + 5, static_cast<unsigned long long>(ct_wait));
        atomicAdd(p.stats + 6, static_cast<unsigned long long>(ct_n));
        atomicAdd(p.stats + 7, 1ull);
      }
    }
    // emit this thread's row.  Short rows (fewer than 32 positive pre-activations) are completed with zeros at the
    // lowest indices not already chosen: the oracle's (value desc, index asc) order for the all-zero tail after ReLU.
    if (row < p.M) {
      const bool whole = !is_piece;
      const uint32_t col_base = static_cast<uint32_t>(nt_begin) * BN;
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
          *reinterpret_cast<int4*>(oi + s) = make_int4(static_cast<int>(~static_cast<uint32_t>(surv[s])),
                                                       static_cast<int>(~static_cast<uint32_t>(surv[s + 1])),
                                                       static_cast<int>(~static_cast<uint32_t>(surv[s + 2])),
                                                       static_cast<int>(~static_cast<uint32_t>(surv[s + 3])));
        }
        if (whole && p.hist != nullptr) {
#pragma unroll
          for (int s = 0; s < kTopK; ++s) atomicAdd(p.hist + ~static_cast<uint32_t>(surv[s]), 1);
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
          if (whole && p.hist != nullptr) atomicAdd(p.hist + idx, 1);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace freud
