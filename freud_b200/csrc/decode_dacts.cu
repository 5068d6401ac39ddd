// Fused sparse decode + activation gradient of the TopK SAE (bf16 mode, k = 32):
//   sae_out[t] = sum_j a[t,j] W_dec[i[t,j]] + b_dec          (topkautoencoder.py:15-18,87-91 eager_decode / decode)
//   e[t]       = sae_out[t] - x[t]                            (:101)      -> bf16 residual, SSE, column sums
//   dacts[t,j] = < bf16(e[t]), W_dec[i[t,j]] >                (autograd of :17-18 w.r.t. the selected activations)
// freud_topk_decode and freud_topk_dacts gather the same 32 decoder rows of a token from L2 back to back (2 x k*d*2
// bytes per token, the dominant traffic of both).  Here a token's rows are fetched ONCE, by one bulk copy
// (cp.async.bulk, the TMA engine) per row issued by the lane that owns the row, straight into shared memory, and both
// passes read them from there.  One warp per token, TOK tokens in flight per CTA.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/freud_b200.h"
#include "device_utils.cuh"
#include "host_common.h"
#include "ptx.cuh"

namespace freud {

__device__ __forceinline__ void bulk_row_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                             uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// (c0, c1) += (a0, a1) * (the two bf16 values packed in w), as ONE packed fp32 FMA (FFMA2, sm_100): each half is an
// IEEE fma, bit-identical to two scalar fmaf.  The kernel is bound by instruction issue (a shift / mask + an FMA per
// bf16 value), not by the L2 gather, so halving the FMA count matters.
__device__ __forceinline__ void fma2(float& c0, float& c1, float a0, float a1, uint32_t w) {
  const float2 r = __ffma2_rn(make_float2(a0, a1),
                              make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)),
                              make_float2(c0, c1));
  c0 = r.x;
  c1 = r.y;
}

// PARTS: a token's rows pass through its warp's slot in PARTS column ranges of D / PARTS columns (the reconstruction of a
// column needs that column of all 32 rows and nothing else, and the activation gradients are sums of per-range dots),
// so a slot is 32 x (D / PARTS) bf16 and TOK = one token per warp stay in flight per CTA whatever D is.
template <int D, int PARTS, int TOK>
struct DecodeDactsCfg {
  static constexpr int kThreads = 32 * TOK;
  static constexpr int kV = 4;                         // columns per 8-byte slice
  static constexpr int kDP = D / PARTS;                // columns per part
  static constexpr int kCH = kDP / (32 * kV);          // slices per lane and part
  static_assert(D % PARTS == 0 && kDP % (32 * kV) == 0, "a part must split into 8-byte slices over the warp");
  static constexpr size_t kPartBytes = static_cast<size_t>(kDP) * 2;
  static constexpr size_t kSlot = 32 * kPartBytes;
  static constexpr size_t kRows = static_cast<size_t>(TOK) * kSlot;
  static constexpr size_t kSmem = kRows + TOK * 8 + D * 4 + 32 * 8;
};

template <int D, int PARTS, int TOK, typename XT>
__global__ void __launch_bounds__(32 * TOK, 1)
decode_dacts_kernel(const float* __restrict__ top_vals, const int32_t* __restrict__ top_idx,
                    const __nv_bfloat16* __restrict__ W, const float* __restrict__ b_dec,
                    const XT* __restrict__ target, float* __restrict__ sae_out,
                    __nv_bfloat16* __restrict__ resid, double* __restrict__ sse, float* __restrict__ colsum,
                    float* __restrict__ dacts, int64_t N) {
  using Cfg = DecodeDactsCfg<D, PARTS, TOK>;
  constexpr int V = Cfg::kV, CH = Cfg::kCH, DP = Cfg::kDP;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kRows);
  float* colsum_s = reinterpret_cast<float*>(smem + Cfg::kRows + TOK * 8);       // [D]
  double* scratch = reinterpret_cast<double*>(colsum_s + D);                     // [32]

  const int lane = threadIdx.x & 31;
  const int p = threadIdx.x >> 5;  // this warp's slot
  uint8_t* my_rows = smem + static_cast<size_t>(p) * Cfg::kSlot;
  const uint8_t* rbase = my_rows + lane * V * 2;  // slice i of slot row j: rbase + j * kPartBytes + i * 32 * V * 2
  uint64_t* bar = bars + p;
  const uint64_t keep = l2_keep_policy();

  if (lane == 0) mbar_init(bar, 1);
  for (int c = threadIdx.x; c < D; c += Cfg::kThreads) colsum_s[c] = 0.f;
  fence_barrier_init();
  __syncthreads();

  double sq = 0.0;
  float csum[PARTS][CH][V];
#pragma unroll
  for (int h = 0; h < PARTS; ++h)
#pragma unroll
    for (int i = 0; i < CH; ++i)
#pragma unroll
      for (int e = 0; e < V; ++e) csum[h][i][e] = 0.f;

  uint32_t phase = 0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * TOK;
  int64_t t = static_cast<int64_t>(blockIdx.x) * TOK + p;
  // the selection of the NEXT token is fetched while the current one is processed: the bulk copies of a token can
  // then be issued without first waiting for an index load
  float nxt_a = 0.f;
  int nxt_i = -1;
  if (t < N) {
    nxt_a = __ldg(top_vals + t * 32 + lane);
    nxt_i = __ldg(top_idx + t * 32 + lane);
  }
  for (; t < N; t += stride) {
    const float my_a = nxt_a;
    const int my_i = nxt_i;
    if (t + stride < N) {
      nxt_a = __ldg(top_vals + (t + stride) * 32 + lane);
      nxt_i = __ldg(top_idx + (t + stride) * 32 + lane);
    }
    const unsigned live = __ballot_sync(0xffffffffu, my_i >= 0);  // -1: entry owned by another dictionary shard
    const uint32_t tx = static_cast<uint32_t>(__popc(live)) * static_cast<uint32_t>(Cfg::kPartBytes);
    const __nv_bfloat16* my_row = W + static_cast<int64_t>(my_i < 0 ? 0 : my_i) * D;
    float dsum = 0.f;
#pragma unroll
    for (int h = 0; h < PARTS; ++h) {
      if (lane == 0) mbar_arrive_expect_tx(bar, tx);
      __syncwarp();
      if (my_i >= 0)
        bulk_row_g2s(my_rows + lane * Cfg::kPartBytes, my_row + h * DP, (uint32_t)Cfg::kPartBytes, bar, keep);
      const int col0 = h * DP + lane * V;  // slice i covers columns col0 + i*32*V .. +V
      float xv[CH][V], acc[CH][V];
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const float4 v = load4_stream(target + t * D + col0 + i * 32 * V);
        xv[i][0] = v.x; xv[i][1] = v.y; xv[i][2] = v.z; xv[i][3] = v.w;
        const float4 b = load4(b_dec + col0 + i * 32 * V);
        acc[i][0] = b.x; acc[i][1] = b.y; acc[i][2] = b.z; acc[i][3] = b.w;
      }
      mbar_wait(bar, phase);
      phase ^= 1;

      // ---- decode: rows in slot order, fp32 FMA chain per column (the order of freud_topk_decode)
      if (live == 0xffffffffu) {
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
          const float a = __shfl_sync(0xffffffffu, my_a, j);
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            const uint2 raw = *reinterpret_cast<const uint2*>(rbase + j * Cfg::kPartBytes + i * 32 * V * 2);
            fma2(acc[i][0], acc[i][1], a, a, raw.x);
            fma2(acc[i][2], acc[i][3], a, a, raw.y);
          }
        }
      } else {
        for (int j = 0; j < 32; ++j) {
          if (!((live >> j) & 1u)) continue;
          const float a = __shfl_sync(0xffffffffu, my_a, j);
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            const uint2 raw = *reinterpret_cast<const uint2*>(rbase + j * Cfg::kPartBytes + i * 32 * V * 2);
            fma2(acc[i][0], acc[i][1], a, a, raw.x);
            fma2(acc[i][2], acc[i][3], a, a, raw.y);
          }
        }
      }

      // ---- reconstruction out, residual (bf16, what the backward gathers), SSE and column sums on the fp32 residual
      float eb[CH][V];
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int64_t o = t * D + col0 + i * 32 * V;
        store4_stream(sae_out + o, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        float err[V];
#pragma unroll
        for (int e = 0; e < V; ++e) err[e] = acc[i][e] - xv[i][e];
        sq += (double)(err[0] * err[0] + err[1] * err[1]) + (double)(err[2] * err[2] + err[3] * err[3]);
#pragma unroll
        for (int e = 0; e < V; ++e) csum[h][i][e] += err[e];
        const __nv_bfloat162 lo = __floats2bfloat162_rn(err[0], err[1]);
        const __nv_bfloat162 hi = __floats2bfloat162_rn(err[2], err[3]);
        uint2 raw;
        raw.x = *reinterpret_cast<const uint32_t*>(&lo);
        raw.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(resid + o) = raw;
        eb[i][0] = __uint_as_float(raw.x << 16);
        eb[i][1] = __uint_as_float(raw.x & 0xffff0000u);
        eb[i][2] = __uint_as_float(raw.y << 16);
        eb[i][3] = __uint_as_float(raw.y & 0xffff0000u);
      }

      // ---- dacts: the same rows against the bf16 residual; lane j ends up with row j's dot over this part's columns
      float part[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const uint2 raw = *reinterpret_cast<const uint2*>(rbase + j * Cfg::kPartBytes + i * 32 * V * 2);
          fma2(s0, s1, eb[i][0], eb[i][1], raw.x);
          fma2(s0, s1, eb[i][2], eb[i][3], raw.y);
        }
        part[j] = ((live >> j) & 1u) ? s0 + s1 : 0.f;  // absent rows hold stale bytes
      }
      dsum += warp_transpose_reduce32(part, lane);
      __syncwarp();  // every lane is done with the slot before its next rows are requested
    }
    dacts[t * 32 + lane] = dsum;
  }

#pragma unroll
  for (int h = 0; h < PARTS; ++h)
#pragma unroll
    for (int i = 0; i < CH; ++i)
#pragma unroll
      for (int e = 0; e < V; ++e) atomicAdd(colsum_s + h * DP + lane * V + i * 32 * V + e, csum[h][i][e]);
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += Cfg::kThreads) atomicAdd(colsum + c, colsum_s[c]);
  const double tot = block_sum(sq, scratch);
  if (threadIdx.x == 0) atomicAdd(sse, tot);
}

template <int D, int PARTS, int TOK, typename XT>
static int launch_decode_dacts_t(const float* tv, const int32_t* ti, const void* W, const float* b_dec,
                                 const void* target, float* sae_out, void* resid, double* sse, float* colsum,
                                 float* dacts, int64_t N, cudaStream_t s) {
  using Cfg = DecodeDactsCfg<D, PARTS, TOK>;
  auto kern = decode_dacts_kernel<D, PARTS, TOK, XT>;
  static bool configured = false;
  if (!configured) {
    FREUD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
    configured = true;
  }
  int64_t grid = (N + TOK - 1) / TOK;
  if (grid > sm_count()) grid = sm_count();
  kern<<<(int)grid, Cfg::kThreads, Cfg::kSmem, s>>>(tv, ti, static_cast<const __nv_bfloat16*>(W), b_dec,
                                                      static_cast<const XT*>(target), sae_out,
                                                      static_cast<__nv_bfloat16*>(resid), sse, colsum, dacts, N);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int D, int PARTS, int TOK>
static int launch_decode_dacts(const float* tv, const int32_t* ti, const void* W, const float* b_dec,
                               const void* target, int target_dtype, float* sae_out, void* resid, double* sse,
                               float* colsum, float* dacts, int64_t N, cudaStream_t s) {
  if (target_dtype == 1)
    return launch_decode_dacts_t<D, PARTS, TOK, __half>(tv, ti, W, b_dec, target, sae_out, resid, sse, colsum, dacts, N, s);
  if (target_dtype == 2)
    return launch_decode_dacts_t<D, PARTS, TOK, __nv_bfloat16>(tv, ti, W, b_dec, target, sae_out, resid, sse, colsum, dacts, N, s);
  return launch_decode_dacts_t<D, PARTS, TOK, float>(tv, ti, W, b_dec, target, sae_out, resid, sse, colsum, dacts, N, s);
}

}  // namespace freud

using namespace freud;

extern "C" int freud_topk_decode_dacts_supported(int64_t d, int64_t k) {
  return k == 32 && (d == 384 || d == 512 || d == 768 || d == 1024 || d == 1280);
}

extern "C" int freud_topk_decode_dacts(const float* top_vals, const int32_t* top_idx, const void* W_dec_bf16,
                                       const float* b_dec, const void* target, int target_dtype, float* sae_out,
                                       void* resid_bf16, double* sse, float* colsum, float* dacts, int64_t N, int64_t d,
                                       int64_t k, void* stream) {
  FREUD_REQUIRE(N > 0 && freud_topk_decode_dacts_supported(d, k),
                "fused decode + dacts needs k == 32 and d in {384, 512, 768, 1024, 1280}");
  FREUD_REQUIRE(target && sae_out && resid_bf16 && sse && colsum && dacts, "all outputs are required");
  FREUD_REQUIRE(target_dtype >= 0 && target_dtype <= 2, "target_dtype: 0 fp32, 1 fp16, 2 bf16");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (d) {
    case 384: return launch_decode_dacts<384, 1, 8>(top_vals, top_idx, W_dec_bf16, b_dec, target, target_dtype, sae_out, resid_bf16, sse, colsum, dacts, N, s);
    case 512: return launch_decode_dacts<512, 1, 6>(top_vals, top_idx, W_dec_bf16, b_dec, target, target_dtype, sae_out, resid_bf16, sse, colsum, dacts, N, s);
    case 768: return launch_decode_dacts<768, 2, 8>(top_vals, top_idx, W_dec_bf16, b_dec, target, target_dtype, sae_out, resid_bf16, sse, colsum, dacts, N, s);
    case 1024: return launch_decode_dacts<1024, 2, 6>(top_vals, top_idx, W_dec_bf16, b_dec, target, target_dtype, sae_out, resid_bf16, sse, colsum, dacts, N, s);
    default: return launch_decode_dacts<1280, 2, 5>(top_vals, top_idx, W_dec_bf16, b_dec, target, target_dtype, sae_out, resid_bf16, sse, colsum, dacts, N, s);
  }
}
