// Bandwidth-bound kernels of the TopK SAE step: operand preparation + total variance, sparse decode,
// gather-dot activation gradients, feature-major (CSC) index, row-sparse weight gradients, bias gradients,
// loss scalars, dead-latent counters and the decoder norm helpers.  All are gather / stream kernels bounded by
// L2 / HBM bandwidth; none is reshaped into a GEMM.
#include <type_traits>

#include <algorithm>
#include <cstdlib>
#include <string>
#include "device_utils.cuh"
#include "host_common.h"
#include "../../include/freud_b200.h"

namespace freud {

// ------------------------------------------------------------------------------------------------ prep_x
// One thread per (t, 4 channels); loops over the batch axis so the per-(t,c) mean of topkautoencoder.py:104 is a
// private register reduction.  Shifted sums (shift = x[0,t,c]) keep sum((x-mean)^2) = s2 - s1^2/B well conditioned.
// XT: storage type of x (fp32, or fp16 / bf16 as collected activation stores hold it: widened here, exactly, instead of
// in a separate pass over the batch).
template <int MODE, typename XT>  // MODE 0: bf16 out, 1: tf32 hi/lo out
__global__ void __launch_bounds__(256) prep_x_kernel(const XT* __restrict__ x, const float* __restrict__ b_dec,
                                                     void* __restrict__ out_hi, void* __restrict__ out_lo,
                                                     double* __restrict__ tv, float* __restrict__ colmean, int B,
                                                     int64_t T, int d) {
  __shared__ double scratch[32];
  const int d4 = d >> 2;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  double part = 0.0;
  if (idx < T * d4) {
    const int64_t t = idx / d4;
    const int c = static_cast<int>(idx - t * d4) * 4;
    const float4 bd = load4(b_dec + c);
    const int64_t stride = T * d;
    const XT* px = x + t * d + c;
    const float4 sh = load4(px);
    float4 s1 = make_float4(0, 0, 0, 0), s2 = make_float4(0, 0, 0, 0);
#pragma unroll 8
    for (int b = 0; b < B; ++b) {
      const float4 v = load4(px + b * stride);
      const float4 xc = make_float4(v.x - bd.x, v.y - bd.y, v.z - bd.z, v.w - bd.w);
      const int64_t o = b * stride + t * d + c;
      if (MODE == 0) {
        store4(reinterpret_cast<__nv_bfloat16*>(out_hi) + o, xc);
      } else {
        const float4 hi = make_float4(to_tf32(xc.x), to_tf32(xc.y), to_tf32(xc.z), to_tf32(xc.w));
        const float4 lo = make_float4(to_tf32(xc.x - hi.x), to_tf32(xc.y - hi.y), to_tf32(xc.z - hi.z),
                                      to_tf32(xc.w - hi.w));
        store4(reinterpret_cast<float*>(out_hi) + o, hi);
        store4(reinterpret_cast<float*>(out_lo) + o, lo);
      }
      const float4 dv = make_float4(v.x - sh.x, v.y - sh.y, v.z - sh.z, v.w - sh.w);
      s1.x += dv.x; s1.y += dv.y; s1.z += dv.z; s1.w += dv.w;
      s2.x += dv.x * dv.x; s2.y += dv.y * dv.y; s2.z += dv.z * dv.z; s2.w += dv.w * dv.w;
    }
    const double inv = 1.0 / B;
    if (colmean) {
      const float fi = 1.f / B;  // mean over the batch axis = shift + s1 / B
      store4(colmean + t * d + c, make_float4(sh.x + s1.x * fi, sh.y + s1.y * fi, sh.z + s1.z * fi, sh.w + s1.w * fi));
    }
    part = (double)s2.x - (double)s1.x * s1.x * inv + (double)s2.y - (double)s1.y * s1.y * inv +
           (double)s2.z - (double)s1.z * s1.z * inv + (double)s2.w - (double)s1.w * s1.w * inv;
  }
  const double tot = block_sum(part, scratch);
  if (threadIdx.x == 0 && tot != 0.0) atomicAdd(tv, tot);
}

template <int MODE>
__global__ void __launch_bounds__(256) split_operand_kernel(const float* __restrict__ w, void* __restrict__ hi,
                                                            void* __restrict__ lo, int64_t n4) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = load4(w + i * 4);
    if (MODE == 0) {
      store4(reinterpret_cast<__nv_bfloat16*>(hi) + i * 4, v);
    } else {
      const float4 h = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
      const float4 l = make_float4(to_tf32(v.x - h.x), to_tf32(v.y - h.y), to_tf32(v.z - h.z), to_tf32(v.w - h.w));
      store4(reinterpret_cast<float*>(hi) + i * 4, h);
      store4(reinterpret_cast<float*>(lo) + i * 4, l);
    }
  }
}

// 16-byte slice of a gathered row: kVec elements, widened to fp32 after the loads of a batch have been issued.
template <typename T> struct RowVec;
template <> struct RowVec<__nv_bfloat16> {
  static constexpr int kVec = 8;
  __device__ static __forceinline__ void unpack(const uint4& raw, float (&v)[8]) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    unpack(__ldg(reinterpret_cast<const uint4*>(p)), v);
  }
  __device__ static __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 raw;
    uint32_t* w = reinterpret_cast<uint32_t*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = raw;
  }
};
template <> struct RowVec<float> {
  static constexpr int kVec = 4;
  __device__ static __forceinline__ void unpack(const uint4& raw, float (&v)[4]) {
    v[0] = __uint_as_float(raw.x); v[1] = __uint_as_float(raw.y);
    v[2] = __uint_as_float(raw.z); v[3] = __uint_as_float(raw.w);
  }
  __device__ static __forceinline__ void load(const float* p, float (&v)[4]) {
    unpack(__ldg(reinterpret_cast<const uint4*>(p)), v);
  }
  __device__ static __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

// Row slice of VB bytes per lane (16, or 8 when the row is not a multiple of 32 x 16 bytes but is one of 32 x 8:
// d = 384 in bf16 keeps all lanes busy with three 8-byte slices instead of one and a half 16-byte ones).
template <typename T, int VB> struct Slice {
  static constexpr int kVec = VB / static_cast<int>(sizeof(T));
  using Raw = typename std::conditional<VB == 16, uint4, uint2>::type;
  __device__ static __forceinline__ Raw load_keep(const T* p, uint64_t pol) {
    if constexpr (VB == 16) {
      return ldg_keep(p, pol);
    } else {
      uint2 r;
      asm("ld.global.nc.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(pol));
      return r;
    }
  }
  __device__ static __forceinline__ void unpack(const Raw& raw, float (&v)[kVec]) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&raw);
    if constexpr (sizeof(T) == 2) {
#pragma unroll
      for (int i = 0; i < kVec / 2; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
      }
    } else {
#pragma unroll
      for (int i = 0; i < kVec; ++i) v[i] = __uint_as_float(w[i]);
    }
  }
};
template <typename T, int D> struct SliceFor {
  static constexpr int kBytes = (D * static_cast<int>(sizeof(T))) % 512 == 0 ? 16
                                : ((D * static_cast<int>(sizeof(T))) % 256 == 0 ? 8 : 16);
  using type = Slice<T, kBytes>;
};

// Pack the entries a shard owns (index >= 0) to the front of the warp's 32 slots, order preserved; returns the count.
__device__ __forceinline__ int compact_valid(int& my_i, float& my_a, int lane) {
  const unsigned m = __ballot_sync(0xffffffffu, my_i >= 0);
  const int cnt = __popc(m);
  if (m != 0xffffffffu) {  // warp-uniform; never taken without feature sharding
    const int src = static_cast<int>(__fns(m, 0, lane + 1)) & 31;
    const int ci = __shfl_sync(0xffffffffu, my_i, src);
    const float ca = __shfl_sync(0xffffffffu, my_a, src);
    my_i = lane < cnt ? ci : -1;
    my_a = lane < cnt ? ca : 0.f;
  }
  return cnt;
}

// ------------------------------------------------------------------------------------- decode, fixed row width
// Same contract as decode_kernel for the activation widths the Whisper family has (D known at compile time):
// lane l owns the 16-byte slices {(32 i + l) * V}, and the k gathered rows are fetched R at a time with every load
// of a batch issued before the first FMA (no branch inside a batch: absent entries gather row 0 with a = 0).
template <typename WT, typename RT, int D>
__global__ void __launch_bounds__(256) decode_fixed_kernel(const float* __restrict__ top_vals,
                                                           const int32_t* __restrict__ top_idx,
                                                           const WT* __restrict__ W, const float* __restrict__ b_dec,
                                                           const float* __restrict__ target,
                                                           float* __restrict__ sae_out, RT* __restrict__ resid,
                                                           double* __restrict__ sse, float* __restrict__ colsum,
                                                           int64_t N, int k) {
  using SL = typename SliceFor<WT, D>::type;
  constexpr int V = SL::kVec;
  constexpr int CH = (D + 32 * V - 1) / (32 * V);
  constexpr int R = CH <= 3 ? 8 : (CH <= 6 ? 4 : 2);  // rows in flight: <= 24 16-byte loads per lane
  __shared__ double scratch[32];
  __shared__ float colsum_s[D];
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const uint64_t keep = l2_keep_policy();
  if (colsum) {
    for (int c = threadIdx.x; c < D; c += blockDim.x) colsum_s[c] = 0.f;
    __syncthreads();
  }
  double sq = 0.0;
  float csum[CH][V];
#pragma unroll
  for (int i = 0; i < CH; ++i)
#pragma unroll
    for (int e = 0; e < V; ++e) csum[i][e] = 0.f;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); t < N;
       t += static_cast<int64_t>(gridDim.x) * warps_per_block) {
    float acc[CH][V];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = (i * 32 + lane) * V;
#pragma unroll
      for (int e = 0; e < V; e += 4) {
        const float4 b = c < D ? load4(b_dec + c + e) : make_float4(0, 0, 0, 0);
        acc[i][e] = b.x; acc[i][e + 1] = b.y; acc[i][e + 2] = b.z; acc[i][e + 3] = b.w;
      }
    }
    for (int j0 = 0; j0 < k; j0 += 32) {
      const int jn = min(32, k - j0);
      float my_a = lane < jn ? __ldg(top_vals + t * k + j0 + lane) : 0.f;
      int my_i = lane < jn ? __ldg(top_idx + t * k + j0 + lane) : -1;
      const int cnt = compact_valid(my_i, my_a, lane);
      for (int jb = 0; jb < cnt; jb += R) {
        typename SL::Raw raw[R][CH];
        float a[R];
#pragma unroll
        for (int u = 0; u < R; ++u) {
          const int fi = __shfl_sync(0xffffffffu, my_i, (jb + u) & 31);
          const float av = __shfl_sync(0xffffffffu, my_a, (jb + u) & 31);
          const bool ok = jb + u < cnt;
          a[u] = ok ? av : 0.f;
          const WT* wr = W + static_cast<int64_t>(ok ? fi : 0) * D;
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            const int c = (i * 32 + lane) * V;
            if (c < D) raw[u][i] = SL::load_keep(wr + c, keep);
          }
        }
#pragma unroll
        for (int u = 0; u < R; ++u) {
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            const int c = (i * 32 + lane) * V;
            if (c < D) {
              float w[V];
              SL::unpack(raw[u][i], w);
#pragma unroll
              for (int e = 0; e < V; ++e) acc[i][e] = fmaf(a[u], w[e], acc[i][e]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = (i * 32 + lane) * V;
      if (c < D) {
#pragma unroll
        for (int e = 0; e < V; e += 4)
          store4_stream(sae_out + t * D + c + e, make_float4(acc[i][e], acc[i][e + 1], acc[i][e + 2], acc[i][e + 3]));
        if (target) {
          float err[V];
#pragma unroll
          for (int e = 0; e < V; e += 4) {
            const float4 xv = load4_stream(target + t * D + c + e);
            err[e] = acc[i][e] - xv.x; err[e + 1] = acc[i][e + 1] - xv.y;
            err[e + 2] = acc[i][e + 2] - xv.z; err[e + 3] = acc[i][e + 3] - xv.w;
            sq += (double)(err[e] * err[e] + err[e + 1] * err[e + 1]) +
                  (double)(err[e + 2] * err[e + 2] + err[e + 3] * err[e + 3]);
          }
          if (resid) {
#pragma unroll
            for (int e = 0; e < V; e += 4)
              store4(resid + t * D + c + e, make_float4(err[e], err[e + 1], err[e + 2], err[e + 3]));
          }
#pragma unroll
          for (int e = 0; e < V; ++e) csum[i][e] += err[e];
        }
      }
    }
  }
  if (colsum) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = (i * 32 + lane) * V;
      if (c < D) {
#pragma unroll
        for (int e = 0; e < V; ++e) atomicAdd(colsum_s + c + e, csum[i][e]);
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) atomicAdd(colsum + c, colsum_s[c]);
  }
  if (sse) {
    const double tot = block_sum(sq, scratch);
    if (threadIdx.x == 0) atomicAdd(sse, tot);
  }
}

// ------------------------------------------------------------------------------------------------ decode
// One warp per token.  Lane l owns channels {4l + 128i .. +3}; the k gathered decoder rows stream through
// registers 4 at a time (independent 16-byte loads in flight), fp32 accumulate.
template <typename WT, typename RT, int CH>
__global__ void __launch_bounds__(256) decode_kernel(const float* __restrict__ top_vals,
                                                     const int32_t* __restrict__ top_idx, const WT* __restrict__ W,
                                                     const float* __restrict__ b_dec, const float* __restrict__ target,
                                                     float* __restrict__ sae_out, RT* __restrict__ resid,
                                                     double* __restrict__ sse, float* __restrict__ colsum, int64_t N,
                                                     int d, int k) {
  __shared__ double scratch[32];
  extern __shared__ float colsum_s[];  // [d] per-CTA partial of the residual column sums
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  if (colsum) {
    for (int c = threadIdx.x; c < d; c += blockDim.x) colsum_s[c] = 0.f;
    __syncthreads();
  }
  double sq = 0.0;
  float4 csum[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) csum[i] = make_float4(0, 0, 0, 0);
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); t < N;
       t += static_cast<int64_t>(gridDim.x) * warps_per_block) {
    float4 acc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = lane * 4 + i * 128;
      acc[i] = c < d ? load4(b_dec + c) : make_float4(0, 0, 0, 0);
    }
    for (int j0 = 0; j0 < k; j0 += 32) {
      const int jn = min(32, k - j0);
      const float my_a = lane < jn ? __ldg(top_vals + t * k + j0 + lane) : 0.f;
      const int my_i = lane < jn ? __ldg(top_idx + t * k + j0 + lane) : 0;
#pragma unroll 8
      for (int j = 0; j < jn; ++j) {
        const float a = __shfl_sync(0xffffffffu, my_a, j);
        const int64_t f = __shfl_sync(0xffffffffu, my_i, j);
        if (f < 0) continue;  // warp-uniform: entry owned by another feature shard
        const WT* wr = W + f * d;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const int c = lane * 4 + i * 128;
          if (c < d) {
            const float4 w = load4(wr + c);
            acc[i].x = fmaf(a, w.x, acc[i].x);
            acc[i].y = fmaf(a, w.y, acc[i].y);
            acc[i].z = fmaf(a, w.z, acc[i].z);
            acc[i].w = fmaf(a, w.w, acc[i].w);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = lane * 4 + i * 128;
      if (c < d) {
        store4(sae_out + t * d + c, acc[i]);
        if (target) {
          const float4 xv = load4(target + t * d + c);
          const float4 e = make_float4(acc[i].x - xv.x, acc[i].y - xv.y, acc[i].z - xv.z, acc[i].w - xv.w);
          if (resid) store4(resid + t * d + c, e);
          sq += (double)(e.x * e.x + e.y * e.y) + (double)(e.z * e.z + e.w * e.w);
          csum[i].x += e.x; csum[i].y += e.y; csum[i].z += e.z; csum[i].w += e.w;
        }
      }
    }
  }
  if (colsum) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = lane * 4 + i * 128;
      if (c < d) {
        atomicAdd(colsum_s + c + 0, csum[i].x);
        atomicAdd(colsum_s + c + 1, csum[i].y);
        atomicAdd(colsum_s + c + 2, csum[i].z);
        atomicAdd(colsum_s + c + 3, csum[i].w);
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += blockDim.x) atomicAdd(colsum + c, colsum_s[c]);
  }
  if (sse) {
    const double tot = block_sum(sq, scratch);
    if (threadIdx.x == 0) atomicAdd(sse, tot);
  }
}

// ------------------------------------------------------------------------------------------------ dacts
// One warp per token: g row in registers, each gathered decoder row reduced to a per-lane partial dot;
// 32 rows are reduced together with a 31-shuffle transposing reduction so lane j ends with row j's dot.
// REFINE (fp32 mode of the fused encoder): g = x, `sub` = b_dec, W = W_enc, `bias` = b_enc and the output is
// relu(dot + bias[idx]) -- the selected pre-activations recomputed with an fp32 FMA chain.  The tensor-core product
// that SELECTS them accumulates 3 x d/8 partial sums in TMEM with truncating adds, good for ~1e-6 of a value; the
// gradients (d act = e . W_dec[f] with e nearly orthogonal to the selected rows) need the values themselves to ~1e-7.
template <typename GT, typename WT, int CH, bool REFINE = false>
__global__ void __launch_bounds__(256) dacts_kernel(const GT* __restrict__ g, const int32_t* __restrict__ top_idx,
                                                    const WT* __restrict__ W, float* __restrict__ dacts, int64_t N,
                                                    int d, int k, const float* __restrict__ sub = nullptr,
                                                    const float* __restrict__ bias = nullptr) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); t < N;
       t += static_cast<int64_t>(gridDim.x) * warps_per_block) {
    float4 gv[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = lane * 4 + i * 128;
      gv[i] = c < d ? load4(g + t * d + c) : make_float4(0, 0, 0, 0);
      if constexpr (REFINE) {
        if (c < d) {
          const float4 sb = load4(sub + c);
          gv[i] = make_float4(gv[i].x - sb.x, gv[i].y - sb.y, gv[i].z - sb.z, gv[i].w - sb.w);
        }
      }
    }
    for (int j0 = 0; j0 < k; j0 += 32) {
      const int jn = min(32, k - j0);
      const int my_i = lane < jn ? __ldg(top_idx + t * k + j0 + lane) : -1;
      float part[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int64_t f = __shfl_sync(0xffffffffu, my_i, j);
        float s = 0.f;
        if (f >= 0) {
          const WT* wr = W + f * d;
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            const int c = lane * 4 + i * 128;
            if (c < d) {
              const float4 w = load4(wr + c);
              s = fmaf(gv[i].x, w.x, s);
              s = fmaf(gv[i].y, w.y, s);
              s = fmaf(gv[i].z, w.z, s);
              s = fmaf(gv[i].w, w.w, s);
            }
          }
        }
        part[j] = s;
      }
      const float tot = warp_transpose_reduce32(part, lane);
      if constexpr (REFINE) {
        if (lane < jn && my_i >= 0) dacts[t * k + j0 + lane] = fmaxf(tot + __ldg(bias + my_i), 0.f);
      } else {
        if (lane < jn) dacts[t * k + j0 + lane] = tot;
      }
    }
  }
}

// Fixed-width twin of dacts_kernel: rows fetched R at a time, all loads of a batch in flight before the dots.
template <typename GT, typename WT, int D>
__global__ void __launch_bounds__(256) dacts_fixed_kernel(const GT* __restrict__ g, const int32_t* __restrict__ top_idx,
                                                          const WT* __restrict__ W, float* __restrict__ dacts,
                                                          int64_t N, int k) {
  using SL = typename SliceFor<WT, D>::type;
  constexpr int V = SL::kVec;
  constexpr int CH = (D + 32 * V - 1) / (32 * V);
  constexpr int R = CH <= 3 ? 8 : (CH <= 6 ? 4 : 2);
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const uint64_t keep = l2_keep_policy();
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); t < N;
       t += static_cast<int64_t>(gridDim.x) * warps_per_block) {
    float gv[CH][V];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = (i * 32 + lane) * V;
#pragma unroll
      for (int e = 0; e < V; e += 4) {
        const float4 v = c < D ? load4(g + t * D + c + e) : make_float4(0, 0, 0, 0);
        gv[i][e] = v.x; gv[i][e + 1] = v.y; gv[i][e + 2] = v.z; gv[i][e + 3] = v.w;
      }
    }
    for (int j0 = 0; j0 < k; j0 += 32) {
      const int jn = min(32, k - j0);
      const int my_i = lane < jn ? __ldg(top_idx + t * k + j0 + lane) : -1;
      const unsigned live = __ballot_sync(0xffffffffu, my_i >= 0);
      float part[32];
#pragma unroll
      for (int jb = 0; jb < 32; jb += R) {
        if ((live >> jb) & ((1u << R) - 1u)) {  // warp-uniform: skip batches with no owned entry
          typename SL::Raw raw[R][CH];
#pragma unroll
          for (int u = 0; u < R; ++u) {
            const int fi = __shfl_sync(0xffffffffu, my_i, jb + u);
            const WT* wr = W + static_cast<int64_t>(fi < 0 ? 0 : fi) * D;
#pragma unroll
            for (int i = 0; i < CH; ++i) {
              const int c = (i * 32 + lane) * V;
              if (c < D) raw[u][i] = SL::load_keep(wr + c, keep);
            }
          }
#pragma unroll
          for (int u = 0; u < R; ++u) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < CH; ++i) {
              const int c = (i * 32 + lane) * V;
              if (c < D) {
                float w[V];
                SL::unpack(raw[u][i], w);
#pragma unroll
                for (int e = 0; e < V; ++e) s = fmaf(gv[i][e], w[e], s);
              }
            }
            part[jb + u] = ((live >> (jb + u)) & 1u) ? s : 0.f;
          }
        } else {
#pragma unroll
          for (int u = 0; u < R; ++u) part[jb + u] = 0.f;
        }
      }
      const float tot = warp_transpose_reduce32(part, lane);
      if (lane < jn) dacts[t * k + j0 + lane] = tot;
    }
  }
}

template <typename OT>
__global__ void __launch_bounds__(256) axpby_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                    const float* __restrict__ coef, OT* __restrict__ out, int64_t n4) {
  const float alpha = coef[0], beta = coef[1];
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 av = load4(a + i * 4);
    float4 o = make_float4(alpha * av.x, alpha * av.y, alpha * av.z, alpha * av.w);
    if (b) {
      const float4 bv = load4(b + i * 4);
      o.x = fmaf(beta, bv.x, o.x); o.y = fmaf(beta, bv.y, o.y); o.z = fmaf(beta, bv.z, o.z); o.w = fmaf(beta, bv.w, o.w);
    }
    store4(out + i * 4, o);
  }
}

// ------------------------------------------------------------------------------------------------ CSC index
__global__ void __launch_bounds__(256) csc_hist_kernel(const int32_t* __restrict__ top_idx, int64_t total,
                                                       int32_t* __restrict__ counts) {
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < total;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int32_t f = __ldg(top_idx + p);
    if (f >= 0) atomicAdd(counts + f, 1);  // negative = entry owned by another feature shard
  }
}

// Block-wide exclusive scan of one value per thread (1024 threads); returns the exclusive prefix, *total = sum.
__device__ __forceinline__ int32_t block_excl_scan_1024(int32_t v, int32_t* warp_tot, int32_t* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) warp_tot[w] = incl;
  __syncthreads();
  if (w == 0) {
    int32_t t = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t u = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += u;
    }
    warp_tot[lane] = t;
  }
  __syncthreads();
  *total = warp_tot[31];
  return (w > 0 ? warp_tot[w - 1] : 0) + incl - v;
}

// Single-block exclusive scan of counts[0..n) -> offsets[0..n]; cursor[f] = offsets[f]; cursor[n] = 0 (the long-list
// work counter of the sort kernels).  Tiles of 4096 counts: thread i owns the 16-byte run {4i .. 4i+3} of a tile, so the
// loads and stores are coalesced (the first version gave every thread one long private run: 24 strided loads per
// thread, 32 us at n = 24 576); the next tile's loads are issued before the block scan of the current one.
__global__ void __launch_bounds__(1024) csc_scan_kernel(int32_t* __restrict__ offsets, int32_t* __restrict__ cursor,
                                                        int n) {
  __shared__ int32_t warp_tot[32];
  const int tiles = (n + 4095) / 4096;
  int32_t carry = 0;
  auto load_tile = [&](int tile) {
    const int i0 = tile * 4096 + static_cast<int>(threadIdx.x) * 4;
    int4 v = make_int4(0, 0, 0, 0);
    if (i0 + 3 < n && (n & 3) == 0) {
      v = *reinterpret_cast<const int4*>(offsets + i0);
    } else {
      if (i0 < n) v.x = offsets[i0];
      if (i0 + 1 < n) v.y = offsets[i0 + 1];
      if (i0 + 2 < n) v.z = offsets[i0 + 2];
      if (i0 + 3 < n) v.w = offsets[i0 + 3];
    }
    return v;
  };
  int4 nxt = load_tile(0);
  for (int tile = 0; tile < tiles; ++tile) {
    const int4 v = nxt;
    if (tile + 1 < tiles) nxt = load_tile(tile + 1);
    int32_t total;
    const int32_t run = carry + block_excl_scan_1024(v.x + v.y + v.z + v.w, warp_tot, &total);
    const int4 o = make_int4(run, run + v.x, run + v.x + v.y, run + v.x + v.y + v.z);
    const int i0 = tile * 4096 + static_cast<int>(threadIdx.x) * 4;
    if (i0 + 3 < n && (n & 3) == 0) {
      *reinterpret_cast<int4*>(offsets + i0) = o;
      *reinterpret_cast<int4*>(cursor + i0) = o;
    } else {
      if (i0 < n) offsets[i0] = cursor[i0] = o.x;
      if (i0 + 1 < n) offsets[i0 + 1] = cursor[i0 + 1] = o.y;
      if (i0 + 2 < n) offsets[i0 + 2] = cursor[i0 + 2] = o.z;
      if (i0 + 3 < n) offsets[i0 + 3] = cursor[i0 + 3] = o.w;
    }
    carry += total;
    __syncthreads();  // warp_tot is reused by the next tile
  }
  if (threadIdx.x == 0) {
    offsets[n] = carry;
    cursor[n] = 0;
  }
}

__global__ void __launch_bounds__(256) csc_fill_kernel(const int32_t* __restrict__ top_idx, int64_t total,
                                                       int32_t* __restrict__ cursor, int32_t* __restrict__ entries) {
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < total;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int32_t f = __ldg(top_idx + p);
    if (f < 0) continue;
    const int32_t pos = atomicAdd(cursor + f, 1);
    entries[pos] = static_cast<int32_t>(p);
  }
}

// Sort each feature's list ascending (token order) so the fp32 accumulation order -- and therefore the
// gradients -- are run-to-run deterministic despite the atomic fill.  Lists of up to kWarpSortMax entries: one WARP per
// feature, bitonic network over registers and shuffles; longer lists up to kSortMax: one CTA per queued feature, bitonic
// sort in shared memory; beyond that they are left in fill order (still correct, not bit-stable).
constexpr int kSortMax = 4096;
constexpr int kWarpSortMax = 1024;

// Ascending bitonic sort of 32 * E keys held E per lane; element e of lane l has index l * E + e.  Exchanges at distance
// j < E stay inside a lane (static register indices), the others cross lanes with one shuffle per element.  No shared
// memory, no barrier: a 512-key sort is ~1400 instructions of one warp (the CTA-wide shared-memory network with a
// barrier per stage took 26 us per C2 list, 157 us for the 6144 lists of a C2 step).
template <int E>
__device__ __forceinline__ void warp_bitonic_sort_asc(int32_t (&a)[E], int lane) {
#pragma unroll
  for (int k = 2; k <= 32 * E; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= E) {  // partner element: same e, lane ^ (j / E)
        const int lj = j / E;
        // ascending block iff bit k of the index is clear; k >= 2 j >= 2 E here, so the bit lies in the lane number
        const bool up = k >= 32 * E ? true : ((lane & (k / E)) == 0);
        const bool lower = (lane & lj) == 0;
        const bool take_min = lower == up;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int32_t o = __shfl_xor_sync(0xffffffffu, a[e], lj);
          a[e] = take_min ? min(a[e], o) : max(a[e], o);
        }
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int l = e ^ j;
          if (l > e) {
            // direction: bit k of the index l * E + e -- inside e when k < E, else in the lane number
            const bool up = k < E ? ((e & k) == 0) : (k >= 32 * E ? true : ((lane & (k / E)) == 0));
            const int32_t lo = min(a[e], a[l]), hi = max(a[e], a[l]);
            a[e] = up ? lo : hi;
            a[l] = up ? hi : lo;
          }
        }
      }
    }
  }
}

template <int E>
__device__ __forceinline__ void csc_sort_list(int32_t* __restrict__ list, int len, int lane) {
  int32_t a[E];
#pragma unroll
  for (int e = 0; e < E; ++e) a[e] = e * 32 + lane < len ? list[e * 32 + lane] : 0x7fffffff;  // coalesced, any order
  warp_bitonic_sort_asc<E>(a, lane);
#pragma unroll
  for (int e = 0; e < E; ++e)
    if (lane * E + e < len) list[lane * E + e] = a[e];
}

__global__ void __launch_bounds__(256) csc_sort_warp_kernel(const int32_t* __restrict__ offsets,
                                                            int32_t* __restrict__ entries, int32_t* __restrict__ worklist,
                                                            int n) {
  const int lane = threadIdx.x & 31;
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (f >= n) return;
  const int beg = offsets[f], len = offsets[f + 1] - beg;
  if (len > kWarpSortMax) {  // left to csc_sort_kernel: queue the feature (cursor is free once the fill is done)
    if (lane == 0 && len <= kSortMax) worklist[atomicAdd(worklist + n, 1)] = f;
    return;
  }
  if (len <= 1) return;
  int32_t* list = entries + beg;
  if (len <= 32) csc_sort_list<1>(list, len, lane);
  else if (len <= 64) csc_sort_list<2>(list, len, lane);
  else if (len <= 128) csc_sort_list<4>(list, len, lane);
  else if (len <= 256) csc_sort_list<8>(list, len, lane);
  else if (len <= 512) csc_sort_list<16>(list, len, lane);
  else csc_sort_list<32>(list, len, lane);
}
// Lists of kWarpSortMax + 1 .. kSortMax entries, queued by csc_sort_warp_kernel: a fixed grid walks the queue (one CTA
// per feature used to be launched -- n CTAs that almost all returned at once cost 16 us at n = 24 576).
__global__ void __launch_bounds__(256) csc_sort_kernel(const int32_t* __restrict__ offsets,
                                                       int32_t* __restrict__ entries,
                                                       const int32_t* __restrict__ worklist, int n) {
  __shared__ int32_t buf[kSortMax];
  const int count = worklist[n];
  for (int item = blockIdx.x; item < count; item += gridDim.x) {
  __syncthreads();  // buf is reused
  const int f = worklist[item];
  const int beg = offsets[f], len = offsets[f + 1] - beg;
  int m = 2;
  while (m < len) m <<= 1;
  for (int i = threadIdx.x; i < m; i += blockDim.x) buf[i] = i < len ? entries[beg + i] : 0x7fffffff;
  __syncthreads();
  for (int k = 2; k <= m; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const int32_t a = buf[i], b = buf[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            buf[i] = b;
            buf[l] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < len; i += blockDim.x) entries[beg + i] = buf[i];
  }
}

// ------------------------------------------------------------------------------------------------ weight grads
// Work item = (feature f, chunk of <= kChunk list entries); real activations make the lists wildly uneven (a few
// dense features fire on most tokens), so long lists are split over many CTAs.  Each CTA stages its chunk's
// (token, a*s_dec, dpre) triples in shared memory, then `groups` row-teams walk the entries, every thread owning
// one 16-byte column slice of the gathered g / xc rows (whole contiguous rows: coalesced), 4 entries in flight:
//   dW_dec[f,:] += a_p * g[t_p,:]      dW_enc[f,:] += dpre_p * xc[t_p,:]      db_enc[f] += dpre_p
// Single-chunk features store their rows; multi-chunk features accumulate with vector atomics into zeroed rows.
constexpr int kChunk = 1024;
constexpr int kShortList = 192;  // lists up to this length: one row-team per feature, no block-level machinery

__global__ void __launch_bounds__(1024) chunk_scan_kernel(const int32_t* __restrict__ offsets,
                                                          int32_t* __restrict__ chunk_off, int n, int short_len) {
  __shared__ int32_t warp_tot[32];
  const int per = (n + 1023) / 1024;
  const int lo = min(n, static_cast<int>(threadIdx.x) * per), hi = min(n, lo + per);
  auto chunks = [&](int i) {
    const int32_t len = offsets[i + 1] - offsets[i];
    return len > short_len ? (len + kChunk - 1) / kChunk : 0;  // short lists belong to the warp kernel
  };
  int32_t sum = 0;
  for (int i = lo; i < hi; ++i) sum += chunks(i);
  int32_t total;
  int32_t run = block_excl_scan_1024(sum, warp_tot, &total);
  for (int i = lo; i < hi; ++i) {
    chunk_off[i] = run;
    run += chunks(i);
  }
  if (threadIdx.x == 0) chunk_off[n] = total;
}

// List-ordered metadata of the CSC entries, so the gradient kernels read (token, a * s_dec, dpre) with coalesced
// loads instead of chasing entries -> top_vals / dacts:  dpre = ReLU mask of the selected pre-activation * s_enc.
__global__ void __launch_bounds__(256) csc_meta_kernel(const int32_t* __restrict__ offsets,
                                                       const int32_t* __restrict__ entries,
                                                       const float* __restrict__ top_vals,
                                                       const float* __restrict__ dacts, const float* __restrict__ scales,
                                                       int32_t* __restrict__ tok, float* __restrict__ a_sc,
                                                       float* __restrict__ dp_sc, int n, int k) {
  const int total = offsets[n];  // entries owned by other feature shards are not in the index
  const float s_dec = scales[0], s_enc = scales[1];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int p = __ldg(entries + i);
    const float a = __ldg(top_vals + p);
    tok[i] = p / k;
    a_sc[i] = a * s_dec;
    dp_sc[i] = a > 0.f ? __ldg(dacts + p) * s_enc : 0.f;
  }
}

// Short lists (a wide dictionary has N*k/n entries per feature on average): one WARP per (32-slice column slab,
// feature) item, slab-major (blockIdx.y = slab): the CTAs resident at any moment all gather the same 512-byte
// column slab of g and xc (2 * N * 512 B: 49 MB at N = 48000), which stays in L2 however wide the rows are.
// The warp reads the metadata of up to 32 entries with one coalesced load per array, broadcasts it by shuffle
// (all shuffles of a batch before any load, in converged code), and gathers its 16-byte row slices 8 entries at a
// time from both matrices.  No shared memory, no barrier, direct row stores -- which also zero the rows of silent
// features, so the gradient matrices need no memset.  Rows of long lists are zeroed here and accumulated by the
// chunk kernel.  Lanes past the row end (last slab of a row that is not a multiple of 32 slices) gather column 0
// and skip the stores, so the warp never diverges.
// (launch bound of four CTAs per SM: 64 registers with a 4-byte spill instead of 80 -- 32 resident warps instead of 24
//  was worth 8 % on a kernel that lives on memory-level parallelism)
template <typename GT, typename XT>
__global__ void __launch_bounds__(256, 4) sparse_grads_warp_kernel(
    const int32_t* __restrict__ offsets, const int32_t* __restrict__ tok, const float* __restrict__ a_sc,
    const float* __restrict__ dp_sc, const GT* __restrict__ g, const XT* __restrict__ xc,
    const float* __restrict__ b_dec, float* __restrict__ dW_dec, float* __restrict__ dW_enc,
    float* __restrict__ db_enc, int n, int d, int short_len, int accumulate) {
  constexpr int V = RowVec<GT>::kVec;
  constexpr int R = 8;
  constexpr bool kRecenter = sizeof(XT) == 4;
  const int lane = threadIdx.x & 31;
  const uint64_t keep = l2_keep_policy();
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int slab = blockIdx.y;
  if (f >= n) return;
  const int c = (slab * 32 + lane) * V;
  const bool active = c < d;
  const int cl = active ? c : 0;
  const int beg = offsets[f], len = offsets[f + 1] - beg;
  float* pd = dW_dec + static_cast<int64_t>(f) * d + c;
  float* pe = dW_enc + static_cast<int64_t>(f) * d + c;
  if (len == 0 || len > short_len) {
    if (!accumulate) {
      if (active) {
#pragma unroll
        for (int e = 0; e < V; e += 4) {
          store4_stream(pd + e, make_float4(0, 0, 0, 0));
          store4_stream(pe + e, make_float4(0, 0, 0, 0));
        }
      }
      if (slab == 0 && lane == 0) db_enc[f] = 0.f;
    }
    return;
  }
  float accd[V], acce[V], bd[V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    accd[e] = 0.f;
    acce[e] = 0.f;
    bd[e] = kRecenter ? b_dec[cl + e] : 0.f;
  }
  const GT* gcol = g + cl;
  const XT* xcol = xc + cl;
  float dpsum = 0.f;
  for (int base = 0; base < len; base += 32) {
    const int cnt = min(32, len - base);
    const int my_t = lane < cnt ? __ldg(tok + beg + base + lane) : 0;
    const float my_a = lane < cnt ? __ldg(a_sc + beg + base + lane) : 0.f;
    const float my_dp = lane < cnt ? __ldg(dp_sc + beg + base + lane) : 0.f;
    dpsum += my_dp;
    for (int jb = 0; jb < cnt; jb += R) {  // slots past the list gather row 0 with zero coefficients
      int t[R];
      float a[R], dp[R];
#pragma unroll
      for (int u = 0; u < R; ++u) {
        t[u] = __shfl_sync(0xffffffffu, my_t, jb + u);
        a[u] = __shfl_sync(0xffffffffu, my_a, jb + u);
        dp[u] = __shfl_sync(0xffffffffu, my_dp, jb + u);
      }
      uint4 rg[R], rx[R];
#pragma unroll
      for (int u = 0; u < R; ++u) {
        rg[u] = ldg_keep(gcol + static_cast<int64_t>(t[u]) * d, keep);
        rx[u] = ldg_keep(xcol + static_cast<int64_t>(t[u]) * d, keep);
      }
      if constexpr (sizeof(GT) == 2 && sizeof(XT) == 2) {
        // both rows bf16: one packed fp32 FMA (FFMA2) per bf16 PAIR -- each half an IEEE fma, bit-identical to the
        // scalar chain; the kernel is bound by instruction issue, and the FMAs were 40 % of its instructions
#pragma unroll
        for (int u = 0; u < R; ++u) {
          const uint32_t wg[4] = {rg[u].x, rg[u].y, rg[u].z, rg[u].w};
          const uint32_t wx[4] = {rx[u].x, rx[u].y, rx[u].z, rx[u].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 rd = __ffma2_rn(make_float2(a[u], a[u]),
                                         make_float2(__uint_as_float(wg[q] << 16), __uint_as_float(wg[q] & 0xffff0000u)),
                                         make_float2(accd[2 * q], accd[2 * q + 1]));
            accd[2 * q] = rd.x;
            accd[2 * q + 1] = rd.y;
            const float2 re = __ffma2_rn(make_float2(dp[u], dp[u]),
                                         make_float2(__uint_as_float(wx[q] << 16), __uint_as_float(wx[q] & 0xffff0000u)),
                                         make_float2(acce[2 * q], acce[2 * q + 1]));
            acce[2 * q] = re.x;
            acce[2 * q + 1] = re.y;
          }
        }
      } else {
#pragma unroll
        for (int u = 0; u < R; ++u) {
          float gv[V], xv[V];
          RowVec<GT>::unpack(rg[u], gv);
          RowVec<XT>::unpack(rx[u], xv);
#pragma unroll
          for (int e = 0; e < V; ++e) {
            accd[e] = fmaf(a[u], gv[e], accd[e]);
            acce[e] = fmaf(dp[u], kRecenter ? xv[e] - bd[e] : xv[e], acce[e]);
          }
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int e = 0; e < V; e += 4) {
      float4 od = make_float4(0, 0, 0, 0), oe = make_float4(0, 0, 0, 0);
      if (accumulate) {
        od = load4_stream(pd + e);
        oe = load4_stream(pe + e);
      }
      store4_stream(pd + e, make_float4(od.x + accd[e], od.y + accd[e + 1], od.z + accd[e + 2], od.w + accd[e + 3]));
      store4_stream(pe + e, make_float4(oe.x + acce[e], oe.y + acce[e + 1], oe.z + acce[e + 2], oe.w + acce[e + 3]));
    }
  }
  if (slab == 0) {
    dpsum = warp_sum(dpsum);
    if (lane == 0) db_enc[f] = (accumulate ? db_enc[f] : 0.f) + dpsum;
  }
}

// WHICH: 0 = both gradients in one pass; 1 = dW_dec only (gathers g); 2 = dW_enc + db_enc only (gathers xc).
// Two single-matrix passes keep the gathered working set (one [N,d] bf16 matrix) inside the 126 MB L2.
template <typename GT, typename XT, int WHICH>
__global__ void __launch_bounds__(256) sparse_grads_kernel(
    const int32_t* __restrict__ offsets, const int32_t* __restrict__ chunk_off, const int32_t* __restrict__ tok,
    const float* __restrict__ a_sc, const float* __restrict__ dp_sc, const GT* __restrict__ g,
    const XT* __restrict__ xc, const float* __restrict__ b_dec, float* __restrict__ dW_dec,
    float* __restrict__ dW_enc, float* __restrict__ db_enc, int n, int d) {
  constexpr int V = RowVec<GT>::kVec;
  static_assert(RowVec<XT>::kVec == V, "g and xc share a storage type");
  constexpr bool kRecenter = sizeof(XT) == 4;  // fp32 path: xc = x - b_dec recomputed on the fly
  extern __shared__ float red_s[];             // [groups][2][d] cross-team reduction
  __shared__ int32_t tok_s[kChunk];
  __shared__ float a_s[kChunk];
  __shared__ float dp_s[kChunk];
  __shared__ int f_s;
  __shared__ float wsum[8];
  const int item = blockIdx.x;
  if (item >= chunk_off[n]) return;
  if (threadIdx.x == 0) {  // binary search: last f with chunk_off[f] <= item
    int lo = 0, hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (chunk_off[mid] <= item) lo = mid; else hi = mid - 1;
    }
    f_s = lo;
  }
  __syncthreads();
  const int f = f_s;
  const int nchunks = chunk_off[f + 1] - chunk_off[f];
  const int beg = offsets[f] + (item - chunk_off[f]) * kChunk;
  const int cnt = min(kChunk, offsets[f + 1] - beg);
  float dpsum = 0.f;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const float dp = dp_sc[beg + i];
    tok_s[i] = tok[beg + i];
    a_s[i] = a_sc[beg + i];
    dp_s[i] = dp;
    dpsum += dp;
  }
  __syncthreads();
  const int tpr = (d + V - 1) / V;              // threads per row
  const int col_passes = (tpr + 255) / 256;     // > 1 only for very wide rows
  const int tpr_eff = col_passes > 1 ? 256 : tpr;
  const int groups = col_passes > 1 ? 1 : max(1, 256 / tpr);
  const int team = threadIdx.x / tpr_eff;
  const int lane_c = threadIdx.x - team * tpr_eff;
  const bool multi = nchunks > 1;
  for (int cp = 0; cp < col_passes; ++cp) {
    const int c = (cp * 256 + lane_c) * V;
    const bool active = team < groups && c < d;
    float accd[V], acce[V], bd[V];
#pragma unroll
    for (int u = 0; u < V; ++u) { accd[u] = 0.f; acce[u] = 0.f; bd[u] = 0.f; }
    if (active) {
      if (kRecenter) {
#pragma unroll
        for (int u = 0; u < V; ++u) bd[u] = b_dec[c + u];
      }
      int i = team;
      for (; i + 3 * groups < cnt; i += 4 * groups) {
        float gv[4][V], xv[4][V];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t t = tok_s[i + u * groups];
          if (WHICH != 2) RowVec<GT>::load(g + t * d + c, gv[u]);
          if (WHICH != 1) RowVec<XT>::load(xc + t * d + c, xv[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float a = a_s[i + u * groups], dp = dp_s[i + u * groups];
#pragma unroll
          for (int e = 0; e < V; ++e) {
            if (WHICH != 2) accd[e] = fmaf(a, gv[u][e], accd[e]);
            if (WHICH != 1) acce[e] = fmaf(dp, xv[u][e] - bd[e], acce[e]);
          }
        }
      }
      for (; i < cnt; i += groups) {
        const int64_t t = tok_s[i];
        float gv[V], xv[V];
        if (WHICH != 2) RowVec<GT>::load(g + t * d + c, gv);
        if (WHICH != 1) RowVec<XT>::load(xc + t * d + c, xv);
        const float a = a_s[i], dp = dp_s[i];
#pragma unroll
        for (int e = 0; e < V; ++e) {
          if (WHICH != 2) accd[e] = fmaf(a, gv[e], accd[e]);
          if (WHICH != 1) acce[e] = fmaf(dp, xv[e] - bd[e], acce[e]);
        }
      }
    }
    // cross-team reduction through shared memory, then one store / vector-atomic per 4 columns
    __syncthreads();
    if (active) {
      float* r = red_s + static_cast<size_t>(team) * 2 * d;
#pragma unroll
      for (int e = 0; e < V; ++e) {
        r[c + e] = accd[e];
        r[d + c + e] = acce[e];
      }
    }
    __syncthreads();
    const int c_lo = cp * 256 * V, c_hi = min(d, c_lo + 256 * V);
    const int span4 = (c_hi - c_lo) / 4;  // d % 4 == 0
    for (int q = threadIdx.x; q < 2 * span4; q += blockDim.x) {
      const int m = q / span4;
      if ((WHICH == 1 && m == 1) || (WHICH == 2 && m == 0)) continue;
      const int cc = c_lo + (q - m * span4) * 4;
      float4 sum = make_float4(0, 0, 0, 0);
      for (int t = 0; t < groups; ++t) {
        const float4 v = *reinterpret_cast<const float4*>(red_s + (static_cast<size_t>(t) * 2 + m) * d + cc);
        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
      }
      float* dst = (m == 0 ? dW_dec : dW_enc) + static_cast<int64_t>(f) * d + cc;
      if (multi) {
        atomicAdd(reinterpret_cast<float4*>(dst), sum);
      } else {
        const float4 old = *reinterpret_cast<const float4*>(dst);  // zero, or earlier decodes when accumulating
        *reinterpret_cast<float4*>(dst) = make_float4(old.x + sum.x, old.y + sum.y, old.z + sum.z, old.w + sum.w);
      }
    }
  }
  dpsum = warp_sum(dpsum);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = dpsum;
  __syncthreads();
  if (threadIdx.x == 0 && WHICH != 1) {
    float tot = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) tot += wsum[w];
    if (multi) atomicAdd(db_enc + f, tot); else db_enc[f] += tot;
  }
}

// db_dec[c] (+)= s*colsum[c] - sum_f db_enc[f] * W_enc[f,c]; grid over column chunks x feature slabs.
__device__ __forceinline__ float ldw(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldw(const __nv_bfloat16* p) {
  return __uint_as_float(static_cast<uint32_t>(__ldg(reinterpret_cast<const unsigned short*>(p))) << 16);
}
template <typename WT>
__global__ void __launch_bounds__(256) bdec_grad_kernel(const float* __restrict__ colsum,
                                                        const float* __restrict__ scales,
                                                        const float* __restrict__ db_enc,
                                                        const WT* __restrict__ W_enc, float* __restrict__ db_dec,
                                                        int n, int d, int slab) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const int f0 = blockIdx.y * slab, f1 = min(n, f0 + slab);
  float acc = 0.f;
  if (db_enc) {
    float part[4] = {0.f, 0.f, 0.f, 0.f};  // four independent chains: the loads of a group are all in flight
    int f = f0;
    for (; f + 4 <= f1; f += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        part[u] = fmaf(-__ldg(db_enc + f + u), ldw(W_enc + static_cast<int64_t>(f + u) * d + c), part[u]);
    }
    for (; f < f1; ++f) part[0] = fmaf(-__ldg(db_enc + f), ldw(W_enc + static_cast<int64_t>(f) * d + c), part[0]);
    acc = (part[0] + part[1]) + (part[2] + part[3]);
  }
  if (blockIdx.y == 0 && colsum) acc = fmaf(scales[0], colsum[c], acc);
  atomicAdd(db_dec + c, acc);
}

__global__ void loss_scalars_kernel(const double* __restrict__ sse, const double* __restrict__ tv,
                                    float* __restrict__ out, double inv_numel) {
  double t = *tv;
  if (t == 0.0) t = 1.0;  // topkautoencoder.py:105-106
  const double s = *sse;
  out[0] = static_cast<float>(s / t);
  out[1] = static_cast<float>(s * inv_numel);
  out[2] = static_cast<float>(2.0 / t);
  out[3] = out[2];
  out[4] = static_cast<float>(t);
}

__global__ void __launch_bounds__(256) dead_update_kernel(const int32_t* __restrict__ offsets,
                                                          int64_t* __restrict__ frames, int n, int64_t n_tokens) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < n) frames[f] = offsets[f + 1] > offsets[f] ? 0 : frames[f] + n_tokens;
}

// One warp per row.
__global__ void __launch_bounds__(256) rownorm_kernel(float* __restrict__ W, int64_t rows, int cols, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  float* w = W + r * cols;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s = fmaf(w[c], w[c], s);
  s = warp_sum(s);
  const float inv = 1.f / (sqrtf(s) + eps);
  for (int c = lane; c < cols; c += 32) w[c] *= inv;
}
__global__ void __launch_bounds__(256) remove_parallel_kernel(float* __restrict__ G, const float* __restrict__ W,
                                                              int64_t rows, int cols) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  float* gr = G + r * cols;
  const float* w = W + r * cols;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s = fmaf(gr[c], w[c], s);
  s = warp_sum(s);
  for (int c = lane; c < cols; c += 32) gr[c] = fmaf(-s, w[c], gr[c]);
}

// dst[r, :] = src[rows[r], :] in 4-byte words (row_words per row); used to compact the dead-latent encoder rows.
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint32_t* __restrict__ src,
                                                          const int32_t* __restrict__ rows,
                                                          uint32_t* __restrict__ dst, int64_t n_rows, int row_words) {
  const int64_t total = n_rows * row_words;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / row_words;
    const int w = static_cast<int>(i - r * row_words);
    dst[i] = __ldg(src + static_cast<int64_t>(rows[r]) * row_words + w);
  }
}
__global__ void __launch_bounds__(256) index_map_kernel(const int32_t* __restrict__ table,
                                                        const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                        int64_t count) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = table[in[i]];
}

// ---------------------------------------------------------------------------------------------- feature sharding
// Global top-32 of G per-shard top-32 lists (each sorted by (value desc, index asc), indices shard-local):
// one warp per token, lane j holds element j; each further shard is folded in with max(cur, reversed other)
// followed by a 5-stage bitonic merge.  Output indices are GLOBAL (local + shard * n_local).
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
  const uint32_t lo = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v), src);
  const uint32_t hi = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), src);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
__global__ void __launch_bounds__(256) shard_merge_kernel(const float* __restrict__ vals, const int32_t* __restrict__ idx,
                                                          float* __restrict__ out_vals, int32_t* __restrict__ out_idx,
                                                          int64_t N, int G, int n_local) {
  const int lane = threadIdx.x & 31;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= N) return;
  uint64_t cur = 0;
  for (int g = 0; g < G; ++g) {
    const int64_t o = (static_cast<int64_t>(g) * N + t) * 32 + lane;
    const uint32_t gi = static_cast<uint32_t>(idx[o] + g * n_local);
    // zero-valued fillers keep value bits 0 but a real index, so they order by index among themselves
    const uint64_t key = (static_cast<uint64_t>(__float_as_uint(vals[o])) << 32) | static_cast<uint32_t>(~gi);
    if (g == 0) {
      cur = key;
    } else {
      const uint64_t rev = shfl64(key, 31 - lane);
      cur = cur > rev ? cur : rev;
#pragma unroll
      for (int j = 16; j > 0; j >>= 1) {
        const uint64_t other = shfl64(cur, lane ^ j);
        const bool lower = (lane & j) == 0;
        const uint64_t mx = cur > other ? cur : other, mn = cur > other ? other : cur;
        cur = lower ? mx : mn;
      }
    }
  }
  out_vals[t * 32 + lane] = __uint_as_float(static_cast<uint32_t>(cur >> 32));
  out_idx[t * 32 + lane] = static_cast<int32_t>(~static_cast<uint32_t>(cur));
}

// Keep the winners this shard owns: local index in [0, n_local) or -1, value or 0.
__global__ void __launch_bounds__(256) shard_localize_kernel(const float* __restrict__ vals,
                                                             const int32_t* __restrict__ gidx,
                                                             float* __restrict__ lvals, int32_t* __restrict__ lidx,
                                                             int64_t count, int lo, int n_local) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int32_t l = gidx[i] - lo;
    const bool own = l >= 0 && l < n_local;
    lidx[i] = own ? l : -1;
    lvals[i] = own ? vals[i] : 0.f;
  }
}

// resid = sae_out - target (+ SSE, column sums) after the partial reconstructions have been reduced.
template <typename RT>
__global__ void __launch_bounds__(256) residual_kernel(const float* __restrict__ sae_out,
                                                       const float* __restrict__ target, RT* __restrict__ resid,
                                                       double* __restrict__ sse, float* __restrict__ colsum,
                                                       int64_t N, int d) {
  __shared__ double scratch[32];
  const int d4 = d >> 2;
  const int c4 = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 4-channel column group
  double sq = 0.0;
  if (c4 < d4) {
    float4 cs = make_float4(0, 0, 0, 0);
    for (int64_t t = blockIdx.y; t < N; t += gridDim.y) {
      const float4 a = load4(sae_out + t * d + c4 * 4), x = load4(target + t * d + c4 * 4);
      const float4 e = make_float4(a.x - x.x, a.y - x.y, a.z - x.z, a.w - x.w);
      if (resid) store4(resid + t * d + c4 * 4, e);
      sq += (double)(e.x * e.x + e.y * e.y) + (double)(e.z * e.z + e.w * e.w);
      cs.x += e.x; cs.y += e.y; cs.z += e.z; cs.w += e.w;
    }
    if (colsum) {
      atomicAdd(colsum + c4 * 4 + 0, cs.x);
      atomicAdd(colsum + c4 * 4 + 1, cs.y);
      atomicAdd(colsum + c4 * 4 + 2, cs.z);
      atomicAdd(colsum + c4 * 4 + 3, cs.w);
    }
  }
  const double tot = block_sum(sq, scratch);
  if (sse && threadIdx.x == 0) atomicAdd(sse, tot);
}

// out[i] = sum_s parts[s, i]  (split-K partial sums)
__global__ void __launch_bounds__(256) sum_splits_kernel(const float* __restrict__ parts, float* __restrict__ out,
                                                         int splits, int64_t n4) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 acc = make_float4(0, 0, 0, 0);
    for (int s = 0; s < splits; ++s) {
      const float4 v = load4(parts + (static_cast<int64_t>(s) * n4 + i) * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    store4(out + i * 4, acc);
  }
}

static inline int grid_for(int64_t work, int block, int max_blocks) {
  int64_t g = (work + block - 1) / block;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace freud

using namespace freud;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int freud_topk_prep_x(const void* x, int x_dtype, const float* b_dec, void* xc_hi, void* xc_lo, double* tv,
                                 float* colmean, int64_t B, int64_t T, int64_t d, int precision, void* stream) {
  FREUD_REQUIRE(B > 0 && T > 0 && d > 0 && d % 4 == 0, "prep_x needs d % 4 == 0");
  FREUD_REQUIRE(x_dtype >= 0 && x_dtype <= 2, "x_dtype: 0 fp32, 1 fp16, 2 bf16");
  FREUD_CHECK_CUDA(cudaMemsetAsync(tv, 0, sizeof(double), STREAM));
  const int64_t work = T * (d / 4);
  const int grid = static_cast<int>((work + 255) / 256);
#define FREUD_PREP(MODE)                                                                                              \
  do {                                                                                                                \
    if (x_dtype == 0)                                                                                                 \
      prep_x_kernel<MODE, float><<<grid, 256, 0, STREAM>>>(static_cast<const float*>(x), b_dec, xc_hi, xc_lo, tv,     \
                                                           colmean, (int)B, T, (int)d);                               \
    else if (x_dtype == 1)                                                                                            \
      prep_x_kernel<MODE, __half><<<grid, 256, 0, STREAM>>>(static_cast<const __half*>(x), b_dec, xc_hi, xc_lo, tv,   \
                                                            colmean, (int)B, T, (int)d);                              \
    else                                                                                                              \
      prep_x_kernel<MODE, __nv_bfloat16><<<grid, 256, 0, STREAM>>>(static_cast<const __nv_bfloat16*>(x), b_dec, xc_hi, \
                                                                   xc_lo, tv, colmean, (int)B, T, (int)d);            \
  } while (0)
  if (precision == FREUD_BF16)
    FREUD_PREP(0);
  else
    FREUD_PREP(1);
#undef FREUD_PREP
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_split_operand(const float* w, void* hi, void* lo, int64_t numel, int precision, void* stream) {
  FREUD_REQUIRE(numel % 4 == 0, "split_operand needs numel % 4 == 0");
  const int grid = grid_for(numel / 4, 256, sm_count() * 8);
  if (precision == FREUD_BF16)
    split_operand_kernel<0><<<grid, 256, 0, STREAM>>>(w, hi, lo, numel / 4);
  else
    split_operand_kernel<1><<<grid, 256, 0, STREAM>>>(w, hi, lo, numel / 4);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

namespace {
template <typename WT, typename RT>
int launch_decode(const float* tv, const int32_t* ti, const void* W, const float* b_dec, const float* target,
                  float* sae_out, void* resid, double* sse, float* colsum, int64_t N, int d, int k, cudaStream_t s) {
  const int grid = grid_for(N, 8, sm_count() * 8);
  const size_t sm = d * sizeof(float);
#define LF(D)                                                                                                    \
  decode_fixed_kernel<WT, RT, D><<<grid, 256, 0, s>>>(tv, ti, static_cast<const WT*>(W), b_dec, target,       \
                                                         sae_out, static_cast<RT*>(resid), sse, colsum, N, k)
  switch (d) {
    case 384: LF(384); return 0;
    case 512: LF(512); return 0;
    case 768: LF(768); return 0;
    case 1024: LF(1024); return 0;
    case 1280: LF(1280); return 0;
    default: break;
  }
#undef LF
#define LD(CH)                                                                                                   \
  decode_kernel<WT, RT, CH><<<grid, 256, sm, s>>>(tv, ti, static_cast<const WT*>(W), b_dec, target, sae_out,     \
                                                   static_cast<RT*>(resid), sse, colsum, N, d, k)
  if (d <= 384) LD(3);
  else if (d <= 768) LD(6);
  else if (d <= 1280) LD(10);
  else LD(16);
#undef LD
  return 0;
}
template <typename GT, typename WT>
int launch_dacts(const void* g, const int32_t* ti, const void* W, float* dacts, int64_t N, int d, int k,
                 cudaStream_t s) {
  const int grid = grid_for(N, 8, sm_count() * 8);
#define LF(D)                                                                                                 \
  dacts_fixed_kernel<GT, WT, D><<<grid, 256, 0, s>>>(static_cast<const GT*>(g), ti, static_cast<const WT*>(W), \
                                                        dacts, N, k)
  switch (d) {
    case 384: LF(384); return 0;
    case 512: LF(512); return 0;
    case 768: LF(768); return 0;
    case 1024: LF(1024); return 0;
    case 1280: LF(1280); return 0;
    default: break;
  }
#undef LF
#define LD(CH) \
  dacts_kernel<GT, WT, CH><<<grid, 256, 0, s>>>(static_cast<const GT*>(g), ti, static_cast<const WT*>(W), dacts, N, d, k)
  if (d <= 384) LD(3);
  else if (d <= 768) LD(6);
  else if (d <= 1280) LD(10);
  else LD(16);
#undef LD
  return 0;
}
}  // namespace

extern "C" int freud_topk_decode(const float* top_vals, const int32_t* top_idx, const void* W_dec, int w_is_bf16,
                                 const float* b_dec, const float* target, float* sae_out, void* resid,
                                 int resid_is_bf16, double* sse, float* colsum, int64_t N, int64_t d, int64_t k,
                                 void* stream) {
  FREUD_REQUIRE(N > 0 && k > 0 && d % 4 == 0 && d <= 2048, "decode needs d % 4 == 0 and d <= 2048");
  FREUD_REQUIRE(target != nullptr || (resid == nullptr && sse == nullptr && colsum == nullptr),
                "residual outputs need a target");
  if (w_is_bf16) {
    if (resid_is_bf16)
      launch_decode<__nv_bfloat16, __nv_bfloat16>(top_vals, top_idx, W_dec, b_dec, target, sae_out, resid, sse, colsum, N, (int)d, (int)k, STREAM);
    else
      launch_decode<__nv_bfloat16, float>(top_vals, top_idx, W_dec, b_dec, target, sae_out, resid, sse, colsum, N, (int)d, (int)k, STREAM);
  } else {
    if (resid_is_bf16)
      launch_decode<float, __nv_bfloat16>(top_vals, top_idx, W_dec, b_dec, target, sae_out, resid, sse, colsum, N, (int)d, (int)k, STREAM);
    else
      launch_decode<float, float>(top_vals, top_idx, W_dec, b_dec, target, sae_out, resid, sse, colsum, N, (int)d, (int)k, STREAM);
  }
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_topk_dacts(const void* g, int g_is_bf16, const int32_t* top_idx, const void* W_dec,
                                int w_is_bf16, float* dacts, int64_t N, int64_t d, int64_t k, void* stream) {
  FREUD_REQUIRE(N > 0 && k > 0 && d % 4 == 0 && d <= 2048, "dacts needs d % 4 == 0 and d <= 2048");
  if (g_is_bf16) {
    if (w_is_bf16) launch_dacts<__nv_bfloat16, __nv_bfloat16>(g, top_idx, W_dec, dacts, N, (int)d, (int)k, STREAM);
    else launch_dacts<__nv_bfloat16, float>(g, top_idx, W_dec, dacts, N, (int)d, (int)k, STREAM);
  } else {
    if (w_is_bf16) launch_dacts<float, __nv_bfloat16>(g, top_idx, W_dec, dacts, N, (int)d, (int)k, STREAM);
    else launch_dacts<float, float>(g, top_idx, W_dec, dacts, N, (int)d, (int)k, STREAM);
  }
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_topk_refine(const float* x, const float* b_dec, const float* W_enc, const float* b_enc,
                                 const int32_t* top_idx, float* top_vals, int64_t N, int64_t d, int64_t k,
                                 void* stream) {
  FREUD_REQUIRE(N > 0 && k > 0 && d % 4 == 0 && d <= 2048, "refine needs d % 4 == 0 and d <= 2048");
  const int grid = grid_for(N, 8, sm_count() * 8);
#define LR(CH)                                                                                                     \
  dacts_kernel<float, float, CH, true><<<grid, 256, 0, STREAM>>>(x, top_idx, W_enc, top_vals, N, (int)d, (int)k, \
                                                                    b_dec, b_enc)
  if (d <= 384) LR(3);
  else if (d <= 768) LR(6);
  else if (d <= 1280) LR(10);
  else LR(16);
#undef LR
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_axpby(const float* a, const float* b, const float* coef, void* out, int out_is_bf16,
                           int64_t numel, void* stream) {
  FREUD_REQUIRE(numel % 4 == 0, "axpby needs numel % 4 == 0");
  const int grid = grid_for(numel / 4, 256, sm_count() * 8);
  if (out_is_bf16)
    axpby_kernel<__nv_bfloat16><<<grid, 256, 0, STREAM>>>(a, b, coef, static_cast<__nv_bfloat16*>(out), numel / 4);
  else
    axpby_kernel<float><<<grid, 256, 0, STREAM>>>(a, b, coef, static_cast<float*>(out), numel / 4);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_gather_rows(const void* src, const int32_t* rows, void* dst, int64_t n_rows, int64_t row_bytes,
                                 void* stream) {
  FREUD_REQUIRE(n_rows > 0 && row_bytes > 0 && row_bytes % 4 == 0, "gather_rows needs row_bytes % 4 == 0");
  const int grid = grid_for(n_rows * (row_bytes / 4), 256, sm_count() * 8);
  gather_rows_kernel<<<grid, 256, 0, STREAM>>>(static_cast<const uint32_t*>(src), rows, static_cast<uint32_t*>(dst),
                                              n_rows, (int)(row_bytes / 4));
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_index_map(const int32_t* table, const int32_t* in, int32_t* out, int64_t count, void* stream) {
  FREUD_REQUIRE(count > 0, "index_map needs count > 0");
  index_map_kernel<<<grid_for(count, 256, sm_count() * 8), 256, 0, STREAM>>>(table, in, out, count);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_shard_merge(const float* vals, const int32_t* idx, float* out_vals, int32_t* out_idx, int64_t N,
                                 int64_t G, int64_t n_local, void* stream) {
  FREUD_REQUIRE(N > 0 && G >= 1 && n_local > 0, "shard_merge: bad sizes");
  shard_merge_kernel<<<(unsigned)((N + 7) / 8), 256, 0, STREAM>>>(vals, idx, out_vals, out_idx, N, (int)G, (int)n_local);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_shard_localize(const float* vals, const int32_t* gidx, float* lvals, int32_t* lidx, int64_t count,
                                    int64_t lo, int64_t n_local, void* stream) {
  FREUD_REQUIRE(count > 0, "shard_localize: empty");
  shard_localize_kernel<<<grid_for(count, 256, sm_count() * 8), 256, 0, STREAM>>>(vals, gidx, lvals, lidx, count, (int)lo,
                                                                                 (int)n_local);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_residual(const float* sae_out, const float* target, void* resid, int resid_is_bf16, double* sse,
                              float* colsum, int64_t N, int64_t d, void* stream) {
  FREUD_REQUIRE(N > 0 && d % 4 == 0, "residual needs d % 4 == 0");
  dim3 grid((unsigned)((d / 4 + 255) / 256), (unsigned)(N < 2048 ? N : 2048));
  if (resid_is_bf16)
    residual_kernel<__nv_bfloat16><<<grid, 256, 0, STREAM>>>(sae_out, target, static_cast<__nv_bfloat16*>(resid), sse, colsum, N, (int)d);
  else
    residual_kernel<float><<<grid, 256, 0, STREAM>>>(sae_out, target, static_cast<float*>(resid), sse, colsum, N, (int)d);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_sum_splits(const float* parts, float* out, int64_t splits, int64_t numel, void* stream) {
  FREUD_REQUIRE(numel > 0 && numel % 4 == 0 && splits >= 1, "sum_splits needs numel % 4 == 0");
  sum_splits_kernel<<<grid_for(numel / 4, 256, sm_count() * 8), 256, 0, STREAM>>>(parts, out, (int)splits, numel / 4);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_csc_build(const int32_t* top_idx, int64_t N, int64_t k, int64_t n, int32_t* offsets,
                               int32_t* entries, int32_t* cursor, int counts_ready, void* stream) {
  const int64_t total = N * k;
  FREUD_REQUIRE(total > 0 && total < (1ll << 31) && n > 0 && n < (1ll << 31), "csc sizes out of range");
  const int grid = grid_for(total, 256, sm_count() * 8);
  if (!counts_ready) {  // else offsets[0..n) already holds the per-feature counts (freud_topk_encode's hist output)
    FREUD_CHECK_CUDA(cudaMemsetAsync(offsets, 0, (n + 1) * sizeof(int32_t), STREAM));
    csc_hist_kernel<<<grid, 256, 0, STREAM>>>(top_idx, total, offsets);
  }
  csc_scan_kernel<<<1, 1024, 0, STREAM>>>(offsets, cursor, (int)n);
  csc_fill_kernel<<<grid, 256, 0, STREAM>>>(top_idx, total, cursor, entries);
  csc_sort_warp_kernel<<<(int)((n + 7) / 8), 256, 0, STREAM>>>(offsets, entries, cursor, (int)n);
  csc_sort_kernel<<<(int)std::min<int64_t>(n, sm_count() * 4), 256, 0, STREAM>>>(offsets, entries, cursor, (int)n);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_csc_meta(const int32_t* offsets, const int32_t* entries, const float* top_vals,
                              const float* dacts, const float* scales, int32_t* meta, int64_t n_entries, int64_t n,
                              int64_t k, void* stream) {
  FREUD_REQUIRE(n_entries > 0 && n_entries < (1ll << 31) && n > 0 && k > 0, "csc_meta: sizes out of range");
  csc_meta_kernel<<<grid_for(n_entries, 256, sm_count() * 8), 256, 0, STREAM>>>(
      offsets, entries, top_vals, dacts, scales, meta, reinterpret_cast<float*>(meta + n_entries),
      reinterpret_cast<float*>(meta + 2 * n_entries), (int)n, (int)k);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_topk_sparse_grads(const int32_t* offsets, const int32_t* meta, const void* g, int g_is_bf16,
                                       const void* xc, int xc_is_bf16, const float* b_dec, float* dW_dec,
                                       float* dW_enc, float* db_enc, int32_t* chunk_off, int64_t n_entries,
                                       int64_t n, int64_t d, int64_t k, int accumulate, void* stream) {
  FREUD_REQUIRE(n > 0 && d % 4 == 0, "sparse_grads needs d % 4 == 0");
  FREUD_REQUIRE(g_is_bf16 == xc_is_bf16, "g and xc must share a storage type");
  FREUD_REQUIRE(!g_is_bf16 || d % 8 == 0, "bf16 rows need d % 8 == 0");
  FREUD_REQUIRE(xc_is_bf16 || b_dec != nullptr, "fp32 path recomputes x - b_dec and needs b_dec");
  FREUD_REQUIRE(n_entries > 0 && n_entries < (1ll << 31), "sparse_grads: entry count out of range");
  const int32_t* tok = meta;
  const float* a_sc = reinterpret_cast<const float*>(meta + n_entries);
  const float* dp_sc = reinterpret_cast<const float*>(meta + 2 * n_entries);
  const int V = g_is_bf16 ? 8 : 4;
  const int tpr = (int)((d + V - 1) / V);
  // short lists (and the zero fill of every other row): one warp per (feature, 32-slice slab)
  const int slabs = (tpr + 31) / 32;
  const dim3 wgrid((unsigned)((n + 7) / 8), (unsigned)slabs);
  if (g_is_bf16)
    sparse_grads_warp_kernel<__nv_bfloat16, __nv_bfloat16><<<wgrid, 256, 0, STREAM>>>(
        offsets, tok, a_sc, dp_sc, static_cast<const __nv_bfloat16*>(g), static_cast<const __nv_bfloat16*>(xc), b_dec,
        dW_dec, dW_enc, db_enc, (int)n, (int)d, kShortList, accumulate);
  else
    sparse_grads_warp_kernel<float, float><<<wgrid, 256, 0, STREAM>>>(
        offsets, tok, a_sc, dp_sc, static_cast<const float*>(g), static_cast<const float*>(xc), b_dec, dW_dec, dW_enc,
        db_enc, (int)n, (int)d, kShortList, accumulate);
  FREUD_CHECK_CUDA(cudaGetLastError());
  chunk_scan_kernel<<<1, 1024, 0, STREAM>>>(offsets, chunk_off, (int)n, kShortList);
  // chunk items only exist for lists longer than kShortList: at most n_entries / kShortList such features
  int64_t max_items = n_entries / kChunk + n;
  if (n_entries / kChunk + n_entries / kShortList + 1 < max_items)
    max_items = n_entries / kChunk + n_entries / kShortList + 1;
  const int groups = tpr > 256 ? 1 : (256 / tpr > 0 ? 256 / tpr : 1);
  const size_t smem = static_cast<size_t>(groups) * 2 * d * sizeof(float);
  FREUD_REQUIRE(smem + 3 * kChunk * 4 + 64 <= 227 * 1024, "activation size too wide for sparse_grads");
  // one pass per gradient matrix when the two gathered [N,d] matrices together would not stay in L2
  const int64_t N_tok = n_entries / k;
  const bool split = 2 * N_tok * d * (g_is_bf16 ? 2 : 4) > (96ll << 20);
#define FREUD_SG_LAUNCH(GT, WHICH)                                                                                   \
  do {                                                                                                               \
    auto kern = sparse_grads_kernel<GT, GT, WHICH>;                                                                  \
    if (smem > 32 * 1024)                                                                                            \
      FREUD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
    kern<<<(unsigned)max_items, 256, smem, STREAM>>>(offsets, chunk_off, tok, a_sc, dp_sc,                           \
                                                     static_cast<const GT*>(g), static_cast<const GT*>(xc), b_dec,   \
                                                     dW_dec, dW_enc, db_enc, (int)n, (int)d);                        \
  } while (0)
  if (g_is_bf16) {
    if (split) {
      FREUD_SG_LAUNCH(__nv_bfloat16, 1);
      FREUD_SG_LAUNCH(__nv_bfloat16, 2);
    } else {
      FREUD_SG_LAUNCH(__nv_bfloat16, 0);
    }
  } else {
    if (split) {
      FREUD_SG_LAUNCH(float, 1);
      FREUD_SG_LAUNCH(float, 2);
    } else {
      FREUD_SG_LAUNCH(float, 0);
    }
  }
#undef FREUD_SG_LAUNCH
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_topk_bdec_grad(const float* colsum, const float* scales, const float* db_enc, const void* W_enc,
                                    int w_is_bf16, float* db_dec, int64_t n, int64_t d, int accumulate, void* stream) {
  FREUD_REQUIRE(d > 0, "bdec_grad needs d > 0");
  FREUD_REQUIRE(db_enc == nullptr || W_enc != nullptr, "db_enc term needs W_enc");
  FREUD_REQUIRE(colsum == nullptr || scales != nullptr, "colsum term needs scales");
  if (!accumulate) FREUD_CHECK_CUDA(cudaMemsetAsync(db_dec, 0, d * sizeof(float), STREAM));
  const int slab = 64;  // n/64 x d/256 CTAs: enough of them in flight to stream W_enc at HBM rate
  dim3 grid((unsigned)((d + 255) / 256), (unsigned)(db_enc ? (n + slab - 1) / slab : 1));
  if (w_is_bf16)
    bdec_grad_kernel<<<grid, 256, 0, STREAM>>>(colsum, scales, db_enc, static_cast<const __nv_bfloat16*>(W_enc), db_dec,
                                               (int)n, (int)d, slab);
  else
    bdec_grad_kernel<<<grid, 256, 0, STREAM>>>(colsum, scales, db_enc, static_cast<const float*>(W_enc), db_dec, (int)n,
                                               (int)d, slab);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_topk_loss_scalars(const double* sse, const double* tv, float* out, int64_t numel, void* stream) {
  loss_scalars_kernel<<<1, 1, 0, STREAM>>>(sse, tv, out, 1.0 / (double)numel);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_dead_latent_update(const int32_t* offsets, int64_t* frames, int64_t n, int64_t n_tokens,
                                        void* stream) {
  dead_update_kernel<<<(int)((n + 255) / 256), 256, 0, STREAM>>>(offsets, frames, (int)n, n_tokens);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_rownorm_project(float* W, int64_t rows, int64_t cols, float eps, void* stream) {
  rownorm_kernel<<<(int)((rows + 7) / 8), 256, 0, STREAM>>>(W, rows, (int)cols, eps);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int freud_remove_parallel_grad(float* G, const float* W, int64_t rows, int64_t cols, void* stream) {
  remove_parallel_kernel<<<(int)((rows + 7) / 8), 256, 0, STREAM>>>(G, W, rows, (int)cols);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}
