// Per-feature max-activation search (src/utils/activations.py:41-132): streaming per-file statistics over the
// stored activations of ONE feature, then a small exact ranking.  HBM-bound scans; no tensor cores.
#include "device_utils.cuh"
#include "host_common.h"
#include "../../include/freud_b200.h"

#include <cuda_fp16.h>

namespace freud {

struct Stat {
  float vmax;  // running max of the signed trace
  int amax;    // first index attaining it
  float aabs;  // running max of |trace|
  float vabs;  // signed value at the first index attaining aabs
  int iabs;
};
__device__ __forceinline__ Stat stat_init() { return Stat{-INFINITY, -1, -1.f, 0.f, 0x7fffffff}; }
__device__ __forceinline__ void stat_push(Stat& s, float v, int t) {
  if (s.amax < 0 || v > s.vmax) { s.vmax = v; s.amax = t; }  // first occurrence wins ties (torch.argmax)
  const float a = fabsf(v);
  if (a > s.aabs) { s.aabs = a; s.vabs = v; s.iabs = t; }
}
__device__ __forceinline__ Stat stat_merge(const Stat& a, const Stat& b) {
  Stat r = a;
  if (b.amax >= 0 && (r.amax < 0 || b.vmax > r.vmax || (b.vmax == r.vmax && b.amax < r.amax))) {
    r.vmax = b.vmax; r.amax = b.amax;
  }
  if (b.aabs > r.aabs || (b.aabs == r.aabs && b.iabs < r.iabs)) { r.aabs = b.aabs; r.vabs = b.vabs; r.iabs = b.iabs; }
  return r;
}
__device__ __forceinline__ Stat stat_shfl_xor(const Stat& s, int m) {
  Stat o;
  o.vmax = __shfl_xor_sync(0xffffffffu, s.vmax, m);
  o.amax = __shfl_xor_sync(0xffffffffu, s.amax, m);
  o.aabs = __shfl_xor_sync(0xffffffffu, s.aabs, m);
  o.vabs = __shfl_xor_sync(0xffffffffu, s.vabs, m);
  o.iabs = __shfl_xor_sync(0xffffffffu, s.iabs, m);
  return o;
}
__device__ __forceinline__ Stat block_stat_reduce(Stat s, Stat* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s = stat_merge(s, stat_shfl_xor(s, o));
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) scratch[w] = s;
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    s = lane < nw ? scratch[lane] : stat_init();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = stat_merge(s, stat_shfl_xor(s, o));
  }
  return s;  // valid on warp 0
}
__device__ __forceinline__ void stat_store(const Stat& s, int64_t file, float* vmax, int32_t* amax, float* vabs) {
  vmax[file] = s.vmax;
  amax[file] = s.amax;
  vabs[file] = s.amax < 0 ? __int_as_float(0x7fc00000) : s.vabs;  // NaN marks an empty (fully trimmed) file
}

// One CTA per file; thread t reads acts[file, t, feature] (one 32-byte sector per frame as stored).
template <typename T>
__global__ void __launch_bounds__(256) search_dense_kernel(const T* __restrict__ acts,
                                                           const int32_t* __restrict__ n_frames, int64_t Tn, int64_t F,
                                                           int64_t feature, float* __restrict__ vmax,
                                                           int32_t* __restrict__ amax, float* __restrict__ vabs,
                                                           float* __restrict__ trace) {
  __shared__ Stat scratch[8];
  const int64_t file = blockIdx.x;
  const int nf = min(static_cast<int64_t>(n_frames[file]), Tn);
  const T* col = acts + file * Tn * F + feature;
  Stat s = stat_init();
  const int limit = trace ? static_cast<int>(Tn) : nf;
  const int bd = blockDim.x;
  // 4 independent strided loads in flight per thread (each pulls one 32-byte sector)
  for (int t0 = threadIdx.x; t0 < limit; t0 += 4 * bd) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u * bd;
      v[u] = t < limit ? static_cast<float>(col[static_cast<int64_t>(t) * F]) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u * bd;
      if (t < limit) {
        if (trace) trace[file * Tn + t] = v[u];
        if (t < nf) stat_push(s, v[u], t);
      }
    }
  }
  s = block_stat_reduce(s, scratch);
  if (threadIdx.x == 0) stat_store(s, file, vmax, amax, vabs);
}

// One CTA per file, one warp per frame: lanes compare k slot indices (coalesced), ballot finds the first match.
template <typename IT>
__global__ void __launch_bounds__(256) search_indexed_kernel(const float* __restrict__ vals,
                                                             const IT* __restrict__ idx,
                                                             const int32_t* __restrict__ n_frames, int64_t Tn,
                                                             int64_t k, int64_t feature, float* __restrict__ vmax,
                                                             int32_t* __restrict__ amax, float* __restrict__ vabs,
                                                             float* __restrict__ trace) {
  __shared__ Stat scratch[8];
  const int64_t file = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int nf = min(static_cast<int64_t>(n_frames[file]), Tn);
  const int limit = trace ? static_cast<int>(Tn) : nf;
  Stat s = stat_init();
  if (k <= 32) {
    // common case (k == 32): 4 frames per warp iteration, 4 independent coalesced index loads in flight per lane
    for (int t0 = w; t0 < limit; t0 += 4 * nw) {
      IT id[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int t = t0 + u * nw;
        id[u] = (t < limit && lane < k) ? idx[(file * Tn + t) * k + lane] : static_cast<IT>(-1);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int t = t0 + u * nw;
        if (t >= limit) break;  // warp-uniform
        const uint32_t m = __ballot_sync(0xffffffffu, static_cast<int64_t>(id[u]) == feature);
        float v = 0.f;  // feature absent from the frame's top-k -> 0 (utils/activations.py:48-56)
        if (m) {
          const int src = __ffs(m) - 1;  // first matching slot, as `.nonzero()` on a unique match
          float mine = 0.f;
          if (lane == src) mine = vals[(file * Tn + t) * k + src];
          v = __shfl_sync(0xffffffffu, mine, src);
        }
        if (lane == 0) {
          if (trace) trace[file * Tn + t] = v;
          if (t < nf) stat_push(s, v, t);
        }
      }
    }
  } else {
    for (int t = w; t < limit; t += nw) {
      const int64_t base = (file * Tn + t) * k;
      float v = 0.f;
      for (int64_t j0 = 0; j0 < k; j0 += 32) {
        const int64_t j = j0 + lane;
        const bool hit = j < k && static_cast<int64_t>(idx[base + j]) == feature;
        const uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (m) {
          const int src = __ffs(m) - 1;
          float mine = 0.f;
          if (lane == src) mine = vals[base + j0 + src];
          v = __shfl_sync(0xffffffffu, mine, src);
          break;
        }
      }
      if (lane == 0) {
        if (trace) trace[file * Tn + t] = v;
        if (t < nf) stat_push(s, v, t);
      }
    }
  }
  s = block_stat_reduce(s, scratch);
  if (threadIdx.x == 0) stat_store(s, file, vmax, amax, vabs);
}

// All-feature table: the per-file statistics of EVERY feature column in one pass over the dense store
// (N_files * T * F * sizeof = 23 GB at C5, 3.5 ms at the HBM rate), after which a query is a column
// gather of three [N_files] vectors and the ranking kernel -- no scan at all.  One CTA per file; thread (c, g) owns
// the feature quad c (one 16-byte load per frame: a frame row is read with fully coalesced loads) and the frames
// g, g + G, ...; four loads in flight per thread; the G frame groups meet in shared memory.
template <typename T>
__device__ __forceinline__ void load_quad(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load_quad<float>(const float* p, float (&v)[4]) {
  const float4 t = __ldcs(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load_quad<__half>(const __half* p, float (&v)[4]) {
  const uint2 raw = __ldcs(reinterpret_cast<const uint2*>(p));
  const __half2 a = *reinterpret_cast<const __half2*>(&raw.x), b = *reinterpret_cast<const __half2*>(&raw.y);
  const float2 fa = __half22float2(a), fb = __half22float2(b);
  v[0] = fa.x; v[1] = fa.y; v[2] = fb.x; v[3] = fb.y;
}

template <typename T>
__global__ void __launch_bounds__(256) search_table_kernel(const T* __restrict__ acts, const int32_t* __restrict__ n_frames,
                                                           int64_t Tn, int F, int G, float* __restrict__ vmax_tab,
                                                           int32_t* __restrict__ amax_tab, float* __restrict__ vabs_tab) {
  extern __shared__ unsigned char table_smem[];
  Stat* merge = reinterpret_cast<Stat*>(table_smem);  // [G - 1][cols of this pass * 4]
  const int64_t file = blockIdx.x;
  const int nf = min(static_cast<int64_t>(n_frames[file]), Tn);
  const int F4 = F >> 2;
  const int per_pass = blockDim.x / G;  // feature quads handled per pass
  const int g = threadIdx.x / per_pass, cq = threadIdx.x - g * per_pass;
  const T* base = acts + file * Tn * F;
  for (int q0 = 0; q0 < F4; q0 += per_pass) {
    const int q = q0 + cq;
    const bool live = g < G && q < F4;
    Stat s[4] = {stat_init(), stat_init(), stat_init(), stat_init()};
    if (live) {
      const T* col = base + 4 * q;
      int t = g;
      for (; t + 3 * G < nf; t += 4 * G) {
        float v[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) load_quad<T>(col + static_cast<int64_t>(t + u * G) * F, v[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int e = 0; e < 4; ++e) stat_push(s[e], v[u][e], t + u * G);
      }
      for (; t < nf; t += G) {
        float v[4];
        load_quad<T>(col + static_cast<int64_t>(t) * F, v);
#pragma unroll
        for (int e = 0; e < 4; ++e) stat_push(s[e], v[e], t);
      }
    }
    if (G > 1) {
      __syncthreads();
      if (live && g > 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) merge[(static_cast<size_t>(g - 1) * per_pass + cq) * 4 + e] = s[e];
      }
      __syncthreads();
      if (live && g == 0) {
        for (int gg = 1; gg < G; ++gg)
#pragma unroll
          for (int e = 0; e < 4; ++e) s[e] = stat_merge(s[e], merge[(static_cast<size_t>(gg - 1) * per_pass + cq) * 4 + e]);
      }
    }
    if (live && g == 0) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int64_t o = file * F + 4 * q + e;
        vmax_tab[o] = s[e].vmax;
        amax_tab[o] = s[e].amax;
        vabs_tab[o] = s[e].amax < 0 ? __int_as_float(0x7fc00000) : s[e].vabs;
      }
    }
  }
}

// Indexed store, k == 32 slots per frame, int32 indices (narrowed from the int64 of the .npy at load time: half the
// scan bytes).  One CTA per file; a warp reads FOUR frames per load instruction (lane l: 16 bytes = slots 4(l%8)..+3
// of frame l/8, 512 contiguous bytes per warp), four such loads in flight.  A lane that finds the feature fetches
// the value and folds it into its own running statistics -- no shuffle on the hit path; frames without the
// feature contribute the value 0 (utils/activations.py:48-56), of which only the FIRST one can matter (arg-max
// ties go to the lower frame), so a per-lane minimum frame index stands for all of them.
__global__ void __launch_bounds__(256) search_indexed32_kernel(const float* __restrict__ vals,
                                                               const int32_t* __restrict__ idx,
                                                               const int32_t* __restrict__ n_frames, int64_t Tn,
                                                               int feature, float* __restrict__ vmax,
                                                               int32_t* __restrict__ amax, float* __restrict__ vabs) {
  __shared__ Stat scratch[8];
  const int64_t file = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int nf = min(static_cast<int64_t>(n_frames[file]), Tn);
  const int sub = lane >> 3, part = lane & 7;  // frame within the group of four, 4-slot part of its row
  const int32_t* ibase = idx + file * Tn * 32;
  const float* vbase = vals + file * Tn * 32;
  Stat s = stat_init();
  int zmin = 0x7fffffff;
  for (int t0 = w * 16; t0 < nf; t0 += nw * 16) {
    int4 id[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u * 4 + sub;
      id[u] = t < nf ? __ldcs(reinterpret_cast<const int4*>(ibase + static_cast<int64_t>(t) * 32 + part * 4))
                     : make_int4(-1, -1, -1, -1);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u * 4 + sub;
      int pos = -1;
      if (id[u].w == feature) pos = 3;
      if (id[u].z == feature) pos = 2;
      if (id[u].y == feature) pos = 1;
      if (id[u].x == feature) pos = 0;
      const uint32_t m = __ballot_sync(0xffffffffu, pos >= 0);
      const uint32_t mine = (m >> (sub * 8)) & 0xffu;  // hits among the 8 lanes of my frame
      if (t < nf) {
        if (mine == 0u) {
          if (part == 0) zmin = min(zmin, t);
        } else if (pos >= 0 && (mine & ((1u << part) - 1u)) == 0u) {  // first matching slot of the frame
          stat_push(s, __ldg(vbase + static_cast<int64_t>(t) * 32 + part * 4 + pos), t);
        }
      }
    }
  }
  if (zmin != 0x7fffffff) {
    Stat z = stat_init();
    stat_push(z, 0.f, zmin);
    s = stat_merge(s, z);
  }
  s = block_stat_reduce(s, scratch);
  if (threadIdx.x == 0) stat_store(s, file, vmax, amax, vabs);
}

// Single CTA: n_top rounds of "largest key strictly after the previous pick" in (key desc, file asc) order.
__global__ void __launch_bounds__(1024) search_topn_kernel(const float* __restrict__ vmax,
                                                           const float* __restrict__ vabs, int n_files, int absolute,
                                                           int use_min, double min_val, int use_max, double max_val,
                                                           int n_top, int32_t* __restrict__ out_files,
                                                           int32_t* __restrict__ out_count) {
  __shared__ float sk[32];
  __shared__ int si[32];
  __shared__ float last_key_s;
  __shared__ int last_idx_s;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float last_key = INFINITY;
  int last_idx = -1;
  int found = 0;
  for (int r = 0; r < n_top; ++r) {
    float bk = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n_files; i += blockDim.x) {
      const float signed_v = absolute ? vabs[i] : vmax[i];
      if (!absolute && !(vmax[i] > -INFINITY)) continue;  // empty (n_frames == 0) files never rank
      if (use_max && (double)signed_v > max_val) continue;        // filter_activation, utils/activations.py:88-93
      if (use_min && (double)signed_v < min_val) continue;
      const float key = absolute ? fabsf(signed_v) : signed_v;
      if (!(key == key)) continue;
      // strictly after the previous pick in (key desc, idx asc) order
      const bool after = key < last_key || (key == last_key && i > last_idx);
      if (!after) continue;
      if (key > bk || (key == bk && i < bi)) { bk = key; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ok = __shfl_xor_sync(0xffffffffu, bk, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ok > bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
    }
    if (lane == 0) { sk[w] = bk; si[w] = bi; }
    __syncthreads();
    if (w == 0) {
      bk = sk[lane]; bi = si[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ok = __shfl_xor_sync(0xffffffffu, bk, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ok > bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
      }
      if (lane == 0) { last_key_s = bk; last_idx_s = bi; }
    }
    __syncthreads();
    last_key = last_key_s;
    last_idx = last_idx_s;
    if (last_idx == 0x7fffffff) break;  // no more candidates (uniform)
    if (threadIdx.x == 0) out_files[r] = last_idx;
    ++found;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    for (int r = found; r < n_top; ++r) out_files[r] = -1;
    *out_count = found;
  }
}

}  // namespace freud

using namespace freud;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int freud_search_dense(const void* acts, int acts_is_fp16, const int32_t* n_frames, int64_t n_files,
                                  int64_t T, int64_t F, int64_t feature, float* vmax, int32_t* amax, float* vabs,
                                  float* trace, void* stream) {
  FREUD_REQUIRE(n_files > 0 && T > 0 && F > 0 && feature >= 0 && feature < F, "search_dense: bad shape or feature");
  if (acts_is_fp16)
    search_dense_kernel<__half><<<(unsigned)n_files, 256, 0, STREAM>>>(static_cast<const __half*>(acts), n_frames, T, F,
                                                                      feature, vmax, amax, vabs, trace);
  else
    search_dense_kernel<float><<<(unsigned)n_files, 256, 0, STREAM>>>(static_cast<const float*>(acts), n_frames, T, F,
                                                                     feature, vmax, amax, vabs, trace);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_search_indexed(const float* vals, const void* idx, int idx_is_int64, const int32_t* n_frames,
                                    int64_t n_files, int64_t T, int64_t k, int64_t feature, float* vmax,
                                    int32_t* amax, float* vabs, float* trace, void* stream) {
  FREUD_REQUIRE(n_files > 0 && T > 0 && k > 0 && feature >= 0, "search_indexed: bad shape or feature");
  if (!idx_is_int64 && k == 32 && trace == nullptr && feature < (1ll << 31) &&
      (reinterpret_cast<uintptr_t>(idx) & 15) == 0) {
    search_indexed32_kernel<<<(unsigned)n_files, 256, 0, STREAM>>>(vals, static_cast<const int32_t*>(idx), n_frames, T,
                                                                  (int)feature, vmax, amax, vabs);
    FREUD_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  if (idx_is_int64)
    search_indexed_kernel<int64_t><<<(unsigned)n_files, 256, 0, STREAM>>>(
        vals, static_cast<const int64_t*>(idx), n_frames, T, k, feature, vmax, amax, vabs, trace);
  else
    search_indexed_kernel<int32_t><<<(unsigned)n_files, 256, 0, STREAM>>>(
        vals, static_cast<const int32_t*>(idx), n_frames, T, k, feature, vmax, amax, vabs, trace);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_search_topn(const float* vmax, const float* vabs, int64_t n_files, int absolute, int use_min,
                                 double min_val, int use_max, double max_val, int64_t n_top, int32_t* out_files,
                                 int32_t* out_count, void* stream) {
  FREUD_REQUIRE(n_files > 0 && n_files < (1ll << 31) && n_top > 0, "search_topn: bad sizes");
  search_topn_kernel<<<1, 1024, 0, STREAM>>>(vmax, vabs, (int)n_files, absolute, use_min, min_val, use_max, max_val,
                                             (int)n_top, out_files, out_count);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_search_table_dense(const void* acts, int acts_is_fp16, const int32_t* n_frames, int64_t n_files,
                                        int64_t T, int64_t F, float* vmax_tab, int32_t* amax_tab, float* vabs_tab,
                                        void* stream) {
  FREUD_REQUIRE(n_files > 0 && T > 0 && F > 0 && F % 4 == 0, "search_table_dense needs F % 4 == 0");
  FREUD_REQUIRE((reinterpret_cast<uintptr_t>(acts) & 15) == 0, "activations must be 16-byte aligned");
  const int F4 = static_cast<int>(F / 4);
  int G = 1, threads = 256;
  if (F4 <= 128) {
    G = 256 / F4;
    if (G > 8) G = 8;
    threads = F4 * G;
  }
  threads = (threads + 31) / 32 * 32;
  const int per_pass = threads / G;
  const size_t smem = G > 1 ? static_cast<size_t>(G - 1) * per_pass * 4 * sizeof(Stat) : 0;
  if (acts_is_fp16)
    search_table_kernel<__half><<<(unsigned)n_files, threads, smem, STREAM>>>(static_cast<const __half*>(acts), n_frames,
                                                                             T, (int)F, G, vmax_tab, amax_tab, vabs_tab);
  else
    search_table_kernel<float><<<(unsigned)n_files, threads, smem, STREAM>>>(static_cast<const float*>(acts), n_frames,
                                                                            T, (int)F, G, vmax_tab, amax_tab, vabs_tab);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}
