// Exact per-row top-k over a materialised fp32 matrix (torch.topk(k, sorted=False), topkautoencoder.py:81,121,135)
// for arbitrary k (AuxK uses k = d/2, multi-TopK 4k) and an optional column mask (dead latents, :118).
// One CTA per row: 4 x 8-bit radix-select passes find the k-th largest key exactly, a final pass emits the
// winners in index order (ties at the threshold resolved towards the lower index), then the k winners are
// sorted (value desc, index asc) in shared memory.  Streams the row from L2/HBM; no tensor cores.
#include "device_utils.cuh"
#include <cuda_bf16.h>
#include "host_common.h"
#include "../../include/freud_b200.h"

namespace freud {

__device__ __forceinline__ uint32_t float_key(float v) {
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending unsigned order == ascending float order
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

constexpr int kSelThreads = 512;
constexpr int kSelMaxK = 2048;

__global__ void __launch_bounds__(kSelThreads) row_topk_kernel(const float* __restrict__ latents,
                                                               const uint8_t* __restrict__ col_mask,
                                                               float* __restrict__ vals, int32_t* __restrict__ idx,
                                                               int n, int k) {
  __shared__ int hist[256];
  __shared__ uint32_t s_prefix;
  __shared__ int s_remaining;
  __shared__ int s_gt_base, s_eq_base;
  __shared__ int warp_gt[kSelThreads / 32], warp_eq[kSelThreads / 32];
  __shared__ uint64_t out_keys[kSelMaxK];
  const int64_t row = blockIdx.x;
  const float* __restrict__ x = latents + row * n;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

  auto key_at = [&](int j) -> uint32_t {
    if (col_mask && col_mask[j] == 0) return float_key(-INFINITY);  // torch.where(mask, x, -inf)
    return float_key(x[j]);
  };

  if (tid == 0) {
    s_prefix = 0;
    s_remaining = k;
  }
  __syncthreads();
  // ---- radix select, most significant digit first
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    const uint32_t prefix = s_prefix;
    const uint32_t hi_mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int i = tid; i < 256; i += kSelThreads) hist[i] = 0;
    __syncthreads();
    for (int j = tid; j < n; j += kSelThreads) {
      const uint32_t key = key_at(j);
      if ((key & hi_mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xff], 1);
    }
    __syncthreads();
    if (tid == 0) {
      int rem = s_remaining;
      int digit = 255;
      for (; digit > 0; --digit) {
        if (hist[digit] >= rem) break;
        rem -= hist[digit];
      }
      s_prefix = prefix | (static_cast<uint32_t>(digit) << shift);
      s_remaining = rem;  // how many keys equal to the final prefix must still be taken
    }
    __syncthreads();
  }
  const uint32_t kth = s_prefix;
  const int need_eq = s_remaining;
  const int n_gt = k - need_eq;
  if (tid == 0) {
    s_gt_base = 0;
    s_eq_base = 0;
  }
  __syncthreads();
  // ---- emit: keys > kth go to slots [0, n_gt) in index order, the first need_eq keys == kth fill the rest
  for (int base = 0; base < n; base += kSelThreads) {
    const int j = base + tid;
    uint32_t key = 0;
    bool gt = false, eq = false;
    if (j < n) {
      key = key_at(j);
      gt = key > kth;
      eq = key == kth;
    }
    const uint32_t mg = __ballot_sync(0xffffffffu, gt), me = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) {
      warp_gt[w] = __popc(mg);
      warp_eq[w] = __popc(me);
    }
    __syncthreads();
    int og = s_gt_base, oe = s_eq_base;
    for (int ww = 0; ww < w; ++ww) {
      og += warp_gt[ww];
      oe += warp_eq[ww];
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
    if (gt) {
      const int pos = og + __popc(mg & lt_mask);
      out_keys[pos] = (static_cast<uint64_t>(key) << 32) | static_cast<uint32_t>(~j);
    } else if (eq) {
      const int r = oe + __popc(me & lt_mask);
      if (r < need_eq) out_keys[n_gt + r] = (static_cast<uint64_t>(key) << 32) | static_cast<uint32_t>(~j);
    }
    __syncthreads();
    if (tid == 0) {
      int tg = 0, te = 0;
      for (int ww = 0; ww < kSelThreads / 32; ++ww) {
        tg += warp_gt[ww];
        te += warp_eq[ww];
      }
      s_gt_base += tg;
      s_eq_base += te;
    }
    __syncthreads();
  }
  // ---- sort the k winners descending by (key, ~index): bitonic in shared memory
  int m = 2;
  while (m < k) m <<= 1;
  for (int i = k + tid; i < m; i += kSelThreads) out_keys[i] = 0ull;
  __syncthreads();
  for (int kk = 2; kk <= m; kk <<= 1) {
    for (int jj = kk >> 1; jj > 0; jj >>= 1) {
      for (int i = tid; i < m; i += kSelThreads) {
        const int l = i ^ jj;
        if (l > i) {
          const uint64_t a = out_keys[i], b = out_keys[l];
          const bool desc = (i & kk) == 0;
          if ((a < b) == desc) {
            out_keys[i] = b;
            out_keys[l] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < k; i += kSelThreads) {
    const uint64_t kv = out_keys[i];
    vals[row * k + i] = key_float(static_cast<uint32_t>(kv >> 32));
    idx[row * k + i] = static_cast<int32_t>(~static_cast<uint32_t>(kv));
  }
}

// Same exact selection, emitted as a DENSE masked row: out[row, j] = x[j] if j is among the row's top-k (ties at the
// k-th value resolved towards the lower index) else 0, written as bf16 with row pitch `ld` (columns [n, ld) zeroed).
// Used by the AuxK branch on the compacted dead-latent subset, where the selection feeds dense GEMMs.
__global__ void __launch_bounds__(kSelThreads) row_topk_mask_kernel(const float* __restrict__ latents,
                                                                    __nv_bfloat16* __restrict__ out, int n, int k,
                                                                    int ld) {
  __shared__ int hist[256];
  __shared__ uint32_t s_prefix;
  __shared__ int s_remaining;
  __shared__ int s_eq_base;
  __shared__ int warp_eq[kSelThreads / 32];
  const int64_t row = blockIdx.x;
  const float* __restrict__ x = latents + row * n;
  __nv_bfloat16* __restrict__ o = out + row * ld;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) {
    s_prefix = 0;
    s_remaining = k;
    s_eq_base = 0;
  }
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    const uint32_t prefix = s_prefix;
    const uint32_t hi_mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int i = tid; i < 256; i += kSelThreads) hist[i] = 0;
    __syncthreads();
    for (int j = tid; j < n; j += kSelThreads) {
      const uint32_t key = float_key(x[j]);
      if ((key & hi_mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xff], 1);
    }
    __syncthreads();
    if (tid == 0) {
      int rem = s_remaining;
      int digit = 255;
      for (; digit > 0; --digit) {
        if (hist[digit] >= rem) break;
        rem -= hist[digit];
      }
      s_prefix = prefix | (static_cast<uint32_t>(digit) << shift);
      s_remaining = rem;
    }
    __syncthreads();
  }
  const uint32_t kth = s_prefix;
  const int need_eq = s_remaining;
  for (int base = 0; base < ld; base += kSelThreads) {
    const int j = base + tid;
    float v = 0.f;
    bool gt = false, eq = false;
    if (j < n) {
      v = x[j];
      const uint32_t key = float_key(v);
      gt = key > kth;
      eq = key == kth;
    }
    const uint32_t me = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) warp_eq[w] = __popc(me);
    __syncthreads();
    int oe = s_eq_base;
    for (int ww = 0; ww < w; ++ww) oe += warp_eq[ww];
    const bool take = gt || (eq && oe + __popc(me & ((1u << lane) - 1u)) < need_eq);
    if (j < ld) o[j] = __float2bfloat16_rn(take ? v : 0.f);
    __syncthreads();
    if (tid == 0) {
      int te = 0;
      for (int ww = 0; ww < kSelThreads / 32; ++ww) te += warp_eq[ww];
      s_eq_base += te;
    }
    __syncthreads();
  }
}

// Warp-per-row version for NON-NEGATIVE rows (post-ReLU pre-activations) that fit in registers (n <= 64 * PAIRS):
// lane l holds the column pairs {2l + 64i, 2l + 64i + 1}, read once with coalesced 8-byte loads.  Only the strictly
// positive values compete (a zero is written as zero whether selected or not, and with fewer than k positives all of
// them are kept), so their raw bit patterns order like the values.  The k-th largest is bracketed by bisection on the
// bit patterns, starting from the row's own [min, max] and counting with register compares + one warp reduction per
// step (no shared-memory atomics: the histogram versions serialised on the few exponent bins a row occupies), until
// at most 32 values remain in the bracket; those are gathered one per lane and ranked with shuffles.  The masked
// row goes out as coalesced bf16x2 stores.  Ties at the k-th value are resolved towards the lower column index, as
// in row_topk_mask_kernel.
template <int PAIRS>
__global__ void __launch_bounds__(256) row_topk_mask_warp_kernel(const float* __restrict__ latents,
                                                                 __nv_bfloat16* __restrict__ out, int64_t rows, int n,
                                                                 int k, int ld_in, int ld_out) {
  __shared__ uint32_t gather_s[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t* gat = gather_s[w];
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + w; row < rows; row += static_cast<int64_t>(gridDim.x) * 8) {
    const float* __restrict__ x = latents + row * ld_in;
    uint32_t u[PAIRS][2];  // raw bits of the strictly positive values, 0 for everything else
    int npos = 0;
    uint32_t umax = 0u, umin = 0xffffffffu;
    // all loads of the row first (the pitch is even, so a pair that starts inside it lies inside it): written with
    // per-element branches the compiler issued load -> use -> load ..., 40 serial memory round trips per row
    float2 raw[PAIRS];
#pragma unroll
    for (int i = 0; i < PAIRS; ++i) {
      const int c = 2 * lane + 64 * i;
      raw[i] = c < ld_in ? __ldcs(reinterpret_cast<const float2*>(x + c)) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < PAIRS; ++i) {
      const int c = 2 * lane + 64 * i;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float tv = (c + e < n) ? (e == 0 ? raw[i].x : raw[i].y) : 0.f;
        const uint32_t b = tv > 0.f ? __float_as_uint(tv) : 0u;
        u[i][e] = b;
        npos += b != 0u;
        umax = max(umax, b);
        umin = min(umin, b != 0u ? b : 0xffffffffu);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      npos += __shfl_xor_sync(0xffffffffu, npos, o);
      umax = max(umax, __shfl_xor_sync(0xffffffffu, umax, o));
      umin = min(umin, __shfl_xor_sync(0xffffffffu, umin, o));
    }
    uint32_t kth = 1u;  // smallest positive pattern: with fewer than k positives every positive value is kept
    int need_eq = 0x7fffffff;
    if (npos >= k) {  // warp-uniform
      // invariants: count(u >= lo) = c_lo >= k,  count(u >= hi) = c_hi < k
      uint32_t lo = umin, hi = umax + 1u;
      int c_lo = npos, c_hi = 0;
      // (an interpolated pivot -- where a count falling linearly over the bracket would cross k -- was measured
      //  slower than plain bisection, 0.56 vs 0.45 ms at N = 48 000, S = 2458: it closes in from one side only and the
      //  loop ends when BOTH sides are within 32 values)
      while (hi - lo > 1u && c_lo - c_hi > 32) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        int c = 0;
#pragma unroll
        for (int i = 0; i < PAIRS; ++i) c += (u[i][0] >= mid) + (u[i][1] >= mid);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (c >= k) {
          lo = mid;
          c_lo = c;
        } else {
          hi = mid;
          c_hi = c;
        }
      }
      if (hi - lo == 1u) {
        kth = lo;
        need_eq = k - c_hi;  // count(u > lo) == count(u >= hi)
      } else {
        // gather the c_lo - c_hi <= 32 values of [lo, hi), one per lane (any order), and rank them
        int mine = 0;
#pragma unroll
        for (int i = 0; i < PAIRS; ++i) mine += (u[i][0] >= lo && u[i][0] < hi) + (u[i][1] >= lo && u[i][1] < hi);
        int base = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v2 = __shfl_up_sync(0xffffffffu, base, o);
          if (lane >= o) base += v2;
        }
        base -= mine;  // exclusive prefix
#pragma unroll
        for (int i = 0; i < PAIRS; ++i) {
#pragma unroll
          for (int e = 0; e < 2; ++e)
            if (u[i][e] >= lo && u[i][e] < hi) gat[base++] = u[i][e];
        }
        __syncwarp();
        const int m = c_lo - c_hi;
        const uint32_t el = lane < m ? gat[lane] : 0u;
        int gt = 0, ge = 0;
        for (int j = 0; j < m; ++j) {
          const uint32_t o2 = __shfl_sync(0xffffffffu, el, j);
          gt += o2 > el;
          ge += o2 >= el;
        }
        const int r = k - c_hi;  // the r-th largest of the gathered values is the k-th largest of the row
        const uint32_t hit = __ballot_sync(0xffffffffu, lane < m && gt < r && r <= ge);
        const int src = __ffs(hit) - 1;  // every hit lane holds the same value
        kth = __shfl_sync(0xffffffffu, el, src);
        need_eq = r - __shfl_sync(0xffffffffu, gt, src);
        __syncwarp();
      }
    }
    int eq_before = 0;
    __nv_bfloat16* __restrict__ o = out + row * ld_out;
#pragma unroll
    for (int i = 0; i < PAIRS; ++i) {
      const int c = 2 * lane + 64 * i;
      const bool g0 = u[i][0] > kth, g1 = u[i][1] > kth;
      const bool e0 = u[i][0] == kth, e1 = u[i][1] == kth;
      const uint32_t b0 = __ballot_sync(0xffffffffu, e0), b1 = __ballot_sync(0xffffffffu, e1);
      const int before = eq_before + __popc(b0 & lt_mask) + __popc(b1 & lt_mask);
      const bool t0 = g0 || (e0 && before < need_eq);
      const bool t1 = g1 || (e1 && before + (e0 ? 1 : 0) < need_eq);
      eq_before += __popc(b0) + __popc(b1);
      if (c < ld_out) {
        const __nv_bfloat162 q =
            __floats2bfloat162_rn(t0 ? __uint_as_float(u[i][0]) : 0.f, t1 ? __uint_as_float(u[i][1]) : 0.f);
        if (c + 1 < ld_out)
          *reinterpret_cast<__nv_bfloat162*>(o + c) = q;
        else
          o[c] = __low2bfloat16(q);
      }
    }
    // columns of the output pitch beyond what this lane's pairs cover
    for (int c = 64 * PAIRS + 2 * lane; c < ld_out; c += 64) {
      o[c] = __float2bfloat16_rn(0.f);
      if (c + 1 < ld_out) o[c + 1] = __float2bfloat16_rn(0.f);
    }
  }
}

// colsum[j] = sum_r x[r, j] for a bf16 matrix [rows, ld] (first n columns; ld % 8 == 0), fp32 accumulation; caller
// zeroes colsum.  A lane owns a 16-byte column chunk (8 columns), the 8 warps of a CTA stride over the rows of the slab
// with four independent loads in flight each, partial sums meet in shared memory and leave as one atomic per column
// and CTA.  (The first version walked the rows serially with 4-byte loads: 42 us for 60 MB at C1.)
__global__ void __launch_bounds__(256) col_sum_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ colsum,
                                                           int64_t rows, int n, int ld, int slab) {
  __shared__ float part[8][32][8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + lane;  // 16-byte chunk of the row
  const bool active = ch * 8 < n;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * slab, r1 = min(rows, r0 + slab);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (active) {
    const __nv_bfloat16* col = x + ch * 8;
    int64_t r = r0 + w;
    for (; r + 24 < r1; r += 32) {
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = __ldcs(reinterpret_cast<const uint4*>(col + (r + 8 * u) * ld));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t wds[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[2 * e] += __uint_as_float(wds[e] << 16);
          acc[2 * e + 1] += __uint_as_float(wds[e] & 0xffff0000u);
        }
      }
    }
    for (; r < r1; r += 8) {
      const uint4 q = __ldcs(reinterpret_cast<const uint4*>(col + r * ld));
      const uint32_t wds[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[2 * e] += __uint_as_float(wds[e] << 16);
        acc[2 * e + 1] += __uint_as_float(wds[e] & 0xffff0000u);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) part[w][lane][e] = acc[e];
  __syncthreads();
  // thread t sums column t of the CTA's 256 columns over the 8 warps
  const int c = threadIdx.x, j = blockIdx.x * 256 + c;
  if (j < n) {
    float s = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) s += part[ww][c >> 3][c & 7];
    atomicAdd(colsum + j, s);
  }
}


// dst[rows_idx[r], :] += src[r, :]   (row scatter-add of the dead-subset gradients into the full matrices)
__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const float* __restrict__ src,
                                                               const int32_t* __restrict__ rows_idx,
                                                               float* __restrict__ dst, int64_t n_rows, int row_elems) {
  const int64_t total = n_rows * row_elems;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / row_elems;
    const int c = static_cast<int>(i - r * row_elems);
    dst[static_cast<int64_t>(rows_idx[r]) * row_elems + c] += src[i];  // rows_idx entries are distinct
  }
}

}  // namespace freud

using namespace freud;

extern "C" int freud_row_topk(const float* latents, const uint8_t* col_mask, float* vals, int32_t* idx, int64_t rows,
                              int64_t n, int64_t k, void* stream) {
  FREUD_REQUIRE(rows > 0 && n > 0 && k > 0 && k <= n, "row_topk needs 0 < k <= n");
  FREUD_REQUIRE(k <= kSelMaxK, "row_topk supports k <= 2048");
  FREUD_REQUIRE(n < (1ll << 31) && rows < (1ll << 31), "row_topk sizes exceed int32");
  row_topk_kernel<<<(unsigned)rows, kSelThreads, 0, static_cast<cudaStream_t>(stream)>>>(latents, col_mask, vals, idx,
                                                                                       (int)n, (int)k);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_row_topk_mask(const float* latents, void* out_bf16, int64_t rows, int64_t n, int64_t k,
                                   int64_t ld_in, int64_t ld, int nonneg, void* stream) {
  FREUD_REQUIRE(rows > 0 && n > 0 && k > 0 && k <= n && ld >= n && ld_in >= n, "row_topk_mask needs 0 < k <= n <= ld");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(out_bf16);
  const bool warp_ok = nonneg && n <= 64 * 64 && ld_in % 2 == 0 && ld % 2 == 0 &&
                       (reinterpret_cast<uintptr_t>(latents) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0;
  if (warp_ok) {
    int64_t grid = (rows + 7) / 8;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 6;
    if (grid > cap) grid = cap;
#define LW(P) \
  row_topk_mask_warp_kernel<P><<<(unsigned)grid, 256, 0, st>>>(latents, out, rows, (int)n, (int)k, (int)ld_in, (int)ld)
    if (n <= 64 * 8) LW(8);
    else if (n <= 64 * 16) LW(16);
    else if (n <= 64 * 24) LW(24);
    else if (n <= 64 * 32) LW(32);
    else if (n <= 64 * 40) LW(40);
    else if (n <= 64 * 48) LW(48);
    else LW(64);
#undef LW
  } else {
    FREUD_REQUIRE(ld_in == n, "the CTA-per-row kernel reads dense rows");
    row_topk_mask_kernel<<<(unsigned)rows, kSelThreads, 0, st>>>(latents, out, (int)n, (int)k, (int)ld);
  }
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_col_sum_bf16(const void* x_bf16, float* colsum, int64_t rows, int64_t n, int64_t ld, void* stream) {
  FREUD_REQUIRE(rows > 0 && n > 0 && ld >= n && ld % 2 == 0, "col_sum_bf16: bad sizes");
  FREUD_CHECK_CUDA(cudaMemsetAsync(colsum, 0, n * sizeof(float), static_cast<cudaStream_t>(stream)));
  FREUD_REQUIRE(ld % 8 == 0, "col_sum_bf16 needs a row pitch of a multiple of 8 elements");
  const int slab = 256;
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)((rows + slab - 1) / slab));
  col_sum_bf16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x_bf16), colsum,
                                                                          rows, (int)n, (int)ld, slab);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_scatter_add_rows(const float* src, const int32_t* rows_idx, float* dst, int64_t n_rows,
                                      int64_t row_elems, void* stream) {
  FREUD_REQUIRE(n_rows > 0 && row_elems > 0, "scatter_add_rows: bad sizes");
  int64_t grid = (n_rows * row_elems + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (grid > cap) grid = cap;
  scatter_add_rows_kernel<<<(unsigned)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, rows_idx, dst, n_rows,
                                                                                        (int)row_elems);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}
