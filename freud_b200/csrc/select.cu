// Exact per-row top-k over a materialised fp32 matrix (torch.topk(k, sorted=False), topkautoencoder.py:81,121,135)
// for arbitrary k (AuxK uses k = d/2, multi-TopK 4k) and an optional column mask (dead latents, :118).
// One CTA per row: 4 x 8-bit radix-select passes find the k-th largest key exactly, a final pass emits the
// winners in index order (ties at the threshold resolved towards the lower index), then the k winners are
// sorted (value desc, index asc) in shared memory.  Streams the row from L2/HBM; no tensor cores.
#include "device_utils.cuh"
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
