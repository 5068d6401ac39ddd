// Validation feature statistics (reference: topk_feature_extraction and the L1 per-file abs-max
// in src/scripts/train_sae.py:70-118,175-178).  The reference materialises a [T,k,n] boolean mask per file; here the
// (frame, slot) entries are scattered with an integer atomic max on the bit pattern of |activation| (non-negative
// floats order like unsigned integers), so the cost is one pass over the file's encoding.
#include "device_utils.cuh"
#include "host_common.h"
#include "../../include/freud_b200.h"

namespace freud {

template <typename IT>
__global__ void __launch_bounds__(256) feature_absmax_kernel(const float* __restrict__ acts, const IT* __restrict__ idx,
                                                             int64_t count, unsigned int* __restrict__ out, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t f = static_cast<int64_t>(idx[i]);
    if (f >= 0 && f < n) atomicMax(out + f, __float_as_uint(fabsf(acts[i])));
  }
}

// grid: (ceil(n/256), row slabs); coalesced across columns, one atomic per (column, slab)
__global__ void __launch_bounds__(256) col_absmax_kernel(const float* __restrict__ x, int64_t rows, int n, int slab,
                                                         unsigned int* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * slab, r1 = min(rows, r0 + slab);
  float m = 0.f;
  for (int64_t r = r0; r < r1; ++r) m = fmaxf(m, fabsf(x[r * n + j]));
  atomicMax(out + j, __float_as_uint(m));
}

}  // namespace freud

using namespace freud;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int freud_feature_absmax(const float* acts, const void* idx, int idx_is_int64, int64_t count, float* out,
                                    int64_t n, void* stream) {
  FREUD_REQUIRE(count > 0 && n > 0, "feature_absmax: empty input");
  FREUD_CHECK_CUDA(cudaMemsetAsync(out, 0, n * sizeof(float), STREAM));
  int64_t grid = (count + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (grid > cap) grid = cap;
  if (idx_is_int64)
    feature_absmax_kernel<int64_t><<<(unsigned)grid, 256, 0, STREAM>>>(acts, static_cast<const int64_t*>(idx), count,
                                                                       reinterpret_cast<unsigned int*>(out), n);
  else
    feature_absmax_kernel<int32_t><<<(unsigned)grid, 256, 0, STREAM>>>(acts, static_cast<const int32_t*>(idx), count,
                                                                       reinterpret_cast<unsigned int*>(out), n);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_col_absmax(const float* x, int64_t rows, int64_t n, float* out, void* stream) {
  FREUD_REQUIRE(rows > 0 && n > 0, "col_absmax: empty input");
  FREUD_CHECK_CUDA(cudaMemsetAsync(out, 0, n * sizeof(float), STREAM));
  const int slab = 128;
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)((rows + slab - 1) / slab));
  col_absmax_kernel<<<grid, 256, 0, STREAM>>>(x, rows, (int)n, slab, reinterpret_cast<unsigned int*>(out));
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}
