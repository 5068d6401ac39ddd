// L2 -> SM read-bandwidth probe (the ceiling DESIGN.md section 5 grades the gather kernels against).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/l2_probe scripts/l2_probe.cu && /tmp/l2_probe
// Every SM streams 16-byte loads out of a table that stays resident in the 126 MB L2 (sizes 19 / 38 / 75 MB: the bf16
// W_dec of C2-like / C3 / the C3 residual), bypassing L1 (ld.global.cg), with U independent loads in flight per thread:
//   linear : consecutive 512-byte warp rows                      -- the upper bound of the L2 read path
//   gather : random 1536-byte rows (a bf16 d = 768 decoder row),  -- the access pattern of freud_topk_decode / dacts /
//            one row per warp iteration, 3 loads per lane            sparse_grads
// and, for scale, the same linear kernel over a 4 GB buffer (HBM).  Prints one JSON document.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint4 ldcg(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

template <int U>
__global__ void __launch_bounds__(256) linear_read(const uint4* __restrict__ buf, size_t n16, int reps, uint32_t* sink) {
  uint32_t acc = 0;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (int r = 0; r < reps; ++r) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride * U) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = (i + u * stride < n16) ? ldcg(buf + i + u * stride) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}

// one warp per gathered row of 1536 bytes (96 uint4: 3 per lane), R rows in flight per warp
template <int R>
__global__ void __launch_bounds__(256) gather_read(const uint4* __restrict__ buf, const int* __restrict__ rows,
                                                   size_t n_gather, uint32_t* sink) {
  uint32_t acc = 0;
  const int lane = threadIdx.x & 31;
  const size_t warp = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5;
  for (size_t g = warp * R; g + R <= n_gather; g += n_warps * R) {
    uint4 v[R][3];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const uint4* row = buf + static_cast<size_t>(rows[g + r]) * 96;
#pragma unroll
      for (int i = 0; i < 3; ++i) v[r][i] = ldcg(row + i * 32 + lane);
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int i = 0; i < 3; ++i) acc ^= v[r][i].x ^ v[r][i].y ^ v[r][i].z ^ v[r][i].w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

template <typename F>
static float time_ms(F launch, int iters) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) launch();
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) launch();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / iters;
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  uint32_t* sink;
  CK(cudaMalloc(&sink, 4));
  printf("{\n \"sm_count\": %d", sms);
  const size_t sizes_mb[3] = {19, 38, 75};
  for (int s = 0; s < 3; ++s) {
    const size_t bytes = sizes_mb[s] << 20, n16 = bytes / 16;
    uint4* buf;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 1, bytes));
    const int reps = 16;
    float best = 1e30f;
    int best_cfg = 0;
    for (int cfg = 0; cfg < 3; ++cfg) {
      const int ctas = sms * (cfg == 0 ? 4 : cfg == 1 ? 8 : 8);
      float ms;
      if (cfg < 2) ms = time_ms([&] { linear_read<8><<<ctas, 256>>>(buf, n16, reps, sink); }, 5);
      else ms = time_ms([&] { linear_read<16><<<ctas, 256>>>(buf, n16, reps, sink); }, 5);
      if (ms < best) { best = ms; best_cfg = cfg; }
    }
    printf(",\n \"l2_linear_%zuMB_GBps\": %.1f, \"l2_linear_%zuMB_cfg\": %d", sizes_mb[s], bytes * (double)reps / best / 1e6,
           sizes_mb[s], best_cfg);
    // gather of 1536-byte rows: 48000 tokens x 32 rows = 1 536 000 row reads (2.36 GB), as one C3 decode launch
    const size_t n_rows = bytes / 1536, n_gather = 1536000;
    std::vector<int> h(n_gather);
    uint64_t st = 88172645463325252ull;
    for (size_t i = 0; i < n_gather; ++i) {
      st ^= st << 13; st ^= st >> 7; st ^= st << 17;
      h[i] = static_cast<int>(st % n_rows);
    }
    int* rows;
    CK(cudaMalloc(&rows, n_gather * 4));
    CK(cudaMemcpy(rows, h.data(), n_gather * 4, cudaMemcpyHostToDevice));
    float gbest = 1e30f;
    for (int cfg = 0; cfg < 3; ++cfg) {
      const int ctas = sms * 8;
      float ms;
      if (cfg == 0) ms = time_ms([&] { gather_read<4><<<ctas, 256>>>(buf, rows, n_gather, sink); }, 5);
      else if (cfg == 1) ms = time_ms([&] { gather_read<8><<<ctas, 256>>>(buf, rows, n_gather, sink); }, 5);
      else ms = time_ms([&] { gather_read<8><<<sms * 4, 256>>>(buf, rows, n_gather, sink); }, 5);
      if (ms < gbest) gbest = ms;
    }
    printf(",\n \"l2_gather1536_%zuMB_GBps\": %.1f", sizes_mb[s], n_gather * 1536.0 / gbest / 1e6);
    CK(cudaFree(rows));
    CK(cudaFree(buf));
  }
  {
    const size_t bytes = 4ull << 30, n16 = bytes / 16;
    uint4* buf;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 1, bytes));
    const float ms = time_ms([&] { linear_read<8><<<sms * 8, 256>>>(buf, n16, 1, sink); }, 5);
    printf(",\n \"hbm_linear_4GB_GBps\": %.1f", bytes / (double)ms / 1e6);
    CK(cudaFree(buf));
  }
  printf(",\n \"note\": \"bytes read / kernel time, best of the listed launch shapes; ld.global.cg (L1 bypass), CUDA events\"\n}\n");
  return 0;
}
