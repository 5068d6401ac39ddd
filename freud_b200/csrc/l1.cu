// L1 (tied-weight) SAE pieces that are not the two forward GEMMs: column normalisation, the fused
// L1 / masked-MSE loss reduction, the ReLU-masked latent gradient with bias reduction and the tied weight
// gradient dW = X^T dz + dxhat^T c.  (Reference: src/models/l1autoencoder.py:29-95; SURVEY.md M2, M3, M3'.)
#include "device_utils.cuh"
#include "host_common.h"
#include "../../include/freud_b200.h"

namespace freud {

// One CTA of 8 warps per group of 32 dictionary columns: thread (w, lane) walks rows w, w+8, ... of column j0 + lane
// (coalesced across the lanes), the 8 row-partials of a column meet in shared memory.  (One thread per column walked
// all d rows twice by itself: 0.09 ms of pure latency for the 384 x 200 matrix of C1, every encode() call.)
__global__ void __launch_bounds__(256) l1_colnorm_kernel(float* __restrict__ W, float* __restrict__ Wt, int d, int n) {
  __shared__ float part[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (j < n) {
    for (int i = w; i < d; i += 8) {
      const float v = W[static_cast<int64_t>(i) * n + j];
      s = fmaf(v, v, s);
    }
  }
  part[w][lane] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) tot += part[k][lane];
  const float inv = 1.f / fmaxf(sqrtf(tot), 1e-12f);  // F.normalize eps
  if (j < n) {
    for (int i = w; i < d; i += 8) {
      const float v = W[static_cast<int64_t>(i) * n + j] * inv;
      W[static_cast<int64_t>(i) * n + j] = v;
      Wt[static_cast<int64_t>(j) * d + i] = v;
    }
  }
}

__global__ void __launch_bounds__(256) l1_loss_kernel(const float* __restrict__ latent,
                                                      const float* __restrict__ x_hat, const float* __restrict__ x,
                                                      float* __restrict__ dxhat, double* __restrict__ acc,
                                                      int64_t n_lat, int64_t n_act) {
  __shared__ double scratch[32];
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  double l1 = 0.0, sse_m = 0.0, cnt = 0.0, sse = 0.0;
  for (int64_t i = i0; i < n_lat; i += stride) l1 += fabsf(latent[i]);
  for (int64_t i = i0; i < n_act; i += stride) {
    const float xv = x[i];
    const float e = x_hat[i] - xv;
    const bool keep = xv != -1.0f;  // mse_loss(..., ignored_index=-1): mask = target == -1
    const float e2 = e * e;
    sse += e2;
    if (keep) {
      sse_m += e2;
      cnt += 1.0;
    }
    if (dxhat) dxhat[i] = keep ? e : 0.f;
  }
  double t;
  t = block_sum(l1, scratch);
  if (threadIdx.x == 0 && t != 0.0) atomicAdd(acc + 0, t);
  t = block_sum(sse_m, scratch);
  if (threadIdx.x == 0 && t != 0.0) atomicAdd(acc + 1, t);
  t = block_sum(cnt, scratch);
  if (threadIdx.x == 0 && t != 0.0) atomicAdd(acc + 2, t);
  t = block_sum(sse, scratch);
  if (threadIdx.x == 0 && t != 0.0) atomicAdd(acc + 3, t);
}

// grid: (ceil(n/256), row_slabs).  dz in place over dc; per-column partial sums -> atomicAdd into db.
__global__ void __launch_bounds__(256) l1_dz_kernel(float* __restrict__ dc, const float* __restrict__ latent,
                                                    const float* __restrict__ scales, float* __restrict__ db,
                                                    int64_t N, int n, int rows_per_slab) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float s_recon = scales[0], s_l1 = scales[1];
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_slab;
  const int64_t r1 = min(N, r0 + rows_per_slab);
  float acc = 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    const int64_t o = r * n + j;
    const float dz = latent[o] > 0.f ? fmaf(s_recon, dc[o], s_l1) : 0.f;
    dc[o] = dz;
    acc += dz;
  }
  atomicAdd(db + j, acc);
}

// dW[i,j] += sum_{t in slab} ( s0*X[t,i]*dz[t,j] + s1*dxhat[t,i]*c[t,j] );  64x64 output tile per CTA,
// split over the token axis (blockIdx.z), 4x4 register micro-tile per thread, 16-token smem stages.
constexpr int kWgTile = 64, kWgK = 16;
__global__ void __launch_bounds__(256) l1_weight_grad_kernel(const float* __restrict__ x,
                                                             const float* __restrict__ dz,
                                                             const float* __restrict__ dxhat,
                                                             const float* __restrict__ latent,
                                                             const float* __restrict__ scales,
                                                             float* __restrict__ dW, int64_t N, int d, int n,
                                                             int64_t tokens_per_slab) {
  __shared__ float As[2][kWgK][kWgTile];  // [term][token][i]  (x, dxhat)
  __shared__ float Bs[2][kWgK][kWgTile];  // [term][token][j]  (dz, c)
  const int i0 = blockIdx.y * kWgTile, j0 = blockIdx.x * kWgTile;
  const int64_t t0 = static_cast<int64_t>(blockIdx.z) * tokens_per_slab;
  const int64_t t1 = min(N, t0 + tokens_per_slab);
  const int ti = (threadIdx.x >> 4) * 4, tj = (threadIdx.x & 15) * 4;
  float acc[2][4][4] = {};
  const int lr = threadIdx.x >> 4;        // 0..15 token row within the stage
  const int lc = (threadIdx.x & 15) * 4;  // 4 consecutive columns
  for (int64_t tb = t0; tb < t1; tb += kWgK) {
    const int64_t t = tb + lr;
    float4 a0 = make_float4(0, 0, 0, 0), a1 = a0, b0 = a0, b1 = a0;
    if (t < t1) {
      if (i0 + lc + 3 < d) {
        a0 = load4(x + t * d + i0 + lc);
        a1 = load4(dxhat + t * d + i0 + lc);
      } else {
        float* pa0 = &a0.x; float* pa1 = &a1.x;
        for (int u = 0; u < 4; ++u)
          if (i0 + lc + u < d) { pa0[u] = x[t * d + i0 + lc + u]; pa1[u] = dxhat[t * d + i0 + lc + u]; }
      }
      if (j0 + lc + 3 < n && (n & 3) == 0) {
        b0 = load4(dz + t * n + j0 + lc);
        b1 = load4(latent + t * n + j0 + lc);
      } else {
        float* pb0 = &b0.x; float* pb1 = &b1.x;
        for (int u = 0; u < 4; ++u)
          if (j0 + lc + u < n) { pb0[u] = dz[t * n + j0 + lc + u]; pb1[u] = latent[t * n + j0 + lc + u]; }
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[0][lr][lc]) = a0;
    *reinterpret_cast<float4*>(&As[1][lr][lc]) = a1;
    *reinterpret_cast<float4*>(&Bs[0][lr][lc]) = b0;
    *reinterpret_cast<float4*>(&Bs[1][lr][lc]) = b1;
    __syncthreads();
#pragma unroll
    for (int term = 0; term < 2; ++term) {
#pragma unroll
      for (int kk = 0; kk < kWgK; ++kk) {
        const float4 av = *reinterpret_cast<const float4*>(&As[term][kk][ti]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[term][kk][tj]);
        const float a[4] = {av.x, av.y, av.z, av.w};
        const float b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int w = 0; w < 4; ++w) acc[term][u][w] = fmaf(a[u], b[w], acc[term][u][w]);
      }
    }
  }
  const float s0 = scales[0], s1 = scales[1];
  for (int u = 0; u < 4; ++u)
    for (int w = 0; w < 4; ++w) {
      const int i = i0 + ti + u, j = j0 + tj + w;
      if (i < d && j < n) atomicAdd(dW + static_cast<int64_t>(i) * n + j, fmaf(s0, acc[0][u][w], s1 * acc[1][u][w]));
    }
}

// (bf16 mode forms the tied weight gradient on the tensor cores instead: two products over the token axis through
//  MN-major operand descriptors, freud_gemm_tn_splitk -- no operand packing pass.)

}  // namespace freud

using namespace freud;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int freud_l1_colnorm(float* W, float* Wt, int64_t d, int64_t n, void* stream) {
  FREUD_REQUIRE(d > 0 && n > 0, "colnorm needs d, n > 0");
  l1_colnorm_kernel<<<(int)((n + 31) / 32), 256, 0, STREAM>>>(W, Wt, (int)d, (int)n);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_l1_loss_reduce(const float* latent, const float* x_hat, const float* x, float* dxhat,
                                    double* acc, int64_t N, int64_t d, int64_t n, void* stream) {
  const int64_t work = N * (d > n ? d : n);
  int64_t grid = (work + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (grid > cap) grid = cap;
  l1_loss_kernel<<<(int)grid, 256, 0, STREAM>>>(latent, x_hat, x, dxhat, acc, N * n, N * d);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_l1_dz(float* dc, const float* latent, const float* scales, float* db, int64_t N, int64_t n,
                           void* stream) {
  const int rows_per_slab = 256;
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)((N + rows_per_slab - 1) / rows_per_slab));
  l1_dz_kernel<<<grid, 256, 0, STREAM>>>(dc, latent, scales, db, N, (int)n, rows_per_slab);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_l1_weight_grad(const float* x, const float* dz, const float* dxhat, const float* latent,
                                    const float* scales, float* dW, int64_t N, int64_t d, int64_t n, void* stream) {
  FREUD_REQUIRE(d % 4 == 0, "l1_weight_grad needs d % 4 == 0");
  FREUD_CHECK_CUDA(cudaMemsetAsync(dW, 0, d * n * sizeof(float), STREAM));
  const int tiles = (int)(((d + kWgTile - 1) / kWgTile) * ((n + kWgTile - 1) / kWgTile));
  int64_t slabs = (static_cast<int64_t>(sm_count()) * 4 + tiles - 1) / tiles;
  int64_t tokens_per_slab = (N + slabs - 1) / slabs;
  tokens_per_slab = ((tokens_per_slab + kWgK - 1) / kWgK) * kWgK;
  slabs = (N + tokens_per_slab - 1) / tokens_per_slab;
  dim3 grid((unsigned)((n + kWgTile - 1) / kWgTile), (unsigned)((d + kWgTile - 1) / kWgTile), (unsigned)slabs);
  l1_weight_grad_kernel<<<grid, 256, 0, STREAM>>>(x, dz, dxhat, latent, scales, dW, N, (int)d, (int)n,
                                                  tokens_per_slab);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}


// Loss values and gradient scales of the L1 SAE from the four accumulated sums (l1autoencoder.py:85-86,29-36):
//   acc = [sum|c|, sum over x != -1 of (x_hat - x)^2, count of x != -1, sum (x_hat - x)^2]
//   out = [l1_loss, reconstruction_loss, mse, d recon / d x_hat scale 2 alpha / count, d l1 / d c scale 1 / N]
// One launch instead of a dozen scalar torch ops: the L1 step is bound by host enqueue time.
__global__ void l1_loss_scalars_kernel(const double* __restrict__ acc, double n_glob, double d, double recon_alpha,
                                       float* __restrict__ out) {
  out[0] = static_cast<float>(acc[0] / n_glob);
  out[1] = static_cast<float>(recon_alpha * acc[1] / acc[2]);
  out[2] = static_cast<float>(acc[3] / (n_glob * d));
  out[3] = static_cast<float>(2.0 * recon_alpha / acc[2]);
  out[4] = static_cast<float>(1.0 / n_glob);
}

extern "C" int freud_l1_loss_scalars(const double* acc, double n_glob, double d, double recon_alpha, float* out,
                                     void* stream) {
  FREUD_REQUIRE(acc != nullptr && out != nullptr && n_glob > 0 && d > 0, "l1_loss_scalars: bad arguments");
  l1_loss_scalars_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(acc, n_glob, d, recon_alpha, out);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}
