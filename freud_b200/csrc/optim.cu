// Optimiser side of the train step (train_sae.py:449-450): global gradient norm, clipping, and fused
// multi-tensor Adam / RAdam updates with the clip coefficient applied on the fly.  28 B/param of HBM traffic.
#include "device_utils.cuh"
#include "host_common.h"
#include "../../include/freud_b200.h"

#include <cmath>

namespace freud {

struct TensorList {
  int count;
  float* param[FREUD_MAX_TENSORS];
  float* grad[FREUD_MAX_TENSORS];
  float* m[FREUD_MAX_TENSORS];
  float* v[FREUD_MAX_TENSORS];
  __nv_bfloat16* shadow[FREUD_MAX_TENSORS];
  int64_t numel[FREUD_MAX_TENSORS];
};

static int convert_list(const freud_tensor_list* in, TensorList* out) {
  if (in == nullptr || in->count < 1 || in->count > FREUD_MAX_TENSORS) return 1;
  out->count = in->count;
  for (int i = 0; i < in->count; ++i) {
    out->param[i] = in->param[i];
    out->grad[i] = in->grad[i];
    out->m[i] = in->exp_avg[i];
    out->v[i] = in->exp_avg_sq[i];
    out->shadow[i] = static_cast<__nv_bfloat16*>(in->bf16_shadow[i]);
    out->numel[i] = in->numel[i];
  }
  return 0;
}

__device__ __forceinline__ float clip_coef(const double* sumsq, float max_norm) {
  // clip_grad.py:165-169: clamp(max_norm / (total_norm + 1e-6), max=1.0)
  const float total = static_cast<float>(sqrt(*sumsq));
  return fminf(max_norm / (total + 1e-6f), 1.0f);
}

__global__ void __launch_bounds__(256) grad_sumsq_kernel(TensorList tl, double* __restrict__ sumsq) {
  __shared__ double scratch[32];
  const int ti = blockIdx.y;
  const float* __restrict__ g = tl.grad[ti];
  const int64_t n = tl.numel[ti];
  double acc = 0.0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if ((n & 3) == 0) {
    const int64_t n4 = n >> 2;
    int64_t i = i0;
    for (; i + 3 * stride < n4; i += 4 * stride) {  // four 16-byte loads in flight per thread
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = load4(g + (i + u * stride) * 4);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        acc += (double)(v[u].x * v[u].x + v[u].y * v[u].y) + (double)(v[u].z * v[u].z + v[u].w * v[u].w);
    }
    for (; i < n4; i += stride) {
      const float4 v = load4(g + i * 4);
      acc += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
    }
  } else {
    for (int64_t i = i0; i < n; i += stride) acc += (double)g[i] * g[i];
  }
  const double tot = block_sum(acc, scratch);
  if (threadIdx.x == 0 && tot != 0.0) atomicAdd(sumsq, tot);
}

__global__ void __launch_bounds__(256) clip_grads_kernel(TensorList tl, const double* __restrict__ sumsq,
                                                         float max_norm, float* __restrict__ norm_out) {
  const int ti = blockIdx.y;
  float* __restrict__ g = tl.grad[ti];
  const int64_t n = tl.numel[ti];
  const float coef = clip_coef(sumsq, max_norm);
  if (ti == 0 && blockIdx.x == 0 && threadIdx.x == 0 && norm_out) *norm_out = static_cast<float>(sqrt(*sumsq));
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  // torch multiplies unconditionally (clip_grad.py:121), coef == 1 included
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) g[i] *= coef;
}

struct AdamArgs {
  float beta1, beta2, omb1, omb2, eps, step_size, bc2_sqrt;  // Adam (omb = 1 - beta, rounded from double)
  float lr, weight_decay, bc1, rect;             // RAdam extras
  int rectify;                                   // RAdam: rho_t > 5
  float max_norm;
};

template <bool RADAM>
__device__ __forceinline__ void update_one(float& p, float g, float& m, float& v, const AdamArgs& a) {
  if (RADAM && a.weight_decay != 0.f) g = fmaf(a.weight_decay, p, g);
  m = m + (g - m) * a.omb1;                 // exp_avg.lerp_(grad, 1 - beta1)
  v = v * a.beta2 + a.omb2 * (g * g);       // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  if (!RADAM) {
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    p = p - a.step_size * (m / denom);
  } else {
    const float m_hat = m / a.bc1;
    if (a.rectify) {
      const float adaptive = a.bc2_sqrt / (sqrtf(v) + a.eps);
      p = p - m_hat * a.lr * adaptive * a.rect;
    } else {
      p = p - m_hat * a.lr;
    }
  }
}

template <bool RADAM>
__global__ void __launch_bounds__(256) adam_kernel(TensorList tl, AdamArgs a, const double* __restrict__ sumsq) {
  const int ti = blockIdx.y;
  float* __restrict__ p = tl.param[ti];
  const float* __restrict__ g = tl.grad[ti];
  float* __restrict__ m = tl.m[ti];
  float* __restrict__ v = tl.v[ti];
  __nv_bfloat16* __restrict__ sh = tl.shadow[ti];
  const int64_t n = tl.numel[ti];
  const float coef = sumsq ? clip_coef(sumsq, a.max_norm) : 1.f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if ((n & 3) == 0) {
    for (int64_t i = i0; i < (n >> 2); i += stride) {
      float4 pv = *reinterpret_cast<const float4*>(p + i * 4);
      const float4 gv = load4(g + i * 4);
      float4 mv = *reinterpret_cast<const float4*>(m + i * 4);
      float4 vv = *reinterpret_cast<const float4*>(v + i * 4);
      update_one<RADAM>(pv.x, gv.x * coef, mv.x, vv.x, a);
      update_one<RADAM>(pv.y, gv.y * coef, mv.y, vv.y, a);
      update_one<RADAM>(pv.z, gv.z * coef, mv.z, vv.z, a);
      update_one<RADAM>(pv.w, gv.w * coef, mv.w, vv.w, a);
      store4(p + i * 4, pv);
      store4(m + i * 4, mv);
      store4(v + i * 4, vv);
      if (sh) store4(sh + i * 4, pv);
    }
  } else {
    for (int64_t i = i0; i < n; i += stride) {
      float pv = p[i], mv = m[i], vv = v[i];
      update_one<RADAM>(pv, g[i] * coef, mv, vv, a);
      p[i] = pv;
      m[i] = mv;
      v[i] = vv;
      if (sh) sh[i] = __float2bfloat16_rn(pv);
    }
  }
}

static dim3 list_grid(const TensorList& tl) {
  int64_t mx = 1;
  for (int i = 0; i < tl.count; ++i) mx = tl.numel[i] > mx ? tl.numel[i] : mx;
  int64_t gx = (mx / 4 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3(static_cast<unsigned>(gx), static_cast<unsigned>(tl.count));
}

}  // namespace freud

using namespace freud;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int freud_grad_sumsq(const freud_tensor_list* host_list, double* sumsq, void* stream) {
  TensorList tl;
  FREUD_REQUIRE(convert_list(host_list, &tl) == 0, "bad tensor list");
  FREUD_CHECK_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(double), STREAM));
  grad_sumsq_kernel<<<list_grid(tl), 256, 0, STREAM>>>(tl, sumsq);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_clip_grads(const freud_tensor_list* host_list, const double* sumsq, float max_norm,
                                float* norm_out, void* stream) {
  TensorList tl;
  FREUD_REQUIRE(convert_list(host_list, &tl) == 0, "bad tensor list");
  clip_grads_kernel<<<list_grid(tl), 256, 0, STREAM>>>(tl, sumsq, max_norm, norm_out);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_adam_step(const freud_tensor_list* host_list, double lr, double beta1, double beta2, double eps,
                               int64_t step, const double* sumsq, float max_norm, void* stream) {
  TensorList tl;
  FREUD_REQUIRE(convert_list(host_list, &tl) == 0, "bad tensor list");
  FREUD_REQUIRE(step >= 1, "step is 1-based");
  AdamArgs a{};
  a.beta1 = static_cast<float>(beta1);
  a.beta2 = static_cast<float>(beta2);
  a.eps = static_cast<float>(eps);
  a.omb1 = static_cast<float>(1.0 - beta1);
  a.omb2 = static_cast<float>(1.0 - beta2);
  const double bc1 = 1.0 - std::pow(beta1, (double)step);
  const double bc2 = 1.0 - std::pow(beta2, (double)step);
  a.step_size = static_cast<float>(lr / bc1);
  a.bc2_sqrt = static_cast<float>(std::sqrt(bc2));
  a.max_norm = max_norm;
  adam_kernel<false><<<list_grid(tl), 256, 0, STREAM>>>(tl, a, sumsq);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_radam_step(const freud_tensor_list* host_list, double lr, double beta1, double beta2, double eps,
                                double weight_decay, int64_t step, const double* sumsq, float max_norm,
                                void* stream) {
  TensorList tl;
  FREUD_REQUIRE(convert_list(host_list, &tl) == 0, "bad tensor list");
  FREUD_REQUIRE(step >= 1, "step is 1-based");
  AdamArgs a{};
  a.beta1 = static_cast<float>(beta1);
  a.beta2 = static_cast<float>(beta2);
  a.eps = static_cast<float>(eps);
  a.omb1 = static_cast<float>(1.0 - beta1);
  a.omb2 = static_cast<float>(1.0 - beta2);
  a.lr = static_cast<float>(lr);
  a.weight_decay = static_cast<float>(weight_decay);
  const double b2t = std::pow(beta2, (double)step);
  const double bc1 = 1.0 - std::pow(beta1, (double)step);
  const double bc2 = 1.0 - b2t;
  a.bc1 = static_cast<float>(bc1);
  a.bc2_sqrt = static_cast<float>(std::sqrt(bc2));
  const double rho_inf = 2.0 / (1.0 - beta2) - 1.0;
  const double rho_t = rho_inf - 2.0 * (double)step * b2t / bc2;
  a.rectify = rho_t > 5.0 ? 1 : 0;
  a.rect = a.rectify ? static_cast<float>(std::sqrt((rho_t - 4) * (rho_t - 2) * rho_inf /
                                                    ((rho_inf - 4) * (rho_inf - 2) * rho_t)))
                     : 1.f;
  a.max_norm = max_norm;
  adam_kernel<true><<<list_grid(tl), 256, 0, STREAM>>>(tl, a, sumsq);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}
