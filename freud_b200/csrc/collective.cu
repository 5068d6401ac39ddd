// Fused data-parallel optimiser step over NVLink peer memory (SURVEY.md 8(e) data parallel, section 5 "allreduce
// fused into the Adam kernel's read").  The reference is single-device; N ranks must reproduce the single-GPU step on
// the concatenated batch (train_sae.py:448-450: backward, clip_grad_norm_, Adam), which needs the SUM of the ranks'
// gradients, its global norm, and identical updated parameters everywhere.
//
// Instead of  all-reduce(gradients) -> grad-norm pass -> replicated Adam  (2 x 151 MB over the links per rank at C3,
// then 1.06 GB of HBM traffic for the update on EVERY rank) the flat parameter space is cut into one contiguous
// slice per rank and every rank runs two kernels on ITS slice only:
//
//   freud_dp_reduce_scatter : g_slice = sum_r peer_r.grad[slice]  (16-byte loads from every peer's gradient buffer
//                             over NVLink / NVSwitch, fixed rank order -> bit-identical whoever owns the slice),
//                             written back in place, with its sum of squares (the owner's share of the global
//                             gradient norm), which the last block posts into every peer's scalar slot
//   -- cross-rank barrier (symmetric-memory signal pads, enqueued by the host side) --
//   freud_dp_adam_allgather : clip coefficient from the G posted partial norms (fixed order), Adam on the slice
//                             (fp32 master, exp_avg, exp_avg_sq: 1/G of the state traffic), and the UPDATED values go
//                             straight to every rank: weights as the bf16 copies the tensor-core / gather kernels
//                             read (2 B/param), biases as fp32 -- peer stores, or one multimem.st per 16 bytes when
//                             the buffers have an NVSwitch multicast mapping
//   -- cross-rank barrier --
//
// Per rank and step the links carry (G-1)/G x 4 B/param in and (G-1)/G x 2 B/param out instead of ~2 x 4 B/param
// each way, the update touches 1/G of the optimiser state, and there is no separate gradient-norm pass.
// Pointers into peer memory come from torch.distributed._symmetric_memory (the host side owns the allocations and
// the communicator, as SURVEY.md 8(b) prescribes); this file only dereferences them.
#include "device_utils.cuh"
#include "host_common.h"
#include "../../include/freud_b200.h"

#include <cmath>

namespace freud {

constexpr int kMaxRanks = 16;
struct PeerF32 {
  float* p[kMaxRanks];
};
struct PeerBf16 {
  __nv_bfloat16* p[kMaxRanks];
};
struct PeerF64 {
  double* p[kMaxRanks];
};

// relaxed system-scope 16-byte load: peer memory is written by other GPUs' kernels (ordered by the barrier before
// this kernel), so the load must not be served from a stale non-coherent path
__device__ __forceinline__ float4 ld_peer(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_peer(__nv_bfloat16* p, uint4 v) {
  asm volatile("st.relaxed.sys.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
// one store, replicated by the NVSwitch into every rank's copy of the buffer
__device__ __forceinline__ void st_multicast(void* mc, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(__uint_as_float(v.x)),
               "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
               : "memory");
}
// in-switch reduction of the G copies of 16 bytes (fp32 add)
__device__ __forceinline__ float4 ld_reduce_multicast(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}

// grad[lo:hi) <- sum over ranks (in place in this rank's buffer); sumsq of the result -> partial[0], and, by the last
// block to finish, into slot `rank` of every peer's scalar buffer.  lo, hi multiples of 4.
template <int GT>  // GT > 0: world size known at compile time (all peer loads of an element in flight); 0: generic
__global__ void __launch_bounds__(256) dp_reduce_scatter_kernel(PeerF32 grads, const float* mc_grad, int G, int rank,
                                                                int64_t lo, int64_t hi, double* partial,
                                                                unsigned int* done_ctr, PeerF64 slots) {
  __shared__ double scratch[32];
  __shared__ bool is_last;
  float* mine = grads.p[rank];
  const int64_t n4 = (hi - lo) >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  double acc = 0.0;
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (mc_grad != nullptr) {
    // in-switch reduction: four independent 16-byte multimem loads in flight per thread (one request per iteration
    // left the links at ~360 GB/s: the round trip through the switch is a few microseconds)
    for (; i + 3 * stride < n4; i += 4 * stride) {
      float4 s[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) s[u] = ld_reduce_multicast(mc_grad + lo + (i + u * stride) * 4);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        *reinterpret_cast<float4*>(mine + lo + (i + u * stride) * 4) = s[u];
        acc += (double)(s[u].x * s[u].x + s[u].y * s[u].y) + (double)(s[u].z * s[u].z + s[u].w * s[u].w);
      }
    }
  }
  for (; i < n4; i += stride) {
    const int64_t e = lo + i * 4;
    float4 s;
    if (mc_grad != nullptr) {
      s = ld_reduce_multicast(mc_grad + e);
    } else if constexpr (GT > 0) {
      float4 v[GT];
#pragma unroll
      for (int r = 0; r < GT; ++r) v[r] = ld_peer(grads.p[r] + e);  // all peers' loads in flight before the first add
      s = v[0];
#pragma unroll
      for (int r = 1; r < GT; ++r) {
        s.x += v[r].x;
        s.y += v[r].y;
        s.z += v[r].z;
        s.w += v[r].w;
      }
    } else {
      s = ld_peer(grads.p[0] + e);
      for (int r = 1; r < G; ++r) {
        const float4 t = ld_peer(grads.p[r] + e);
        s.x += t.x;
        s.y += t.y;
        s.z += t.z;
        s.w += t.w;
      }
    }
    *reinterpret_cast<float4*>(mine + e) = s;
    acc += (double)(s.x * s.x + s.y * s.y) + (double)(s.z * s.z + s.w * s.w);
  }
  const double tot = block_sum(acc, scratch);
  if (threadIdx.x == 0) {
    if (tot != 0.0) atomicAdd(partial, tot);
    __threadfence();
    is_last = atomicAdd(done_ctr, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {  // block-uniform
    __threadfence();
    const double total = *reinterpret_cast<volatile double*>(partial);
    __syncthreads();  // every thread holds the total before thread 0 re-arms the accumulator
    if (threadIdx.x < G) {
      asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(slots.p[threadIdx.x] + rank), "d"(total) : "memory");
      __threadfence_system();
    }
    if (threadIdx.x == 0) {  // self-reset for the next step (same stream: no launch overlaps this one)
      *partial = 0.0;
      *done_ctr = 0u;
    }
  }
}

struct DpRegions {
  int64_t begin[4], end[4];  // element ranges of the flat space, multiples of 8
  int is_weight[4];          // 1: broadcast as bf16 into the shadow buffers; 0: broadcast as fp32 into the params
  int count;
};
struct DpAdamArgs {
  float beta1, beta2, omb1, omb2, eps, step_size, bc2_sqrt, max_norm;
};

// Adam on [lo, hi) of the flat space + broadcast of the updated values.  m / v hold only this rank's slice
// (index e - lo).  8 elements per thread and iteration: one 16-byte bf16 store per destination.
__global__ void __launch_bounds__(256) dp_adam_allgather_kernel(PeerF32 params, PeerBf16 shadows, float* mc_param,
                                                                __nv_bfloat16* mc_shadow, const float* __restrict__ grad,
                                                                float* __restrict__ m, float* __restrict__ v, int G,
                                                                int rank, int64_t lo, int64_t hi, DpRegions regs,
                                                                DpAdamArgs a, const double* __restrict__ slots,
                                                                int clip) {
  float coef = 1.f;
  if (clip) {
    double sumsq = 0.0;
    for (int r = 0; r < G; ++r) sumsq += slots[r];  // fixed order: the same total on every rank
    const float total = static_cast<float>(sqrt(sumsq));
    coef = fminf(a.max_norm / (total + 1e-6f), 1.0f);  // clip_grad.py:165-169
  }
  float* mine = params.p[rank];
  const int64_t n8 = (hi - lo) >> 3;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const int64_t e = lo + i * 8, l = i * 8;
    int weight = 0, inside = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (r < regs.count && e >= regs.begin[r] && e < regs.end[r]) {
        inside = 1;
        weight = regs.is_weight[r];
      }
    if (!inside) continue;  // padding between tensors
    float pv[8], gv[8], mv[8], vv[8];
    *reinterpret_cast<float4*>(pv) = *reinterpret_cast<const float4*>(mine + e);
    *reinterpret_cast<float4*>(pv + 4) = *reinterpret_cast<const float4*>(mine + e + 4);
    *reinterpret_cast<float4*>(gv) = *reinterpret_cast<const float4*>(grad + e);
    *reinterpret_cast<float4*>(gv + 4) = *reinterpret_cast<const float4*>(grad + e + 4);
    *reinterpret_cast<float4*>(mv) = *reinterpret_cast<const float4*>(m + l);
    *reinterpret_cast<float4*>(mv + 4) = *reinterpret_cast<const float4*>(m + l + 4);
    *reinterpret_cast<float4*>(vv) = *reinterpret_cast<const float4*>(v + l);
    *reinterpret_cast<float4*>(vv + 4) = *reinterpret_cast<const float4*>(v + l + 4);
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // torch/optim/adam.py:347, same operation order as adam_kernel (optim.cu)
      const float g = gv[j] * coef;
      mv[j] = mv[j] + (g - mv[j]) * a.omb1;
      vv[j] = vv[j] * a.beta2 + a.omb2 * (g * g);
      const float denom = sqrtf(vv[j]) / a.bc2_sqrt + a.eps;
      pv[j] = pv[j] - a.step_size * (mv[j] / denom);
    }
    *reinterpret_cast<float4*>(m + l) = *reinterpret_cast<const float4*>(mv);
    *reinterpret_cast<float4*>(m + l + 4) = *reinterpret_cast<const float4*>(mv + 4);
    *reinterpret_cast<float4*>(v + l) = *reinterpret_cast<const float4*>(vv);
    *reinterpret_cast<float4*>(v + l + 4) = *reinterpret_cast<const float4*>(vv + 4);
    const float4 p0 = *reinterpret_cast<const float4*>(pv), p1 = *reinterpret_cast<const float4*>(pv + 4);
    *reinterpret_cast<float4*>(mine + e) = p0;  // fp32 master (this rank's slice)
    *reinterpret_cast<float4*>(mine + e + 4) = p1;
    if (weight) {
      uint4 q;
      uint32_t* w = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(pv[2 * j], pv[2 * j + 1]);
        w[j] = *reinterpret_cast<const uint32_t*>(&h);
      }
      if (mc_shadow != nullptr) {
        st_multicast(mc_shadow + e, q);
      } else {
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r)
          if (r < G) st_peer(shadows.p[r] + e, q);
      }
    } else {
      if (mc_param != nullptr) {
        st_multicast(mc_param + e, *reinterpret_cast<const uint4*>(&p0));
        st_multicast(mc_param + e + 4, *reinterpret_cast<const uint4*>(&p1));
      } else {
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r)
          if (r < G && r != rank) {
            st_peer(params.p[r] + e, p0);
            st_peer(params.p[r] + e + 4, p1);
          }
      }
    }
  }
}

}  // namespace freud

using namespace freud;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int freud_dp_reduce_scatter(void* const* peer_grads, const float* mc_grad, int64_t world, int64_t rank,
                                       int64_t lo, int64_t hi, double* partial, unsigned int* done_ctr,
                                       void* const* peer_slots, void* stream) {
  FREUD_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "dp_reduce_scatter: bad rank / world");
  FREUD_REQUIRE(lo >= 0 && hi > lo && lo % 8 == 0 && hi % 8 == 0, "slice bounds must be multiples of 8");
  PeerF32 g{};
  PeerF64 s{};
  for (int r = 0; r < world; ++r) {
    g.p[r] = static_cast<float*>(peer_grads[r]);
    s.p[r] = static_cast<double*>(peer_slots[r]);
  }
  const int64_t n4 = (hi - lo) / 4;
  int64_t grid = (n4 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 6;
  if (grid > cap) grid = cap;
#define LAUNCH_RS(GT)                                                                                          \
  dp_reduce_scatter_kernel<GT><<<(unsigned)grid, 256, 0, STREAM>>>(g, mc_grad, (int)world, (int)rank, lo, hi, partial, \
                                                                   done_ctr, s)
  switch (world) {
    case 1: LAUNCH_RS(1); break;
    case 2: LAUNCH_RS(2); break;
    case 4: LAUNCH_RS(4); break;
    case 8: LAUNCH_RS(8); break;
    default: LAUNCH_RS(0); break;
  }
#undef LAUNCH_RS
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int freud_dp_adam_allgather(void* const* peer_params, void* const* peer_shadows, float* mc_param,
                                       void* mc_shadow, const float* grad, float* exp_avg, float* exp_avg_sq,
                                       int64_t world, int64_t rank, int64_t lo, int64_t hi, const int64_t* region_begin,
                                       const int64_t* region_end, const int32_t* region_is_weight, int64_t n_regions,
                                       double lr, double beta1, double beta2, double eps, int64_t step,
                                       const double* slots, float max_norm, int clip, void* stream) {
  FREUD_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "dp_adam_allgather: bad rank / world");
  FREUD_REQUIRE(lo >= 0 && hi > lo && lo % 8 == 0 && hi % 8 == 0, "slice bounds must be multiples of 8");
  FREUD_REQUIRE(n_regions >= 1 && n_regions <= 4 && step >= 1, "1..4 regions, 1-based step");
  PeerF32 p{};
  PeerBf16 sh{};
  for (int r = 0; r < world; ++r) {
    p.p[r] = static_cast<float*>(peer_params[r]);
    sh.p[r] = peer_shadows ? static_cast<__nv_bfloat16*>(peer_shadows[r]) : nullptr;
  }
  DpRegions regs{};
  regs.count = static_cast<int>(n_regions);
  for (int r = 0; r < n_regions; ++r) {
    FREUD_REQUIRE(region_begin[r] % 8 == 0 && region_end[r] % 8 == 0, "region bounds must be multiples of 8");
    FREUD_REQUIRE(!region_is_weight[r] || peer_shadows != nullptr, "weight regions need the bf16 shadow buffers");
    regs.begin[r] = region_begin[r];
    regs.end[r] = region_end[r];
    regs.is_weight[r] = region_is_weight[r];
  }
  DpAdamArgs a{};
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
  const int64_t n8 = (hi - lo) / 8;
  int64_t grid = (n8 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 4;
  if (grid > cap) grid = cap;
  dp_adam_allgather_kernel<<<(unsigned)grid, 256, 0, STREAM>>>(p, sh, mc_param, static_cast<__nv_bfloat16*>(mc_shadow),
                                                               grad, exp_avg, exp_avg_sq, (int)world, (int)rank, lo, hi,
                                                               regs, a, slots, clip);
  FREUD_CHECK_CUDA(cudaGetLastError());
  return 0;
}
