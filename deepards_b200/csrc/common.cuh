// Shared device/host helpers for the deepards_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/deepards_b200.h"

namespace dards {

void set_error(const char* fmt, ...);
void count_launch();  // every kernel launch of the library goes through DARDS_CHECK_LAUNCH

#define DARDS_CHECK_ARG(cond, ...)            \
  do {                                        \
    if (!(cond)) {                            \
      ::dards::set_error(__VA_ARGS__);        \
      return DARDS_ERR_INVALID_ARGUMENT;      \
    }                                         \
  } while (0)

#define DARDS_CHECK_LAUNCH(name)                                                   \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      ::dards::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));  \
      return DARDS_ERR_CUDA;                                                       \
    }                                                                              \
    ::dards::count_launch();                                                       \
  } while (0)

// ---- element type traits: activations are fp32 or bf16, arithmetic is always fp32 ----
template <typename T> struct Elem;
template <> struct Elem<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
  static __device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
  static __device__ __forceinline__ float round(float v) { return v; }
};
template <> struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
  static __device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
    uint2 r = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  static __device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
  }
  // value as it will be seen by the next reader of the stored tensor
  static __device__ __forceinline__ float round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int RED_THREADS = 256;  // block size assumed by the last-CTA reduction helpers below

// Ticket: true in every thread of the CTA that is the last of `expected` CTAs to arrive at `counter`.
// Global writes made by ANY thread of the CTA before the call are visible to the last CTA after it: bar.sync orders
// them before thread 0's gpu-scope fence, and fences are cumulative.  Only ONE thread fences -- a MEMBAR.GPU waits for
// the outstanding stores and invalidates the SM's L1, so call this BEFORE a kernel's bulk output stores, right after
// the few values the last CTA needs have been written.
__device__ __forceinline__ bool last_cta_arrives(unsigned int* counter, unsigned int expected, int* flag_smem) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int old = atomicAdd(counter, 1u);
    const int last = (old == expected - 1u);
    if (last) {
      *counter = 0u;  // self-reset: the next launch (or CUDA-graph replay) starts from zero
      __threadfence();
    }
    *flag_smem = last;
  }
  __syncthreads();
  return *flag_smem != 0;
}

// ---- last-CTA reductions over the groups ------------------------------------------------------------------
// One CTA (256 threads) reduces [n_groups][c] fp32 tables over the groups for `nch` channels starting at ch0
// (nch a multiple of 4, <= 64): threads = (nch/4 float4 columns) x (256/(nch/4) group lanes); every lane keeps 8
// independent 16-byte loads in flight (the reduction is pure L2 latency otherwise) and the lane sums are combined
// in a fixed order -> deterministic.  `wfun(g)` is the weight of group g.
template <typename WF>
__device__ __forceinline__ float4 group_reduce4(const float* table, int n_groups, int c, int ch0, int nch, WF wfun,
                                                float4* scratch /* 256 float4 */) {
  const int ncol = nch >> 2;
  const int lanes = RED_THREADS / ncol;
  const int col = threadIdx.x % ncol, ln = threadIdx.x / ncol;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ln < lanes) {
    const float* base = table + ch0 + col * 4;
#pragma unroll 8
    for (int g = ln; g < n_groups; g += lanes) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(base + (size_t)g * c));
      const float w = wfun(g);
      s.x = fmaf(w, v.x, s.x); s.y = fmaf(w, v.y, s.y); s.z = fmaf(w, v.z, s.z); s.w = fmaf(w, v.w, s.w);
    }
  }
  __syncthreads();
  scratch[threadIdx.x] = s;
  __syncthreads();
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
  if (threadIdx.x < ncol) {
    for (int k = 0; k < lanes; ++k) {
      const float4 v = scratch[k * ncol + threadIdx.x];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
  }
  return t;  // valid in threads [0, ncol): channels ch0 + 4*threadIdx.x .. +3
}

// nn.BatchNorm1d updates its running statistics once per group IN ORDER:  r <- (1-m) r + m v_g,  g = 0..G-1.
// Closed form (SURVEY.md hard part 6):  r_G = (1-m)^G r_0 + m * sum_g (1-m)^(G-1-g) v_g  -- a weighted reduction.
// The variance is recovered from the saved rstd: var_g = 1/rstd_g^2 - eps (biased) -> unbiased.
// Called by all 256 threads of one CTA for the channels [ch0, ch0 + nch).
__device__ __forceinline__ void running_update_tile(const float* save_mean, const float* save_rstd, float* rm, float* rv,
                                                    int n_groups, int rows, int c, int ch0, int nch, float momentum,
                                                    float eps, float4* scratch) {
  const float unbias = rows > 1 ? (float)rows / (float)(rows - 1) : 1.f;
  const float lg = log2f(1.f - momentum);
  auto wfun = [&](int g) { return momentum * exp2f(lg * (float)(n_groups - 1 - g)); };
  const float4 tm = group_reduce4(save_mean, n_groups, c, ch0, nch, wfun, scratch);
  // weighted sum of 1/rstd^2 - eps: reduce the transformed values
  const int ncol = nch >> 2;
  const int lanes = RED_THREADS / ncol;
  const int col = threadIdx.x % ncol, ln = threadIdx.x / ncol;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ln < lanes) {
    const float* base = save_rstd + ch0 + col * 4;
#pragma unroll 8
    for (int g = ln; g < n_groups; g += lanes) {
      const float4 r = __ldcg(reinterpret_cast<const float4*>(base + (size_t)g * c));
      const float w = wfun(g) * unbias;
      s.x = fmaf(w, fmaxf(1.f / (r.x * r.x) - eps, 0.f), s.x);
      s.y = fmaf(w, fmaxf(1.f / (r.y * r.y) - eps, 0.f), s.y);
      s.z = fmaf(w, fmaxf(1.f / (r.z * r.z) - eps, 0.f), s.z);
      s.w = fmaf(w, fmaxf(1.f / (r.w * r.w) - eps, 0.f), s.w);
    }
  }
  __syncthreads();
  scratch[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < ncol) {
    float4 tv = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < lanes; ++k) {
      const float4 v = scratch[k * ncol + threadIdx.x];
      tv.x += v.x; tv.y += v.y; tv.z += v.z; tv.w += v.w;
    }
    const float decay = exp2f(lg * (float)n_groups);
    float4* pm = reinterpret_cast<float4*>(rm + ch0) + threadIdx.x;
    float4* pv = reinterpret_cast<float4*>(rv + ch0) + threadIdx.x;
    float4 m = *pm, v = *pv;
    m.x = decay * m.x + tm.x; m.y = decay * m.y + tm.y; m.z = decay * m.z + tm.z; m.w = decay * m.w + tm.w;
    v.x = decay * v.x + tv.x; v.y = decay * v.y + tv.y; v.z = decay * v.z + tv.z; v.w = decay * v.w + tv.w;
    *pm = m;
    *pv = v;
  }
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// forward / dgrad implicit GEMM of the CUDA-core path (conv_simt.cu)
struct ConvGemmArgs {
  const void* in;
  const void* w;       // [t][c_red][c_cols]
  void* out;
  const void* addend;  // may be null
  long long m_total;   // n_breaths * l_dst
  int l_src, l_dst, c_red, c_cols, src_stride, dst_stride, addend_stride;
  int ktaps, q_mul, t_mul, off, div;  // source position = (q*q_mul + t*t_mul + off) / div
};

#define DARDS_DISPATCH_DTYPE(dtype, ...)                                  \
  if ((dtype) == DARDS_F32) {                                             \
    using T = float;                                                      \
    __VA_ARGS__                                                           \
  } else if ((dtype) == DARDS_BF16) {                                     \
    using T = __nv_bfloat16;                                              \
    __VA_ARGS__                                                           \
  } else {                                                                \
    ::dards::set_error("unknown dtype %d", (int)(dtype));                 \
    return DARDS_ERR_INVALID_ARGUMENT;                                    \
  }

}  // namespace dards
