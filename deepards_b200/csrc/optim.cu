// Post-all-reduce step: gradient scale + clamp + SGD-Nesterov / Adam on flat fp32 buffers.
// One pass over (param, grad, state): purely bandwidth bound, float4 vectorised.
#include "common.cuh"

namespace dards {

__device__ __forceinline__ float clampf(float g, float clip) { return clip > 0.f ? fminf(fmaxf(g, -clip), clip) : g; }

__global__ void clamp_sgd_nesterov_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                          long long n, float lr, float mom, float wd, float clip, float gscale,
                                          int first) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float w = p[i];
    float d = clampf(g[i] * gscale, clip);
    d = fmaf(wd, w, d);                      // d_p = grad + wd * p
    float buf = first ? d : fmaf(mom, m[i], d);  // buf = mom*buf + d_p
    m[i] = buf;
    d = fmaf(mom, buf, d);                   // nesterov: d_p + mom*buf
    p[i] = w - lr * d;
  }
}

__global__ void clamp_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ ea,
                                  float* __restrict__ es, long long n, float lr, float b1, float b2, float eps,
                                  float clip, float gscale, float bc1, float bc2_sqrt) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float d = clampf(g[i] * gscale, clip);
    float a = ea[i] = b1 * ea[i] + (1.f - b1) * d;
    float s = es[i] = b2 * es[i] + (1.f - b2) * d * d;
    float denom = sqrtf(s) / bc2_sqrt + eps;
    p[i] = p[i] - (lr / bc1) * (a / denom);
  }
}

static int flat_grid(long long n) {
  long long b = (n + 255) / 256;
  if (b > 148LL * 8) b = 148LL * 8;
  if (b < 1) b = 1;
  return (int)b;
}

int launch_clamp_sgd(float* p, const float* g, float* m, long long n, float lr, float mom, float wd, float clip,
                     float gscale, int first, cudaStream_t st) {
  if (n == 0) return DARDS_OK;
  clamp_sgd_nesterov_kernel<<<flat_grid(n), 256, 0, st>>>(p, g, m, n, lr, mom, wd, clip, gscale, first);
  DARDS_CHECK_LAUNCH("clamp_sgd_nesterov");
  return DARDS_OK;
}

int launch_clamp_adam(float* p, const float* g, float* ea, float* es, long long n, float lr, float b1, float b2,
                      float eps, float clip, float gscale, int step, cudaStream_t st) {
  if (n == 0) return DARDS_OK;
  DARDS_CHECK_ARG(step >= 1, "adam: step is 1-based");
  float bc1 = 1.f - powf(b1, (float)step);
  float bc2 = 1.f - powf(b2, (float)step);
  clamp_adam_kernel<<<flat_grid(n), 256, 0, st>>>(p, g, ea, es, n, lr, b1, b2, eps, clip, gscale, bc1, sqrtf(bc2));
  DARDS_CHECK_LAUNCH("clamp_adam");
  return DARDS_OK;
}

}  // namespace dards
