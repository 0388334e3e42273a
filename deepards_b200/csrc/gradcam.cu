// GradCAM maps on the device (deepards/gradcam.py:40-65, 83-107, 125-162, 195-205; patient_gradcam.py:213-229).
//
// The reference runs one forward and one one-hot backward through the whole DenseNet per sequence, copies the norm5
// activations A and their gradient dA (20,128,7) to the host and reduces them with numpy.  The head behind A is
// ReLU -> AvgPool1d(7) -> flatten -> Linear, so dA is known in closed form:
//     dA[n,c,l] = (A[n,c,l] > 0) ? W[target, n*F + c] / 7 : 0            (bit-exact with autograd, tests/golden)
// and no convolution backward is needed.  One CTA per sequence computes the logits, picks the target, and reduces
// A and dA to the per-breath ("read") map, the per-sequence map, their min-max-normalised uint8 forms and, when asked,
// the linearly resized 224-sample rows -- one launch for any number of sequences, nothing but the maps goes to the host.
#include "common.cuh"

namespace dards {

constexpr int CAM_THREADS = 256;
constexpr int CAM_WARPS = CAM_THREADS / 32;
constexpr int CAM_MAXL = 8;
constexpr int CAM_MAX_OUT = 8;
constexpr int CAM_MAX_GF = 16384;  // group * F: one byte of shared memory each

using GradcamArgs = dards_gradcam_desc;  // include/deepards_b200.h

// MaxMinNormCam.normalize (gradcam.py:156-161) for one row of `l` values held by the first `l` lanes of a warp:
// relu, (v - min) / (max - min), truncation of v * 255 to uint8.  A constant row is 0/0 in the reference; it maps to 0.
__device__ __forceinline__ uint8_t cam_normalize_lane(float v, int lane, int l) {
  const bool on = lane < l;
  float r = fmaxf(v, 0.f);
  float mn = on ? r : INFINITY, mx = on ? r : -INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const float den = __fsub_rn(mx, mn);
  if (!(den > 0.f)) return 0;
  const float q = __fmul_rn(__fdiv_rn(__fsub_rn(r, mn), den), 255.f);
  return (uint8_t)q;
}

// cv2.resize(column, (1, out_len)) for 8-bit data, default INTER_LINEAR: half-pixel centres, 11-bit coefficients,
// rows pre-scaled by 2^11, vertical pass ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2.
__device__ __forceinline__ uint8_t cam_resize_at(const uint8_t* row, int n, int d, float scale) {
  float fy = ((float)d + 0.5f) * scale - 0.5f;
  int sy = (int)floorf(fy);
  fy -= (float)sy;
  const int y0 = min(max(sy, 0), n - 1), y1 = min(max(sy + 1, 0), n - 1);
  const int b0 = __float2int_rn((1.f - fy) * 2048.f), b1 = __float2int_rn(fy * 2048.f);
  const int s0 = (int)row[y0] * 2048, s1 = (int)row[y1] * 2048;
  return (uint8_t)((((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2);
}

template <typename T>
__global__ void __launch_bounds__(CAM_THREADS) gradcam_kernel(GradcamArgs p) {
  extern __shared__ uint8_t posmask[];  // [group*F]: bit l set <=> A[n,c,l] > 0
  __shared__ float red[CAM_WARPS][CAM_MAX_OUT];
  __shared__ float s_logits[CAM_MAX_OUT];
  __shared__ int s_target;
  __shared__ uint8_t s_read_u8[64 * CAM_MAXL];
  __shared__ uint8_t s_seq_u8[CAM_MAXL];

  const int seq = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int group = p.group, L = p.l, F = p.f, GF = group * F;
  const T* a_seq = static_cast<const T*>(p.a) + (size_t)seq * group * L * p.a_stride;
  const float invL = (float)L;

  // ---- phase 1: pooled features -> logits; sign bits; optional copy of A in the reference layout --------------
  float acc[CAM_MAX_OUT];
#pragma unroll
  for (int j = 0; j < CAM_MAX_OUT; ++j) acc[j] = 0.f;
  for (int idx = tid; idx < GF; idx += CAM_THREADS) {
    const int n = idx / F, c = idx - n * F;
    float s = 0.f;
    uint32_t bits = 0;
    for (int l = 0; l < L; ++l) {
      const float v = Elem<T>::ld(a_seq + (size_t)(n * L + l) * p.a_stride + c);
      s += fmaxf(v, 0.f);
      bits |= (v > 0.f ? 1u : 0u) << l;
      if (p.conv_out) p.conv_out[((size_t)(seq * group + n) * F + c) * L + l] = v;
    }
    posmask[idx] = (uint8_t)bits;
    const float pooled = __fdiv_rn(s, invL);
    for (int j = 0; j < p.n_out; ++j) acc[j] = fmaf(pooled, p.w[(size_t)j * GF + idx], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < CAM_MAX_OUT; ++j) {
    const float v = warp_sum(acc[j]);
    if (lane == 0) red[warp][j] = v;
  }
  __syncthreads();
  if (tid < p.n_out) {
    float v = p.bias[tid];
    for (int w = 0; w < CAM_WARPS; ++w) v += red[w][tid];
    s_logits[tid] = v;
    p.logits[(size_t)seq * p.n_out + tid] = v;
  }
  __syncthreads();
  if (tid == 0) {
    int t = p.target_dev ? p.target_dev[seq] : p.target;
    if (t < 0 || t >= p.n_out) {  // np.argmax: first maximum
      t = 0;
      for (int j = 1; j < p.n_out; ++j)
        if (s_logits[j] > s_logits[t]) t = j;
    }
    s_target = t;
    if (p.target_used) p.target_used[seq] = t;
  }
  __syncthreads();
  const float* wt = p.w + (size_t)s_target * GF;

  // ---- phase 2: per-breath map: cam[n,l] = sum_c mean_l(dA[n,c,:]) * A[n,c,l]; one warp per breath -------------
  for (int n = warp; n < group; n += CAM_WARPS) {
    float cam[CAM_MAXL];
#pragma unroll
    for (int l = 0; l < CAM_MAXL; ++l) cam[l] = 0.f;
    for (int c = lane; c < F; c += 32) {
      const float w7 = __fdiv_rn(wt[n * F + c], invL);
      const uint32_t bits = posmask[n * F + c];
      float s = 0.f;
      for (int l = 0; l < L; ++l) s += ((bits >> l) & 1u) ? w7 : 0.f;  // np.mean(grad, axis=2): sum, then / L
      const float wgt = __fdiv_rn(s, invL);
#pragma unroll
      for (int l = 0; l < CAM_MAXL; ++l)
        if (l < L) {
          const float v = Elem<T>::ld(a_seq + (size_t)(n * L + l) * p.a_stride + c);
          cam[l] = fmaf(wgt, v, cam[l]);
          if (p.grad_out) p.grad_out[((size_t)(seq * group + n) * F + c) * L + l] = ((bits >> l) & 1u) ? w7 : 0.f;
        }
    }
    float mine = 0.f;
#pragma unroll
    for (int l = 0; l < CAM_MAXL; ++l) {
      const float v = warp_sum(cam[l]);
      if (lane == l) mine = v;
    }
    const uint8_t u = cam_normalize_lane(mine, lane, L);
    if (lane < L) {
      const size_t o = (size_t)(seq * group + n) * L + lane;
      if (p.read_raw) p.read_raw[o] = mine;
      if (p.read_u8) p.read_u8[o] = u;
      if (n < 64) s_read_u8[n * CAM_MAXL + lane] = u;
    }
  }

  // ---- phase 3: per-sequence map: weights[c] = mean_{n,l} dA, conv[c,l] = mean_n A; cam[l] = sum_c w[c]*conv[c,l] --
  float term[CAM_MAXL];
#pragma unroll
  for (int l = 0; l < CAM_MAXL; ++l) term[l] = 0.f;
  if (p.seq_raw || p.seq_u8 || p.seq_resized) {
    for (int c = tid; c < F; c += CAM_THREADS) {
      float gs = 0.f, ca[CAM_MAXL];
#pragma unroll
      for (int l = 0; l < CAM_MAXL; ++l) ca[l] = 0.f;
      for (int n = 0; n < group; ++n) {
        const float w7 = __fdiv_rn(wt[n * F + c], invL);
        const uint32_t bits = posmask[n * F + c];
#pragma unroll
        for (int l = 0; l < CAM_MAXL; ++l)
          if (l < L) {
            gs += ((bits >> l) & 1u) ? w7 : 0.f;
            ca[l] += Elem<T>::ld(a_seq + (size_t)(n * L + l) * p.a_stride + c);
          }
      }
      const float wc = __fdiv_rn(gs, (float)(group * L));
#pragma unroll
      for (int l = 0; l < CAM_MAXL; ++l) term[l] = fmaf(wc, __fdiv_rn(ca[l], (float)group), term[l]);
    }
    __syncthreads();  // `red` is reused
#pragma unroll
    for (int l = 0; l < CAM_MAXL; ++l) {
      const float v = warp_sum(term[l]);
      if (lane == 0) red[warp][l] = v;
    }
    __syncthreads();
    if (warp == 0) {
      float v = 0.f;
      if (lane < L)
        for (int w = 0; w < CAM_WARPS; ++w) v += red[w][lane];
      const uint8_t u = cam_normalize_lane(v, lane, L);
      if (lane < L) {
        if (p.seq_raw) p.seq_raw[(size_t)seq * L + lane] = v;
        if (p.seq_u8) p.seq_u8[(size_t)seq * L + lane] = u;
        s_seq_u8[lane] = u;
      }
    }
  }
  __syncthreads();

  // ---- phase 4: cv2.resize of every uint8 row to resized_len samples ------------------------------------------
  if (p.resized_len > 0) {
    const float scale = (float)L / (float)p.resized_len;
    if (p.read_resized)
      for (int i = tid; i < group * p.resized_len; i += CAM_THREADS) {
        const int n = i / p.resized_len, d = i - n * p.resized_len;
        p.read_resized[(size_t)(seq * group + n) * p.resized_len + d] = cam_resize_at(s_read_u8 + n * CAM_MAXL, L, d, scale);
      }
    if (p.seq_resized)
      for (int d = tid; d < p.resized_len; d += CAM_THREADS)
        p.seq_resized[(size_t)seq * p.resized_len + d] = cam_resize_at(s_seq_u8, L, d, scale);
  }
}

int launch_gradcam(const dards_gradcam_desc& p, cudaStream_t st) {
  const int n_groups = p.n_groups, dtype = p.dtype;
  DARDS_CHECK_ARG(p.l >= 1 && p.l <= CAM_MAXL, "gradcam: map length must be in [1,%d]", CAM_MAXL);
  DARDS_CHECK_ARG(p.n_out >= 1 && p.n_out <= CAM_MAX_OUT, "gradcam: n_out must be in [1,%d]", CAM_MAX_OUT);
  DARDS_CHECK_ARG(p.group >= 1 && p.group <= 64 && p.f >= 1 && (long long)p.group * p.f <= CAM_MAX_GF,
                  "gradcam: group must be in [1,64] and group*F <= %d", CAM_MAX_GF);
  DARDS_CHECK_ARG(p.a && p.w && p.bias && p.logits, "gradcam: null pointer");
  DARDS_CHECK_ARG(p.a_stride >= p.f, "gradcam: row stride smaller than the channel count");
  if (n_groups == 0) return DARDS_OK;
  const size_t smem = (size_t)p.group * p.f;
  DARDS_DISPATCH_DTYPE(dtype, { gradcam_kernel<T><<<n_groups, CAM_THREADS, smem, st>>>(p); })
  DARDS_CHECK_LAUNCH("gradcam");
  return DARDS_OK;
}

// ---- input contract: (data - mu) / std in float64, one rounding to float32 (dataset.py:1375-1379) -----------------
template <typename R>
__global__ void scale_windows_kernel(const R* __restrict__ raw, float* __restrict__ out, long long n, double mu,
                                     double std, int padded) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = (double)raw[i];
    // padded_breath_by_breath: the zero padding stays zero (dataset.py:1375-1377, 1406-1409)
    const double m = (padded && v == 0.0) ? 0.0 : mu;
    out[i] = (float)__ddiv_rn(__dsub_rn(v, m), std);
  }
}

int launch_scale_windows(const void* raw, int raw_f64, float* out, long long n, double mu, double std, int padded,
                         cudaStream_t st) {
  DARDS_CHECK_ARG(std != 0.0, "scale_windows: std must not be 0");
  if (n == 0) return DARDS_OK;
  DARDS_CHECK_ARG(raw && out, "scale_windows: null pointer");
  long long b = (n + 255) / 256;
  if (b > 148LL * 8) b = 148LL * 8;
  if (raw_f64)
    scale_windows_kernel<double><<<(int)b, 256, 0, st>>>(static_cast<const double*>(raw), out, n, mu, std, padded);
  else
    scale_windows_kernel<float><<<(int)b, 256, 0, st>>>(static_cast<const float*>(raw), out, n, mu, std, padded);
  DARDS_CHECK_LAUNCH("scale_windows");
  return DARDS_OK;
}

}  // namespace dards
