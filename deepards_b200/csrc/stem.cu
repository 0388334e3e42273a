// Stem: Conv1d(1->C0, k7, s2, p3) + grouped BatchNorm + ReLU + Max/AvgPool1d(3,2,1), fused.
//
// The 7-tap, single-input-channel convolution is far too thin for an MMA (K = 7), and its output
// (N,112,C0) is the largest activation of the whole network.  So it is never materialised: a CTA owns one
// BatchNorm group x 16 output channels, keeps the group's input (group x 224 fp32, 17.5 KB for a 20-breath
// sequence) in shared memory and recomputes the convolution in each sweep (7 FMA per output):
//   forward : sweep 1 mean, sweep 2 centred variance, sweep 3 conv+BN+ReLU+pool -> (N,56,C0)
//   backward: from x, the saved statistics and d(pool out) only -- pool arg-max, ReLU mask, BN backward
//             and the weight gradient are all recomputed; there is no gradient w.r.t. the input.
// HBM traffic: x once per channel slice (L2-resident after the first), the pooled output once.
// thread layout: 256 threads = 16 channels x 16 row lanes.
#include "common.cuh"

namespace dards {

constexpr int STEM_THREADS = 256;
constexpr int STEM_CS = 16;                          // channels per CTA
constexpr int STEM_LANES = STEM_THREADS / STEM_CS;   // 16 row lanes
constexpr int STEM_L = 224, STEM_LC = 112, STEM_LP = 56, STEM_K = 7;
constexpr int STEM_SMEM_MAX_BREATHS = 48;  // 48*224*4 = 43 KB of dynamic smem (under the 48 KB default limit)

__device__ __forceinline__ float stem_conv_at(const float* __restrict__ xb, const float (&w)[STEM_K], int l) {
  // y[l] = sum_t w[t] * x[2l + t - 3], zero outside [0,224)
  float y = 0.f;
  const int base = 2 * l - 3;
#pragma unroll
  for (int t = 0; t < STEM_K; ++t) {
    int p = base + t;
    float xv = (p >= 0 && p < STEM_L) ? xb[p] : 0.f;
    y = fmaf(w[t], xv, y);
  }
  return y;
}

// sum `v` over the 16 row lanes (threads with the same channel); every thread gets the total
__device__ __forceinline__ float stem_lane_sum(float v, float* red, int c, int rl) {
  __syncthreads();
  red[rl * STEM_CS + c] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < STEM_LANES; ++r) s += red[r * STEM_CS + c];
  return s;
}

template <typename T>
__global__ void __launch_bounds__(STEM_THREADS)
    stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ gamma,
                    const float* __restrict__ beta, T* __restrict__ out, float* __restrict__ save_mean,
                    float* __restrict__ save_rstd, int group, int c0, int out_stride, float eps, int pool,
                    int use_smem) {
  extern __shared__ float xs_dyn[];
  __shared__ float red[STEM_THREADS];
  const int g = blockIdx.x;
  const int c = threadIdx.x % STEM_CS, rl = threadIdx.x / STEM_CS;
  const int ch = blockIdx.y * STEM_CS + c;  // c0 is a multiple of 16
  const float* xg = x + (size_t)g * group * STEM_L;
  if (use_smem) {
    for (int i = threadIdx.x; i < group * STEM_L; i += STEM_THREADS) xs_dyn[i] = xg[i];
    __syncthreads();
    xg = xs_dyn;
  }
  float wr[STEM_K];
#pragma unroll
  for (int t = 0; t < STEM_K; ++t) wr[t] = w[ch * STEM_K + t];

  const int n_conv = group * STEM_LC;
  const float inv_n = 1.f / (float)n_conv;
  float s = 0.f;
  for (int i = rl; i < n_conv; i += STEM_LANES) s += stem_conv_at(xg + (i / STEM_LC) * STEM_L, wr, i % STEM_LC);
  const float mean = stem_lane_sum(s, red, c, rl) * inv_n;
  float q = 0.f;
  for (int i = rl; i < n_conv; i += STEM_LANES) {
    float d = stem_conv_at(xg + (i / STEM_LC) * STEM_L, wr, i % STEM_LC) - mean;
    q = fmaf(d, d, q);
  }
  const float var = stem_lane_sum(q, red, c, rl) * inv_n + eps;
  float rstd = rsqrtf(var);
  rstd = rstd * (1.5f - 0.5f * var * rstd * rstd);
  if (rl == 0) {
    save_mean[(size_t)g * c0 + ch] = mean;
    save_rstd[(size_t)g * c0 + ch] = rstd;
  }
  const float sc = rstd * gamma[ch], sh = beta[ch] - mean * sc;
  const int n_pool = group * STEM_LP;
  for (int i = rl; i < n_pool; i += STEM_LANES) {
    const int b = i / STEM_LP, lp = i % STEM_LP;
    const float* xb = xg + b * STEM_L;
    float acc = 0.f;  // ReLU output >= 0, so 0 is the identity of the max as well as of the sum
#pragma unroll
    for (int d = -1; d <= 1; ++d) {
      int l = 2 * lp + d;
      if (l < 0 || l >= STEM_LC) continue;
      float z = fmaxf(fmaf(stem_conv_at(xb, wr, l), sc, sh), 0.f);
      acc = pool == 0 ? fmaxf(acc, z) : acc + z;
    }
    if (pool != 0) acc *= (1.f / 3.f);  // count_include_pad=True
    Elem<T>::st(out + ((size_t)(g * group + b) * STEM_LP + lp) * out_stride + ch, acc);
  }
}

// number of distinct entries of the symmetric 7x7 autocorrelation matrix
constexpr int STEM_NR = STEM_K * (STEM_K + 1) / 2;  // 28
__device__ __forceinline__ int stem_r_index(int a, int b) {  // a <= b
  return a * STEM_K - a * (a - 1) / 2 + (b - a);
}

template <typename T>
__global__ void __launch_bounds__(STEM_THREADS)
    stem_bwd_kernel(const T* __restrict__ dout, const float* __restrict__ x, const float* __restrict__ w,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ save_mean,
                    const float* __restrict__ save_rstd, float* __restrict__ dw_part, float* __restrict__ dgamma_part,
                    float* __restrict__ dbeta_part, int group, int c0, int dout_stride, int pool, int use_smem) {
  extern __shared__ float xs_dyn[];
  __shared__ float red[STEM_THREADS * 9];
  __shared__ float xr[STEM_K + STEM_NR];  // X_t (7) then R (28)
  const int g = blockIdx.x;
  const int tid = threadIdx.x;
  const int c = tid % STEM_CS, rl = tid / STEM_CS;
  const int ch = blockIdx.y * STEM_CS + c;
  const float* xg = x + (size_t)g * group * STEM_L;
  if (use_smem) {
    for (int i = tid; i < group * STEM_L; i += STEM_THREADS) xs_dyn[i] = xg[i];
    __syncthreads();
    xg = xs_dyn;
  }
  const int n_conv = group * STEM_LC;

  // ---- A. channel-independent input moments: X_t = sum xin_t, R[a][b] = sum xin_a * xin_b --------------
  {
    float acc[STEM_K + STEM_NR];
#pragma unroll
    for (int i = 0; i < STEM_K + STEM_NR; ++i) acc[i] = 0.f;
    for (int i = tid; i < n_conv; i += STEM_THREADS) {
      const float* xb = xg + (i / STEM_LC) * STEM_L;
      const int base = 2 * (i % STEM_LC) - 3;
      float xin[STEM_K];
#pragma unroll
      for (int t = 0; t < STEM_K; ++t) {
        int p = base + t;
        xin[t] = (p >= 0 && p < STEM_L) ? xb[p] : 0.f;
      }
      int k = STEM_K;
#pragma unroll
      for (int a = 0; a < STEM_K; ++a) {
        acc[a] += xin[a];
#pragma unroll
        for (int b = a; b < STEM_K; ++b) {
          acc[k] = fmaf(xin[a], xin[b], acc[k]);
          ++k;
        }
      }
    }
    // block reduce: warp shuffle, then 8 warps through smem
#pragma unroll
    for (int i = 0; i < STEM_K + STEM_NR; ++i) acc[i] = warp_sum(acc[i]);
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < STEM_K + STEM_NR; ++i) red[warp * (STEM_K + STEM_NR) + i] = acc[i];
    }
    __syncthreads();
    if (tid < STEM_K + STEM_NR) {
      float s = 0.f;
      for (int wv = 0; wv < STEM_THREADS / 32; ++wv) s += red[wv * (STEM_K + STEM_NR) + tid];
      xr[tid] = s;
    }
    __syncthreads();
  }

  // ---- B. per channel: S1 = sum g, S2 = sum g*xhat, G_t = sum g * xin_t ----------------------------------
  float wr[STEM_K];
#pragma unroll
  for (int t = 0; t < STEM_K; ++t) wr[t] = w[ch * STEM_K + t];
  const float mean = save_mean[(size_t)g * c0 + ch], rstd = save_rstd[(size_t)g * c0 + ch];
  const float gm = gamma[ch];
  const float sc = rstd * gm, sh = beta[ch] - mean * sc;
  float s1 = 0.f, s2 = 0.f, gt[STEM_K];
#pragma unroll
  for (int t = 0; t < STEM_K; ++t) gt[t] = 0.f;
  const int n_pool = group * STEM_LP;
  for (int i = rl; i < n_pool; i += STEM_LANES) {
    const int b = i / STEM_LP, lp = i % STEM_LP;
    const float* xb = xg + b * STEM_L;
    const float dp = Elem<T>::ld(dout + ((size_t)(g * group + b) * STEM_LP + lp) * dout_stride + ch);
    float yv[3], zv[3];
    bool ok[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      int l = 2 * lp + d - 1;
      ok[d] = (l >= 0 && l < STEM_LC);
      yv[d] = ok[d] ? stem_conv_at(xb, wr, l) : 0.f;
      zv[d] = ok[d] ? fmaxf(fmaf(yv[d], sc, sh), 0.f) : 0.f;
    }
    float gsel[3] = {0.f, 0.f, 0.f};
    if (pool == 0) {
      // first strict maximum wins (ATen max_pool semantics); a zero maximum carries no gradient (ReLU)
      int win = -1;
      float best = 0.f;
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (ok[d] && zv[d] > best) {
          best = zv[d];
          win = d;
        }
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (d == win) gsel[d] = dp;
    } else {
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (ok[d] && zv[d] > 0.f) gsel[d] = dp * (1.f / 3.f);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float gv = gsel[d];
      if (gv != 0.f) {
        const int base = 2 * (2 * lp + d - 1) - 3;
        s1 += gv;
        s2 = fmaf(gv, (yv[d] - mean) * rstd, s2);
#pragma unroll
        for (int t = 0; t < STEM_K; ++t) {
          int p = base + t;
          float xv = (p >= 0 && p < STEM_L) ? xb[p] : 0.f;
          gt[t] = fmaf(gv, xv, gt[t]);
        }
      }
    }
  }
  // reduce the 9 accumulators over the row lanes
  __syncthreads();
  red[tid * 9 + 0] = s1;
  red[tid * 9 + 1] = s2;
#pragma unroll
  for (int t = 0; t < STEM_K; ++t) red[tid * 9 + 2 + t] = gt[t];
  __syncthreads();
  if (rl != 0) return;
  s1 = 0.f;
  s2 = 0.f;
#pragma unroll
  for (int t = 0; t < STEM_K; ++t) gt[t] = 0.f;
  for (int r = 0; r < STEM_LANES; ++r) {
    const int o = (r * STEM_CS + c) * 9;
    s1 += red[o + 0];
    s2 += red[o + 1];
#pragma unroll
    for (int t = 0; t < STEM_K; ++t) gt[t] += red[o + 2 + t];
  }
  // ---- C. BN backward folded into the weight gradient ------------------------------------------------
  // dy = gamma*rstd*(g - S1/n - xhat*S2/n);  dW_t = sum dy*xin_t
  //    = gamma*rstd*(G_t - S1/n * X_t - S2/n * H_t),  H_t = sum xhat*xin_t = rstd*(sum_a w_a R[a][t] - mean*X_t)
  const float inv_n = 1.f / (float)n_conv;
  dbeta_part[(size_t)g * c0 + ch] = s1;
  dgamma_part[(size_t)g * c0 + ch] = s2;
#pragma unroll
  for (int t = 0; t < STEM_K; ++t) {
    float wr_r = 0.f;
#pragma unroll
    for (int a = 0; a < STEM_K; ++a) {
      int lo = a < t ? a : t, hi = a < t ? t : a;
      wr_r = fmaf(wr[a], xr[STEM_K + stem_r_index(lo, hi)], wr_r);
    }
    float h = rstd * (wr_r - mean * xr[t]);
    dw_part[((size_t)g * c0 + ch) * STEM_K + t] = sc * (gt[t] - s1 * inv_n * xr[t] - s2 * inv_n * h);
  }
}

static bool stem_c0_ok(int c0) { return c0 > 0 && c0 % STEM_CS == 0; }

int launch_stem_fwd(const float* x, const float* w, const float* gamma, const float* beta, void* out, float* save_mean,
                    float* save_rstd, int n_groups, int group, int c0, int out_stride, float eps, int pool, int dtype,
                    cudaStream_t st) {
  DARDS_CHECK_ARG(stem_c0_ok(c0), "stem: initial planes must be a multiple of %d (got %d)", STEM_CS, c0);
  DARDS_CHECK_ARG(group > 0, "stem: empty group");
  if (n_groups == 0) return DARDS_OK;
  DARDS_CHECK_ARG(n_groups <= 0x7fffffff / 1 && c0 / STEM_CS <= 65535, "stem: grid too large");
  int use_smem = group <= STEM_SMEM_MAX_BREATHS;
  size_t smem = use_smem ? (size_t)group * STEM_L * sizeof(float) : 0;
  dim3 grid(n_groups, c0 / STEM_CS);
  DARDS_DISPATCH_DTYPE(dtype, {
    stem_fwd_kernel<T><<<grid, STEM_THREADS, smem, st>>>(x, w, gamma, beta, static_cast<T*>(out), save_mean, save_rstd,
                                                         group, c0, out_stride, eps, pool, use_smem);
  })
  DARDS_CHECK_LAUNCH("stem_fwd");
  return DARDS_OK;
}

int launch_stem_bwd(const void* dout, const float* x, const float* w, const float* gamma, const float* beta,
                    const float* save_mean, const float* save_rstd, float* dw_part, float* dgamma_part,
                    float* dbeta_part, int n_groups, int group, int c0, int dout_stride, int pool, int dtype,
                    cudaStream_t st) {
  DARDS_CHECK_ARG(stem_c0_ok(c0), "stem: initial planes must be a multiple of %d (got %d)", STEM_CS, c0);
  if (n_groups == 0) return DARDS_OK;
  int use_smem = group <= STEM_SMEM_MAX_BREATHS;
  size_t smem = use_smem ? (size_t)group * STEM_L * sizeof(float) : 0;
  dim3 grid(n_groups, c0 / STEM_CS);
  DARDS_DISPATCH_DTYPE(dtype, {
    stem_bwd_kernel<T><<<grid, STEM_THREADS, smem, st>>>(static_cast<const T*>(dout), x, w, gamma, beta, save_mean,
                                                         save_rstd, dw_part, dgamma_part, dbeta_part, group, c0,
                                                         dout_stride, pool, use_smem);
  })
  DARDS_CHECK_LAUNCH("stem_bwd");
  return DARDS_OK;
}

}  // namespace dards
