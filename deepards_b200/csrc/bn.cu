// Grouped BatchNorm1d (+residual, +ReLU) forward/backward on channels-last activations.
//
// Bandwidth-bound kernels: one CTA owns (one group of `rows` = group*L rows) x (32 channels).  The group
// tile (<= 2240 x 32 elements) is streamed once from HBM and re-read from L1/L2 for the second and third
// sweep, so HBM traffic is one read of every input and one write of every output.  Statistics use the
// two-sweep (mean, then centred sum of squares) formulation for fp32-faithful variance.
//
// thread layout: 256 threads = 32 row lanes x 8 channel quads (4 consecutive channels, one 16-byte
// (fp32) / 8-byte (bf16) vector access per row).
#include "common.cuh"

namespace dards {

constexpr int BN_CT = 32;        // channels per CTA
constexpr int BN_THREADS = 256;  // 32 row lanes x 8 quads
constexpr int BN_LANES = 32;

// reduce v[4] over the 32 row lanes; result for channel (quad*4+j) is returned to every thread of that quad
__device__ __forceinline__ void lane_reduce4(float (&v)[4], float (*red)[BN_CT + 1], float* bcast, int rl, int cq) {
  __syncthreads();  // protect red/bcast from the previous use
#pragma unroll
  for (int j = 0; j < 4; ++j) red[rl][cq * 4 + j] = v[j];
  __syncthreads();
  if (threadIdx.x < BN_CT) {
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < BN_LANES; ++r) s += red[r][threadIdx.x];
    bcast[threadIdx.x] = s;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = bcast[cq * 4 + j];
}

template <typename T>
__global__ void __launch_bounds__(BN_THREADS) gbn_fwd_kernel(const T* x, T* out, const T* res,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ save_mean,
                                                             float* __restrict__ save_rstd, int rows, int c, int x_stride,
                                                             int out_stride, int res_stride, float eps, int relu) {
  __shared__ float red[BN_LANES][BN_CT + 1];
  __shared__ float bcast[BN_CT];
  const int g = blockIdx.y;
  const int rl = threadIdx.x >> 3, cq = threadIdx.x & 7;
  const int c0 = blockIdx.x * BN_CT + cq * 4;
  const bool active = c0 < c;
  const size_t row_base = (size_t)g * rows;
  const float inv_n = 1.f / (float)rows;

  float s[4] = {0.f, 0.f, 0.f, 0.f};
  if (active)
    for (int r = rl; r < rows; r += BN_LANES) {
      float4 v = Elem<T>::ld4(x + (row_base + r) * x_stride + c0);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    }
  lane_reduce4(s, red, bcast, rl, cq);
  float mean[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) mean[j] = s[j] * inv_n;

  float q[4] = {0.f, 0.f, 0.f, 0.f};
  if (active)
    for (int r = rl; r < rows; r += BN_LANES) {
      float4 v = Elem<T>::ld4(x + (row_base + r) * x_stride + c0);
      float d0 = v.x - mean[0], d1 = v.y - mean[1], d2 = v.z - mean[2], d3 = v.w - mean[3];
      q[0] = fmaf(d0, d0, q[0]); q[1] = fmaf(d1, d1, q[1]); q[2] = fmaf(d2, d2, q[2]); q[3] = fmaf(d3, d3, q[3]);
    }
  lane_reduce4(q, red, bcast, rl, cq);
  if (!active) return;
  float rstd[4], sc[4], sh[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    rstd[j] = rsqrtf(q[j] * inv_n + eps);
    // one Newton step: rsqrtf is 2 ulp; the reference divides by sqrt()
    float v = q[j] * inv_n + eps;
    rstd[j] = rstd[j] * (1.5f - 0.5f * v * rstd[j] * rstd[j]);
    sc[j] = rstd[j] * gamma[c0 + j];
    sh[j] = beta[c0 + j] - mean[j] * sc[j];
  }
  if (rl == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      save_mean[(size_t)g * c + c0 + j] = mean[j];
      save_rstd[(size_t)g * c + c0 + j] = rstd[j];
    }
  }
  for (int r = rl; r < rows; r += BN_LANES) {
    float4 v = Elem<T>::ld4(x + (row_base + r) * x_stride + c0);
    v.x = fmaf(v.x, sc[0], sh[0]); v.y = fmaf(v.y, sc[1], sh[1]);
    v.z = fmaf(v.z, sc[2], sh[2]); v.w = fmaf(v.w, sc[3], sh[3]);
    if (res) {
      float4 e = Elem<T>::ld4(res + (row_base + r) * res_stride + c0);
      v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
    }
    if (relu) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    }
    Elem<T>::st4(out + (row_base + r) * out_stride + c0, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
    gbn_bwd_kernel(const T* dout, const T* x, const T* mask_src,
                   const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ save_mean,
                   const float* __restrict__ save_rstd, T* dx, int accumulate_dx, T* dres, float* __restrict__ dgamma_part,
                   float* __restrict__ dbeta_part, int rows, int c, int dout_stride, int x_stride, int mask_stride,
                   int dx_stride, int dres_stride, int relu_mode) {
  __shared__ float red[BN_LANES][BN_CT + 1];
  __shared__ float bcast[BN_CT];
  const int g = blockIdx.y;
  const int rl = threadIdx.x >> 3, cq = threadIdx.x & 7;
  const int c0 = blockIdx.x * BN_CT + cq * 4;
  const bool active = c0 < c;
  const size_t row_base = (size_t)g * rows;
  const float inv_n = 1.f / (float)rows;

  float mean[4] = {0, 0, 0, 0}, rstd[4] = {0, 0, 0, 0}, gm[4] = {0, 0, 0, 0}, bt[4] = {0, 0, 0, 0};
  if (active) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mean[j] = save_mean[(size_t)g * c + c0 + j];
      rstd[j] = save_rstd[(size_t)g * c + c0 + j];
      gm[j] = gamma[c0 + j];
      bt[j] = beta[c0 + j];
    }
  }
  // masked upstream gradient and xhat for one row
  auto load_row = [&](int r, float (&gv)[4], float (&xh)[4]) {
    float4 d = Elem<T>::ld4(dout + (row_base + r) * dout_stride + c0);
    float4 v = Elem<T>::ld4(x + (row_base + r) * x_stride + c0);
    gv[0] = d.x; gv[1] = d.y; gv[2] = d.z; gv[3] = d.w;
    xh[0] = (v.x - mean[0]) * rstd[0]; xh[1] = (v.y - mean[1]) * rstd[1];
    xh[2] = (v.z - mean[2]) * rstd[2]; xh[3] = (v.w - mean[3]) * rstd[3];
    if (relu_mode == 1) {
      // same arithmetic as the forward: fmaf(x, sc, sh) with sc = rstd*gamma, sh = beta - mean*sc
      const float xin[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float sc = rstd[j] * gm[j];
        float y = fmaf(xin[j], sc, bt[j] - mean[j] * sc);
        if (!(y > 0.f)) gv[j] = 0.f;
      }
    } else if (relu_mode == 2) {
      float4 m = Elem<T>::ld4(mask_src + (row_base + r) * mask_stride + c0);
      if (!(m.x > 0.f)) gv[0] = 0.f;
      if (!(m.y > 0.f)) gv[1] = 0.f;
      if (!(m.z > 0.f)) gv[2] = 0.f;
      if (!(m.w > 0.f)) gv[3] = 0.f;
    }
  };

  float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
  if (active)
    for (int r = rl; r < rows; r += BN_LANES) {
      float gv[4], xh[4];
      load_row(r, gv, xh);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s1[j] += gv[j];
        s2[j] = fmaf(gv[j], xh[j], s2[j]);
      }
    }
  lane_reduce4(s1, red, bcast, rl, cq);
  lane_reduce4(s2, red, bcast, rl, cq);
  if (!active) return;
  if (rl == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (dbeta_part) dbeta_part[(size_t)g * c + c0 + j] = s1[j];
      if (dgamma_part) dgamma_part[(size_t)g * c + c0 + j] = s2[j];
    }
  }
  float k0[4], m1[4], m2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    k0[j] = gm[j] * rstd[j];
    m1[j] = s1[j] * inv_n;
    m2[j] = s2[j] * inv_n;
  }
  for (int r = rl; r < rows; r += BN_LANES) {
    float gv[4], xh[4];
    load_row(r, gv, xh);
    float4 o;
    o.x = k0[0] * (gv[0] - m1[0] - xh[0] * m2[0]);
    o.y = k0[1] * (gv[1] - m1[1] - xh[1] * m2[1]);
    o.z = k0[2] * (gv[2] - m1[2] - xh[2] * m2[2]);
    o.w = k0[3] * (gv[3] - m1[3] - xh[3] * m2[3]);
    if (dres) Elem<T>::st4(dres + (row_base + r) * dres_stride + c0, make_float4(gv[0], gv[1], gv[2], gv[3]));
    T* dst = dx + (row_base + r) * dx_stride + c0;
    if (accumulate_dx) {
      float4 e = Elem<T>::ld4(dst);
      o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
    }
    Elem<T>::st4(dst, o);
  }
}

// 256 threads = 32 channels x 8 row lanes; fixed summation order -> deterministic
constexpr int RR_LANES = 8;
__global__ void __launch_bounds__(256) reduce_rows_kernel(const float* __restrict__ part, float* __restrict__ out,
                                                          int rows, int c, int accumulate) {
  __shared__ float red[RR_LANES][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + cl;
  float s = 0.f;
  if (i < c)
    for (int r = rl; r < rows; r += RR_LANES) s += part[(size_t)r * c + i];
  red[rl][cl] = s;
  __syncthreads();
  if (rl == 0 && i < c) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < RR_LANES; ++k) t += red[k][cl];
    out[i] = accumulate ? out[i] + t : t;
  }
}

// nn.BatchNorm1d updates its running statistics once per group, in order:  r <- (1-m) r + m v_g,  g = 0..G-1.
// Closed form (SURVEY.md hard part 6):  r_G = (1-m)^G r_0 + m * sum_g (1-m)^(G-1-g) v_g  -- a weighted reduction,
// done here by 8 row lanes per channel instead of a G-long dependent chain.
__global__ void __launch_bounds__(256)
    bn_running_update_kernel(const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                             float* __restrict__ rm, float* __restrict__ rv, long long* nbt, int n_groups, int rows, int c,
                             float momentum, float eps) {
  __shared__ float red_m[RR_LANES][33], red_v[RR_LANES][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + cl;
  if (blockIdx.x == 0 && threadIdx.x == 0 && nbt) *nbt += n_groups;
  const float unbias = rows > 1 ? (float)rows / (float)(rows - 1) : 1.f;
  const float lg = log2f(1.f - momentum);
  float sm = 0.f, sv = 0.f;
  if (i < c)
    for (int g = rl; g < n_groups; g += RR_LANES) {
      const float w = momentum * exp2f(lg * (float)(n_groups - 1 - g));
      const float r = save_rstd[(size_t)g * c + i];
      sm = fmaf(w, save_mean[(size_t)g * c + i], sm);
      sv = fmaf(w, fmaxf(1.f / (r * r) - eps, 0.f) * unbias, sv);
    }
  red_m[rl][cl] = sm;
  red_v[rl][cl] = sv;
  __syncthreads();
  if (rl == 0 && i < c) {
    float tm = 0.f, tv = 0.f;
#pragma unroll
    for (int k = 0; k < RR_LANES; ++k) {
      tm += red_m[k][cl];
      tv += red_v[k][cl];
    }
    const float decay = exp2f(lg * (float)n_groups);
    rm[i] = decay * rm[i] + tm;
    rv[i] = decay * rv[i] + tv;
  }
}

// ---- host launchers ------------------------------------------------------------------------------
int launch_gbn_fwd(const void* x, void* out, const void* res, const float* gamma, const float* beta, float* save_mean,
                   float* save_rstd, int n_groups, int rows, int c, int x_stride, int out_stride, int res_stride,
                   float eps, int relu, int dtype, cudaStream_t st) {
  DARDS_CHECK_ARG(c % 4 == 0 && x_stride % 4 == 0 && out_stride % 4 == 0 && (!res || res_stride % 4 == 0),
                  "gbn_fwd: channels and strides must be multiples of 4");
  DARDS_CHECK_ARG(rows > 0, "gbn_fwd: empty group");
  if (n_groups == 0) return DARDS_OK;
  DARDS_CHECK_ARG(n_groups <= 65535, "gbn_fwd: too many groups (%d)", n_groups);
  dim3 grid(ceil_div(c, BN_CT), n_groups);
  DARDS_DISPATCH_DTYPE(dtype, {
    gbn_fwd_kernel<T><<<grid, BN_THREADS, 0, st>>>(static_cast<const T*>(x), static_cast<T*>(out),
                                                   static_cast<const T*>(res), gamma, beta, save_mean, save_rstd, rows,
                                                   c, x_stride, out_stride, res_stride, eps, relu);
  })
  DARDS_CHECK_LAUNCH("gbn_fwd");
  return DARDS_OK;
}

int launch_gbn_bwd(const void* dout, const void* x, const void* mask_src, const float* gamma, const float* beta,
                   const float* save_mean, const float* save_rstd, void* dx, int accumulate_dx, void* dres,
                   float* dgamma_part, float* dbeta_part, int n_groups, int rows, int c, int dout_stride, int x_stride,
                   int mask_stride, int dx_stride, int dres_stride, int relu_mode, int dtype, cudaStream_t st) {
  DARDS_CHECK_ARG(c % 4 == 0 && dout_stride % 4 == 0 && x_stride % 4 == 0 && dx_stride % 4 == 0,
                  "gbn_bwd: channels and strides must be multiples of 4");
  DARDS_CHECK_ARG(relu_mode != 2 || (mask_src && mask_stride % 4 == 0), "gbn_bwd: relu_mode 2 needs mask_src");
  DARDS_CHECK_ARG(!dres || dres_stride % 4 == 0, "gbn_bwd: dres stride");
  if (n_groups == 0) return DARDS_OK;
  DARDS_CHECK_ARG(n_groups <= 65535, "gbn_bwd: too many groups (%d)", n_groups);
  dim3 grid(ceil_div(c, BN_CT), n_groups);
  DARDS_DISPATCH_DTYPE(dtype, {
    gbn_bwd_kernel<T><<<grid, BN_THREADS, 0, st>>>(
        static_cast<const T*>(dout), static_cast<const T*>(x), static_cast<const T*>(mask_src), gamma, beta, save_mean,
        save_rstd, static_cast<T*>(dx), accumulate_dx, static_cast<T*>(dres), dgamma_part, dbeta_part, rows, c,
        dout_stride, x_stride, mask_stride, dx_stride, dres_stride, relu_mode);
  })
  DARDS_CHECK_LAUNCH("gbn_bwd");
  return DARDS_OK;
}

int launch_reduce_rows(const float* part, float* out, int rows, int c, int accumulate, cudaStream_t st) {
  if (c == 0) return DARDS_OK;
  reduce_rows_kernel<<<ceil_div(c, 32), 256, 0, st>>>(part, out, rows, c, accumulate);
  DARDS_CHECK_LAUNCH("reduce_rows");
  return DARDS_OK;
}

int launch_bn_running_update(const float* save_mean, const float* save_rstd, float* rm, float* rv, long long* nbt,
                             int n_groups, int rows, int c, float momentum, float eps, cudaStream_t st) {
  bn_running_update_kernel<<<ceil_div(c, 32), 256, 0, st>>>(save_mean, save_rstd, rm, rv, nbt, n_groups, rows, c,
                                                             momentum, eps);
  DARDS_CHECK_LAUNCH("bn_running_update");
  return DARDS_OK;
}

}  // namespace dards
