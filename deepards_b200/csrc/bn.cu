// Grouped BatchNorm1d (+residual, +ReLU) forward/backward on channels-last activations.
//
// Bandwidth-bound kernels: one CTA owns (one group of `rows` = group*L rows) x (8 x VEC channels), VEC = the
// number of elements in a 16-byte access (4 fp32 / 8 bf16).  The group tile (<= 1120 x 128 bytes for a 20-breath
// sequence) is streamed once from HBM and re-read from L1/L2 for the second and third sweep.  Statistics use
// the two-sweep (mean, then centred sum of squares) formulation for fp32-faithful variance.
//
// thread layout: 256 threads = 32 row lanes x 8 channel vectors.
#include "common.cuh"

namespace dards {

constexpr int BN_THREADS = 256;
constexpr int BN_LANES = 32;
constexpr int BN_QUADS = 8;
constexpr int BN_MAXCT = BN_QUADS * 8;  // 64 channels per CTA for bf16, 32 for fp32

template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    float4 r = *reinterpret_cast<const float4*>(p);
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      float2 f = __bfloat1622float2(b);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&b);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// reduce v[V] over the 32 row lanes; the total for channel (cq*V + j) is returned to every thread of vector cq
template <int V>
__device__ __forceinline__ void lane_reduce(float (&v)[V], float (*red)[BN_MAXCT + 1], float* bcast, int rl, int cq) {
  __syncthreads();  // protect red/bcast from the previous use
#pragma unroll
  for (int j = 0; j < V; ++j) red[rl][cq * V + j] = v[j];
  __syncthreads();
  if (threadIdx.x < BN_QUADS * V) {
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < BN_LANES; ++r) s += red[r][threadIdx.x];
    bcast[threadIdx.x] = s;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < V; ++j) v[j] = bcast[cq * V + j];
}

template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
    gbn_fwd_kernel(const T* x, T* out, const T* res, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ save_mean, float* __restrict__ save_rstd, int rows, int c, int x_stride,
                   int out_stride, int res_stride, float eps, int relu) {
  constexpr int V = Vec<T>::N;
  __shared__ float red[BN_LANES][BN_MAXCT + 1];
  __shared__ float bcast[BN_MAXCT];
  const int g = blockIdx.y;
  const int rl = threadIdx.x >> 3, cq = threadIdx.x & 7;
  const int c0 = blockIdx.x * (BN_QUADS * V) + cq * V;
  const bool active = c0 < c;
  const size_t row_base = (size_t)g * rows;
  const float inv_n = 1.f / (float)rows;

  float s[V];
#pragma unroll
  for (int j = 0; j < V; ++j) s[j] = 0.f;
  if (active)
    for (int r = rl; r < rows; r += BN_LANES) {
      float v[V];
      Vec<T>::ld(x + (row_base + r) * x_stride + c0, v);
#pragma unroll
      for (int j = 0; j < V; ++j) s[j] += v[j];
    }
  lane_reduce<V>(s, red, bcast, rl, cq);
  float mean[V];
#pragma unroll
  for (int j = 0; j < V; ++j) mean[j] = s[j] * inv_n;

  float q[V];
#pragma unroll
  for (int j = 0; j < V; ++j) q[j] = 0.f;
  if (active)
    for (int r = rl; r < rows; r += BN_LANES) {
      float v[V];
      Vec<T>::ld(x + (row_base + r) * x_stride + c0, v);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float d = v[j] - mean[j];
        q[j] = fmaf(d, d, q[j]);
      }
    }
  lane_reduce<V>(q, red, bcast, rl, cq);
  if (!active) return;
  float sc[V], sh[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const float var = q[j] * inv_n + eps;
    float rstd = rsqrtf(var);
    rstd = rstd * (1.5f - 0.5f * var * rstd * rstd);  // one Newton step: the reference divides by sqrt()
    sc[j] = rstd * gamma[c0 + j];
    sh[j] = beta[c0 + j] - mean[j] * sc[j];
    if (rl == 0) {
      save_mean[(size_t)g * c + c0 + j] = mean[j];
      save_rstd[(size_t)g * c + c0 + j] = rstd;
    }
  }
  for (int r = rl; r < rows; r += BN_LANES) {
    float v[V];
    Vec<T>::ld(x + (row_base + r) * x_stride + c0, v);
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
    if (res) {
      float e[V];
      Vec<T>::ld(res + (row_base + r) * res_stride + c0, e);
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] += e[j];
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    Vec<T>::st(out + (row_base + r) * out_stride + c0, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
    gbn_bwd_kernel(const T* dout, const T* x, const T* mask_src, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                   T* dx, int accumulate_dx, T* dres, float* __restrict__ dgamma_part, float* __restrict__ dbeta_part,
                   int rows, int c, int dout_stride, int x_stride, int mask_stride, int dx_stride, int dres_stride,
                   int relu_mode) {
  constexpr int V = Vec<T>::N;
  __shared__ float red[BN_LANES][BN_MAXCT + 1];
  __shared__ float bcast[BN_MAXCT];
  const int g = blockIdx.y;
  const int rl = threadIdx.x >> 3, cq = threadIdx.x & 7;
  const int c0 = blockIdx.x * (BN_QUADS * V) + cq * V;
  const bool active = c0 < c;
  const size_t row_base = (size_t)g * rows;
  const float inv_n = 1.f / (float)rows;

  float mean[V], rstd[V], sc[V], sh[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    mean[j] = active ? save_mean[(size_t)g * c + c0 + j] : 0.f;
    rstd[j] = active ? save_rstd[(size_t)g * c + c0 + j] : 0.f;
    const float gm = active ? gamma[c0 + j] : 0.f, bt = active ? beta[c0 + j] : 0.f;
    sc[j] = rstd[j] * gm;          // same arithmetic as the forward: y = fmaf(x, sc, sh)
    sh[j] = bt - mean[j] * sc[j];
  }
  // masked upstream gradient and xhat for one row
  auto load_row = [&](int r, float (&gv)[V], float (&xh)[V]) {
    float v[V];
    Vec<T>::ld(dout + (row_base + r) * dout_stride + c0, gv);
    Vec<T>::ld(x + (row_base + r) * x_stride + c0, v);
#pragma unroll
    for (int j = 0; j < V; ++j) xh[j] = (v[j] - mean[j]) * rstd[j];
    if (relu_mode == 1) {
#pragma unroll
      for (int j = 0; j < V; ++j)
        if (!(fmaf(v[j], sc[j], sh[j]) > 0.f)) gv[j] = 0.f;
    } else if (relu_mode == 2) {
      float m[V];
      Vec<T>::ld(mask_src + (row_base + r) * mask_stride + c0, m);
#pragma unroll
      for (int j = 0; j < V; ++j)
        if (!(m[j] > 0.f)) gv[j] = 0.f;
    }
  };

  float s1[V], s2[V];
#pragma unroll
  for (int j = 0; j < V; ++j) s1[j] = s2[j] = 0.f;
  if (active)
    for (int r = rl; r < rows; r += BN_LANES) {
      float gv[V], xh[V];
      load_row(r, gv, xh);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        s1[j] += gv[j];
        s2[j] = fmaf(gv[j], xh[j], s2[j]);
      }
    }
  lane_reduce<V>(s1, red, bcast, rl, cq);
  lane_reduce<V>(s2, red, bcast, rl, cq);
  if (!active) return;
  if (rl == 0) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (dbeta_part) dbeta_part[(size_t)g * c + c0 + j] = s1[j];
      if (dgamma_part) dgamma_part[(size_t)g * c + c0 + j] = s2[j];
    }
  }
  float m1[V], m2[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    m1[j] = s1[j] * inv_n;
    m2[j] = s2[j] * inv_n;
  }
  for (int r = rl; r < rows; r += BN_LANES) {
    float gv[V], xh[V], o[V];
    load_row(r, gv, xh);
#pragma unroll
    for (int j = 0; j < V; ++j) o[j] = sc[j] * (gv[j] - m1[j] - xh[j] * m2[j]);
    if (dres) Vec<T>::st(dres + (row_base + r) * dres_stride + c0, gv);
    T* dst = dx + (row_base + r) * dx_stride + c0;
    if (accumulate_dx) {
      float e[V];
      Vec<T>::ld(dst, e);
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] += e[j];
    }
    Vec<T>::st(dst, o);
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
static int vec_of(int dtype) { return dtype == DARDS_BF16 ? 8 : 4; }

int launch_gbn_fwd(const void* x, void* out, const void* res, const float* gamma, const float* beta, float* save_mean,
                   float* save_rstd, int n_groups, int rows, int c, int x_stride, int out_stride, int res_stride,
                   float eps, int relu, int dtype, cudaStream_t st) {
  const int v = vec_of(dtype);
  DARDS_CHECK_ARG(c % v == 0 && x_stride % v == 0 && out_stride % v == 0 && (!res || res_stride % v == 0),
                  "gbn_fwd: channels and strides must be multiples of %d", v);
  DARDS_CHECK_ARG(rows > 0, "gbn_fwd: empty group");
  if (n_groups == 0) return DARDS_OK;
  DARDS_CHECK_ARG(n_groups <= 65535, "gbn_fwd: too many groups (%d)", n_groups);
  dim3 grid(ceil_div(c, BN_QUADS * v), n_groups);
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
  const int v = vec_of(dtype);
  DARDS_CHECK_ARG(c % v == 0 && dout_stride % v == 0 && x_stride % v == 0 && dx_stride % v == 0,
                  "gbn_bwd: channels and strides must be multiples of %d", v);
  DARDS_CHECK_ARG(relu_mode != 2 || (mask_src && mask_stride % v == 0), "gbn_bwd: relu_mode 2 needs mask_src");
  DARDS_CHECK_ARG(!dres || dres_stride % v == 0, "gbn_bwd: dres stride");
  if (n_groups == 0) return DARDS_OK;
  DARDS_CHECK_ARG(n_groups <= 65535, "gbn_bwd: too many groups (%d)", n_groups);
  dim3 grid(ceil_div(c, BN_QUADS * v), n_groups);
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
