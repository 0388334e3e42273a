// Grouped BatchNorm1d (+residual, +ReLU) forward/backward on channels-last activations.
//
// Bandwidth-bound kernels: one CTA owns (one group of `rows` = group*L rows) x (8 x VEC channels), VEC = the
// number of elements in a 16-byte access (4 fp32 / 8 bf16).  The group tile (<= 1120 x 128 bytes for a 20-breath
// sequence) is streamed once from HBM and re-read from L1/L2 for the later sweeps.  Statistics use the
// two-sweep (mean, then centred sum of squares) formulation for fp32-faithful variance.  Every sweep keeps
// BN_UNROLL rows (16-byte loads of every operand) in flight per thread: the kernels are latency-bound otherwise.
//
// Cross-group reductions (running statistics in the forward; dgamma / dbeta in the backward) are done by the
// LAST CTA to finish a channel tile (threadfence + atomic ticket, the counter resets itself): it reads the
// per-group values written by all CTAs and reduces them in a fixed order, so the result is deterministic and
// no extra kernel launch is needed.
//
// thread layout: 256 threads = 32 row lanes x 8 channel vectors.
#include "common.cuh"

namespace dards {

constexpr int BN_THREADS = 256;
constexpr int BN_LANES = 32;
constexpr int BN_QUADS = 8;
constexpr int BN_MAXCT = BN_QUADS * 8;  // 64 channels per CTA for bf16, 32 for fp32
constexpr int BN_UNROLL = 4;

template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void unpack(const uint4& r, float (&v)[4]) {
    v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
  }
  static __device__ __forceinline__ uint4 pack(const float (&v)[4]) {
    return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
  }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void unpack(const uint4& r, float (&v)[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ uint4 pack(const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&b);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};

template <typename T> __device__ __forceinline__ uint4 ld16(const T* p) { return *reinterpret_cast<const uint4*>(p); }
template <typename T> __device__ __forceinline__ void st16(T* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

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
                   float* save_mean, float* save_rstd, int rows, int c, int x_stride, int out_stride, int res_stride,
                   float eps, int relu, float* rm, float* rv, long long* nbt, float momentum, unsigned int* counters) {
  constexpr int V = Vec<T>::N;
  constexpr int U = BN_UNROLL;
  __shared__ float red[BN_LANES][BN_MAXCT + 1];
  __shared__ float bcast[BN_MAXCT];
  __shared__ int last_flag;
  __shared__ float4 scratch4[BN_THREADS];
  const int g = blockIdx.y;
  const int rl = threadIdx.x >> 3, cq = threadIdx.x & 7;
  const int c0 = blockIdx.x * (BN_QUADS * V) + cq * V;
  const bool active = c0 < c;
  const size_t row_base = (size_t)g * rows;
  const float inv_n = 1.f / (float)rows;
  const T* xp = x + row_base * x_stride + c0;

  float s[V];
#pragma unroll
  for (int j = 0; j < V; ++j) s[j] = 0.f;
  if (active) {
    int r = rl;
    for (; r + (U - 1) * BN_LANES < rows; r += U * BN_LANES) {
      uint4 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) raw[u] = ld16(xp + (size_t)(r + u * BN_LANES) * x_stride);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float v[V];
        Vec<T>::unpack(raw[u], v);
#pragma unroll
        for (int j = 0; j < V; ++j) s[j] += v[j];
      }
    }
    for (; r < rows; r += BN_LANES) {
      float v[V];
      Vec<T>::unpack(ld16(xp + (size_t)r * x_stride), v);
#pragma unroll
      for (int j = 0; j < V; ++j) s[j] += v[j];
    }
  }
  lane_reduce<V>(s, red, bcast, rl, cq);
  float mean[V];
#pragma unroll
  for (int j = 0; j < V; ++j) mean[j] = s[j] * inv_n;

  float q[V];
#pragma unroll
  for (int j = 0; j < V; ++j) q[j] = 0.f;
  if (active) {
    int r = rl;
    for (; r + (U - 1) * BN_LANES < rows; r += U * BN_LANES) {
      uint4 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) raw[u] = ld16(xp + (size_t)(r + u * BN_LANES) * x_stride);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float v[V];
        Vec<T>::unpack(raw[u], v);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float d = v[j] - mean[j];
          q[j] = fmaf(d, d, q[j]);
        }
      }
    }
    for (; r < rows; r += BN_LANES) {
      float v[V];
      Vec<T>::unpack(ld16(xp + (size_t)r * x_stride), v);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float d = v[j] - mean[j];
        q[j] = fmaf(d, d, q[j]);
      }
    }
  }
  lane_reduce<V>(q, red, bcast, rl, cq);
  float sc[V], sh[V];
#pragma unroll
  for (int j = 0; j < V; ++j) sc[j] = sh[j] = 0.f;
  if (active) {
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
  }
  // running statistics: the last CTA of this channel tile to get here folds all groups' statistics in (before the
  // output sweep, so the ticket's fence has no bulk stores to wait for)
  if (rm != nullptr && last_cta_arrives(counters + blockIdx.x, gridDim.y, &last_flag)) {
    const int ch0 = blockIdx.x * (BN_QUADS * V);
    const int nch = (c - ch0) < BN_QUADS * V ? (c - ch0) : BN_QUADS * V;
    running_update_tile(save_mean, save_rstd, rm, rv, gridDim.y, rows, c, ch0, nch, momentum, eps, scratch4);
    if (blockIdx.x == 0 && threadIdx.x == 0 && nbt) *nbt += gridDim.y;
  }
  if (active) {
    T* op = out + row_base * out_stride + c0;
    const T* rp = res ? res + row_base * res_stride + c0 : nullptr;
    auto finish = [&](const uint4& rx, const uint4& rr, int r) {
      float v[V];
      Vec<T>::unpack(rx, v);
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
      if (rp) {
        float e[V];
        Vec<T>::unpack(rr, e);
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] += e[j];
      }
      if (relu) {
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      st16(op + (size_t)r * out_stride, Vec<T>::pack(v));
    };
    int r = rl;
    for (; r + (U - 1) * BN_LANES < rows; r += U * BN_LANES) {
      uint4 raw[U], rres[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        raw[u] = ld16(xp + (size_t)(r + u * BN_LANES) * x_stride);
        rres[u] = rp ? ld16(rp + (size_t)(r + u * BN_LANES) * res_stride) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) finish(raw[u], rres[u], r + u * BN_LANES);
    }
    for (; r < rows; r += BN_LANES)
      finish(ld16(xp + (size_t)r * x_stride), rp ? ld16(rp + (size_t)r * res_stride) : make_uint4(0u, 0u, 0u, 0u), r);
  }
}

template <typename T>
__global__ void __launch_bounds__(BN_THREADS, 2)
    gbn_bwd_kernel(const T* dout, const T* x, const T* mask_src, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                   T* dx, int accumulate_dx, T* dres, float* dgamma_part, float* dbeta_part, float* dgamma, float* dbeta,
                   unsigned int* counters, int rows, int c, int dout_stride, int x_stride, int mask_stride, int dx_stride,
                   int dres_stride, int relu_mode) {
  constexpr int V = Vec<T>::N;
  constexpr int U = BN_UNROLL;
  __shared__ float red[BN_LANES][BN_MAXCT + 1];
  __shared__ float bcast[BN_MAXCT];
  __shared__ int last_flag;
  __shared__ float4 scratch4[BN_THREADS];
  const int g = blockIdx.y;
  const int rl = threadIdx.x >> 3, cq = threadIdx.x & 7;
  const int c0 = blockIdx.x * (BN_QUADS * V) + cq * V;
  const bool active = c0 < c;
  const size_t row_base = (size_t)g * rows;
  const float inv_n = 1.f / (float)rows;
  const T* gp = dout + row_base * dout_stride + c0;
  const T* xp = x + row_base * x_stride + c0;
  const T* mp = relu_mode == 2 ? mask_src + row_base * mask_stride + c0 : nullptr;

  float mean[V], rstd[V], sc[V], sh[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    mean[j] = active ? save_mean[(size_t)g * c + c0 + j] : 0.f;
    rstd[j] = active ? save_rstd[(size_t)g * c + c0 + j] : 0.f;
    const float gm = active ? gamma[c0 + j] : 0.f, bt = active ? beta[c0 + j] : 0.f;
    sc[j] = rstd[j] * gm;          // same arithmetic as the forward: y = fmaf(x, sc, sh)
    sh[j] = bt - mean[j] * sc[j];
  }
  // masked upstream gradient (gv) and the raw input (xv) of one row
  auto decode = [&](const uint4& rg, const uint4& rx, const uint4& rm_, float (&gv)[V], float (&xv)[V]) {
    Vec<T>::unpack(rg, gv);
    Vec<T>::unpack(rx, xv);
    if (relu_mode == 1) {
#pragma unroll
      for (int j = 0; j < V; ++j)
        if (!(fmaf(xv[j], sc[j], sh[j]) > 0.f)) gv[j] = 0.f;
    } else if (relu_mode == 2) {
      float m[V];
      Vec<T>::unpack(rm_, m);
#pragma unroll
      for (int j = 0; j < V; ++j)
        if (!(m[j] > 0.f)) gv[j] = 0.f;
    }
  };
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);

  float s1[V], s2[V];
#pragma unroll
  for (int j = 0; j < V; ++j) s1[j] = s2[j] = 0.f;
  if (active) {
    auto acc = [&](const uint4& rg, const uint4& rx, const uint4& rm_) {
      float gv[V], xv[V];
      decode(rg, rx, rm_, gv, xv);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        s1[j] += gv[j];
        s2[j] = fmaf(gv[j], (xv[j] - mean[j]) * rstd[j], s2[j]);
      }
    };
    int r = rl;
    for (; r + (U - 1) * BN_LANES < rows; r += U * BN_LANES) {
      uint4 rg[U], rx[U], rk[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t rr = (size_t)(r + u * BN_LANES);
        rg[u] = ld16(gp + rr * dout_stride);
        rx[u] = ld16(xp + rr * x_stride);
        rk[u] = mp ? ld16(mp + rr * mask_stride) : zero4;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) acc(rg[u], rx[u], rk[u]);
    }
    for (; r < rows; r += BN_LANES)
      acc(ld16(gp + (size_t)r * dout_stride), ld16(xp + (size_t)r * x_stride),
          mp ? ld16(mp + (size_t)r * mask_stride) : zero4);
  }
  lane_reduce<V>(s1, red, bcast, rl, cq);
  lane_reduce<V>(s2, red, bcast, rl, cq);
  if (active && rl == 0) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (dbeta_part) dbeta_part[(size_t)g * c + c0 + j] = s1[j];
      if (dgamma_part) dgamma_part[(size_t)g * c + c0 + j] = s2[j];
    }
  }
  // parameter gradients: the last CTA of this channel tile to get here sums the per-group partials in group order
  // (before the dx sweep, so the ticket's fence has no bulk stores to wait for)
  if (dgamma != nullptr && last_cta_arrives(counters + blockIdx.x, gridDim.y, &last_flag)) {
    const int n_groups = gridDim.y;
    const int ch0 = blockIdx.x * (BN_QUADS * V);
    const int nch = (c - ch0) < BN_QUADS * V ? (c - ch0) : BN_QUADS * V;
    auto one = [](int) { return 1.f; };
    const float4 tg = group_reduce4(dgamma_part, n_groups, c, ch0, nch, one, scratch4);
    const float4 tb = group_reduce4(dbeta_part, n_groups, c, ch0, nch, one, scratch4);
    if (threadIdx.x < (nch >> 2)) {
      reinterpret_cast<float4*>(dgamma + ch0)[threadIdx.x] = tg;
      reinterpret_cast<float4*>(dbeta + ch0)[threadIdx.x] = tb;
    }
  }
  if (active) {
    // dx = sc*(g - m1 - xhat*m2) = sc*g + kb*x + kc
    float kb[V], kc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float m1 = s1[j] * inv_n, m2 = s2[j] * inv_n;
      kb[j] = -sc[j] * m2 * rstd[j];
      kc[j] = -sc[j] * m1 - kb[j] * mean[j];
    }
    T* dxp = dx + row_base * dx_stride + c0;
    T* drp = dres ? dres + row_base * dres_stride + c0 : nullptr;
    auto finish = [&](const uint4& rg, const uint4& rx, const uint4& rm_, const uint4& rold, int r) {
      float gv[V], xv[V], o[V];
      decode(rg, rx, rm_, gv, xv);
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = fmaf(sc[j], gv[j], fmaf(kb[j], xv[j], kc[j]));
      if (drp) st16(drp + (size_t)r * dres_stride, Vec<T>::pack(gv));
      if (accumulate_dx) {
        float e[V];
        Vec<T>::unpack(rold, e);
#pragma unroll
        for (int j = 0; j < V; ++j) o[j] += e[j];
      }
      st16(dxp + (size_t)r * dx_stride, Vec<T>::pack(o));
    };
    int r = rl;
    for (; r + (U - 1) * BN_LANES < rows; r += U * BN_LANES) {
      uint4 rg[U], rx[U], rk[U], ro[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t rr = (size_t)(r + u * BN_LANES);
        rg[u] = ld16(gp + rr * dout_stride);
        rx[u] = ld16(xp + rr * x_stride);
        rk[u] = mp ? ld16(mp + rr * mask_stride) : zero4;
        ro[u] = accumulate_dx ? ld16(dxp + rr * dx_stride) : zero4;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) finish(rg[u], rx[u], rk[u], ro[u], r + u * BN_LANES);
    }
    for (; r < rows; r += BN_LANES)
      finish(ld16(gp + (size_t)r * dout_stride), ld16(xp + (size_t)r * x_stride),
             mp ? ld16(mp + (size_t)r * mask_stride) : zero4, accumulate_dx ? ld16(dxp + (size_t)r * dx_stride) : zero4, r);
  }
}

// out[i] (+)= sum_r part[r][i]: 256 threads = 32 columns x 8 row lanes, fixed summation order -> deterministic
constexpr int RR_LANES = 8;
__global__ void __launch_bounds__(256) reduce_rows_kernel(const float* __restrict__ part, float* __restrict__ out,
                                                          int rows, int c, int accumulate) {
  __shared__ float red[RR_LANES][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + cl;
  float s = 0.f;
  if (i < c) {
#pragma unroll 8
    for (int r = rl; r < rows; r += RR_LANES) s += part[(size_t)r * c + i];
  }
  red[rl][cl] = s;
  __syncthreads();
  if (rl == 0 && i < c) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < RR_LANES; ++k) t += red[k][cl];
    out[i] = accumulate ? out[i] + t : t;
  }
}

// stand-alone running-statistics update (the stem's BatchNorm; gbn_fwd does its own)
__global__ void __launch_bounds__(BN_THREADS)
    bn_running_update_kernel(const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                             float* __restrict__ rm, float* __restrict__ rv, long long* nbt, int n_groups, int rows, int c,
                             float momentum, float eps) {
  __shared__ float4 scratch[BN_THREADS];
  if (blockIdx.x == 0 && threadIdx.x == 0 && nbt) *nbt += n_groups;
  const int ch0 = blockIdx.x * 64;
  const int nch = (c - ch0) < 64 ? (c - ch0) : 64;
  running_update_tile(save_mean, save_rstd, rm, rv, n_groups, rows, c, ch0, nch, momentum, eps, scratch);
}

// ---- host launchers ------------------------------------------------------------------------------
static int vec_of(int dtype) { return dtype == DARDS_BF16 ? 8 : 4; }

int launch_gbn_fwd(const void* x, void* out, const void* res, const float* gamma, const float* beta, float* save_mean,
                   float* save_rstd, int n_groups, int rows, int c, int x_stride, int out_stride, int res_stride,
                   float eps, int relu, float* rm, float* rv, long long* nbt, float momentum, unsigned int* counters,
                   int dtype, cudaStream_t st) {
  const int v = vec_of(dtype);
  DARDS_CHECK_ARG(c % v == 0 && x_stride % v == 0 && out_stride % v == 0 && (!res || res_stride % v == 0),
                  "gbn_fwd: channels and strides must be multiples of %d", v);
  DARDS_CHECK_ARG(rows > 0, "gbn_fwd: empty group");
  DARDS_CHECK_ARG((rm == nullptr) == (rv == nullptr), "gbn_fwd: running_mean and running_var go together");
  DARDS_CHECK_ARG(rm == nullptr || counters != nullptr, "gbn_fwd: the running-statistics update needs sync_counters");
  if (n_groups == 0) return DARDS_OK;
  DARDS_CHECK_ARG(n_groups <= 65535, "gbn_fwd: too many groups (%d)", n_groups);
  dim3 grid(ceil_div(c, BN_QUADS * v), n_groups);
  DARDS_DISPATCH_DTYPE(dtype, {
    gbn_fwd_kernel<T><<<grid, BN_THREADS, 0, st>>>(static_cast<const T*>(x), static_cast<T*>(out),
                                                   static_cast<const T*>(res), gamma, beta, save_mean, save_rstd, rows,
                                                   c, x_stride, out_stride, res_stride, eps, relu, rm, rv, nbt, momentum,
                                                   counters);
  })
  DARDS_CHECK_LAUNCH("gbn_fwd");
  return DARDS_OK;
}

int launch_gbn_bwd(const void* dout, const void* x, const void* mask_src, const float* gamma, const float* beta,
                   const float* save_mean, const float* save_rstd, void* dx, int accumulate_dx, void* dres,
                   float* dgamma_part, float* dbeta_part, float* dgamma, float* dbeta, unsigned int* counters,
                   int n_groups, int rows, int c, int dout_stride, int x_stride, int mask_stride, int dx_stride,
                   int dres_stride, int relu_mode, int dtype, cudaStream_t st) {
  const int v = vec_of(dtype);
  DARDS_CHECK_ARG(c % v == 0 && dout_stride % v == 0 && x_stride % v == 0 && dx_stride % v == 0,
                  "gbn_bwd: channels and strides must be multiples of %d", v);
  DARDS_CHECK_ARG(relu_mode != 2 || (mask_src && mask_stride % v == 0), "gbn_bwd: relu_mode 2 needs mask_src");
  DARDS_CHECK_ARG(!dres || dres_stride % v == 0, "gbn_bwd: dres stride");
  DARDS_CHECK_ARG((dgamma == nullptr) == (dbeta == nullptr), "gbn_bwd: dgamma and dbeta go together");
  DARDS_CHECK_ARG(dgamma == nullptr || (counters && dgamma_part && dbeta_part),
                  "gbn_bwd: the fused dgamma/dbeta reduction needs the partial buffers and sync_counters");
  if (n_groups == 0) return DARDS_OK;
  DARDS_CHECK_ARG(n_groups <= 65535, "gbn_bwd: too many groups (%d)", n_groups);
  dim3 grid(ceil_div(c, BN_QUADS * v), n_groups);
  DARDS_DISPATCH_DTYPE(dtype, {
    gbn_bwd_kernel<T><<<grid, BN_THREADS, 0, st>>>(
        static_cast<const T*>(dout), static_cast<const T*>(x), static_cast<const T*>(mask_src), gamma, beta, save_mean,
        save_rstd, static_cast<T*>(dx), accumulate_dx, static_cast<T*>(dres), dgamma_part, dbeta_part, dgamma, dbeta,
        counters, rows, c, dout_stride, x_stride, mask_stride, dx_stride, dres_stride, relu_mode);
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
  bn_running_update_kernel<<<ceil_div(c, 64), BN_THREADS, 0, st>>>(save_mean, save_rstd, rm, rv, nbt, n_groups, rows, c,
                                                                   momentum, eps);
  DARDS_CHECK_LAUNCH("bn_running_update");
  return DARDS_OK;
}

}  // namespace dards
