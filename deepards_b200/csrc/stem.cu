// Stem: Conv1d(1->C0, k7, s2, p3) + grouped BatchNorm + ReLU + Max/AvgPool1d(3,2,1), fused.
//
// The 7-tap, single-input-channel convolution is far too thin for an MMA (K = 7), and its output
// (N,112,C0) is the largest activation of the whole network.  So it is never materialised: a CTA owns one
// BatchNorm group x 16 output channels, keeps the group's input (group x 224 fp32, zero-padded by 8 on both
// sides: 19 KB for a 20-breath sequence) in shared memory and recomputes the convolution where needed:
//   forward : mean from the 7 tap-input sums (channel independent), sweep 1 centred variance,
//             sweep 2 conv+BN+ReLU+pool -> (N,56,C0)
//   backward: from x, the saved statistics and d(pool out) only -- pool arg-max, ReLU mask, BN backward
//             and the weight gradient are all recomputed; there is no gradient w.r.t. the input.
// Every sweep walks RUNS of 28 consecutive conv outputs (= 14 pool outputs) of one breath with the 7-tap
// input window held in registers: two new inputs (one 8-byte shared-memory load, a broadcast within the
// 16 channel threads) and 7 FMAs per conv output -- the kernels are FMA-bound, not LDS-bound.
// HBM traffic: x once per channel slab (L2-resident after the first), the pooled output once.
// thread layout: 256 threads = 16 channels x 16 row lanes.
#include "common.cuh"

namespace dards {

constexpr int STEM_THREADS = 256;
constexpr int STEM_CS = 16;                          // channels per CTA
constexpr int STEM_LANES = STEM_THREADS / STEM_CS;   // 16 row lanes
constexpr int STEM_L = 224, STEM_LC = 112, STEM_LP = 56, STEM_K = 7;
constexpr int STEM_PAD = 8;                          // zero floats before and after every breath in smem
constexpr int STEM_BS = STEM_L + 2 * STEM_PAD;       // 240: smem breath stride (even -> 8-byte aligned pairs)
constexpr int STEM_RUN = 28;                         // conv outputs per run; 4 runs per breath
constexpr int STEM_MAX_GROUP = 236;                  // 236 * 240 * 4 B = 226 KB of dynamic smem

// 16-byte asynchronous global -> shared copy (L2-only caching: the data is read once)
__device__ __forceinline__ void stem_cp_async16(const void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void stem_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// group input -> zero-padded shared copy
__device__ __forceinline__ void stem_load_group(const float* __restrict__ xg, float* xs, int group) {
  for (int i = threadIdx.x; i < group * STEM_BS; i += STEM_THREADS) {
    const int b = i / STEM_BS, p = i % STEM_BS - STEM_PAD;
    xs[i] = (p >= 0 && p < STEM_L) ? xg[b * STEM_L + p] : 0.f;
  }
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

__device__ __forceinline__ float stem_dot7(const float (&w)[STEM_K], float x0, float x1, float x2, float x3, float x4,
                                           float x5, float x6) {
  float y = w[0] * x0;
  y = fmaf(w[1], x1, y);
  y = fmaf(w[2], x2, y);
  y = fmaf(w[3], x3, y);
  y = fmaf(w[4], x4, y);
  y = fmaf(w[5], x5, y);
  y = fmaf(w[6], x6, y);
  return y;
}

// X_t = sum over all conv positions of the t-th tap input (channel independent) -> xt[0..6], all threads
__device__ __forceinline__ void stem_tap_sums(const float* xs, int group, float* red /* 8*7 */, float* xt) {
  float acc[STEM_K];
#pragma unroll
  for (int t = 0; t < STEM_K; ++t) acc[t] = 0.f;
  const int n_conv = group * STEM_LC;
  for (int i = threadIdx.x; i < n_conv; i += STEM_THREADS) {
    const float* xb = xs + (i / STEM_LC) * STEM_BS + STEM_PAD + 2 * (i % STEM_LC) - 3;
#pragma unroll
    for (int t = 0; t < STEM_K; ++t) acc[t] += xb[t];
  }
#pragma unroll
  for (int t = 0; t < STEM_K; ++t) acc[t] = warp_sum(acc[t]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int t = 0; t < STEM_K; ++t) red[warp * STEM_K + t] = acc[t];
  }
  __syncthreads();
  if (threadIdx.x < STEM_K) {
    float s = 0.f;
    for (int wv = 0; wv < STEM_THREADS / 32; ++wv) s += red[wv * STEM_K + threadIdx.x];
    xt[threadIdx.x] = s;
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(STEM_THREADS)
    stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ gamma,
                    const float* __restrict__ beta, T* __restrict__ out, float* __restrict__ save_mean,
                    float* __restrict__ save_rstd, int group, int c0, int out_stride, float eps, int pool) {
  extern __shared__ __align__(16) float xs[];
  __shared__ float red[STEM_THREADS];
  __shared__ float xt[STEM_K];
  const int g = blockIdx.x;
  const int c = threadIdx.x % STEM_CS, rl = threadIdx.x / STEM_CS;
  const int ch = blockIdx.y * STEM_CS + c;  // c0 is a multiple of 16
  stem_load_group(x + (size_t)g * group * STEM_L, xs, group);
  __syncthreads();
  float wr[STEM_K];
#pragma unroll
  for (int t = 0; t < STEM_K; ++t) wr[t] = w[ch * STEM_K + t];

  const float inv_n = 1.f / (float)(group * STEM_LC);
  stem_tap_sums(xs, group, red, xt);
  float mean = 0.f;
#pragma unroll
  for (int t = 0; t < STEM_K; ++t) mean = fmaf(wr[t], xt[t], mean);
  mean *= inv_n;

  const int n_runs = group * (STEM_LC / STEM_RUN);
  // ---- sweep 1: centred sum of squares ------------------------------------------------------------------
  float q = 0.f;
  for (int run = rl; run < n_runs; run += STEM_LANES) {
    const float* xb = xs + (run >> 2) * STEM_BS + STEM_PAD + 2 * (run & 3) * STEM_RUN - 3;
    float x0 = xb[0], x1 = xb[1], x2 = xb[2], x3 = xb[3], x4 = xb[4], x5 = xb[5], x6 = xb[6];
#pragma unroll
    for (int j = 0; j < STEM_RUN; ++j) {
      const float d = stem_dot7(wr, x0, x1, x2, x3, x4, x5, x6) - mean;
      q = fmaf(d, d, q);
      const float2 nx = *reinterpret_cast<const float2*>(xb + 7 + 2 * j);  // absolute index is even
      x0 = x2; x1 = x3; x2 = x4; x3 = x5; x4 = x6; x5 = nx.x; x6 = nx.y;
    }
  }
  const float var = stem_lane_sum(q, red, c, rl) * inv_n + eps;
  float rstd = rsqrtf(var);
  rstd = rstd * (1.5f - 0.5f * var * rstd * rstd);
  if (rl == 0) {
    save_mean[(size_t)g * c0 + ch] = mean;
    save_rstd[(size_t)g * c0 + ch] = rstd;
  }
  const float sc = rstd * gamma[ch], sh = beta[ch] - mean * sc;

  // ---- sweep 2: conv + BN + ReLU + pool; run = 14 pool outputs = conv positions 2*lp0-1 .. 2*lp0+27 ----------
  for (int run = rl; run < n_runs; run += STEM_LANES) {
    const int b = run >> 2, lp0 = (run & 3) * (STEM_RUN / 2);
    const float* xb = xs + b * STEM_BS + STEM_PAD + 2 * (2 * lp0 - 1) - 3;
    float x0 = xb[0], x1 = xb[1], x2 = xb[2], x3 = xb[3], x4 = xb[4], x5 = xb[5], x6 = xb[6];
    // ReLU output >= 0, so 0 is the identity of the max as well as of the (count_include_pad) sum
    float zl = lp0 == 0 ? 0.f : fmaxf(fmaf(stem_dot7(wr, x0, x1, x2, x3, x4, x5, x6), sc, sh), 0.f);
    T* op = out + ((size_t)(g * group + b) * STEM_LP + lp0) * out_stride + ch;
#pragma unroll
    for (int j = 0; j < STEM_RUN / 2; ++j) {
      float2 nx = *reinterpret_cast<const float2*>(xb + 7 + 4 * j);
      x0 = x2; x1 = x3; x2 = x4; x3 = x5; x4 = x6; x5 = nx.x; x6 = nx.y;
      const float zm = fmaxf(fmaf(stem_dot7(wr, x0, x1, x2, x3, x4, x5, x6), sc, sh), 0.f);
      nx = *reinterpret_cast<const float2*>(xb + 9 + 4 * j);
      x0 = x2; x1 = x3; x2 = x4; x3 = x5; x4 = x6; x5 = nx.x; x6 = nx.y;
      const float zr = fmaxf(fmaf(stem_dot7(wr, x0, x1, x2, x3, x4, x5, x6), sc, sh), 0.f);
      const float o = pool == 0 ? fmaxf(fmaxf(zl, zm), zr) : (zl + zm + zr) * (1.f / 3.f);
      Elem<T>::st(op + (size_t)j * out_stride, o);
      zl = zr;
    }
  }
}

// number of distinct entries of the symmetric 7x7 autocorrelation matrix
constexpr int STEM_NR = STEM_K * (STEM_K + 1) / 2;  // 28
__device__ __forceinline__ int stem_r_index(int a, int b) {  // a <= b
  return a * STEM_K - a * (a - 1) / 2 + (b - a);
}

// STAGE: the CTA's slice of dout (group x 56 rows x 16 channels) is brought into shared memory with cp.async while the
// input moments are computed, and sweep B reads it from there.  With the gradient loaded from global memory inside the
// sweep the kernel spent most of its time waiting on those loads (ncu: long-scoreboard stalls 5.9 per issue, issue
// slots 47 % busy).  Groups too large for the extra 35 KB per 20 breaths use the direct path.
template <typename T, bool STAGE>
__global__ void __launch_bounds__(STEM_THREADS)
    stem_bwd_kernel(const T* __restrict__ dout, const float* __restrict__ x, const float* __restrict__ w,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ save_mean,
                    const float* __restrict__ save_rstd, float* __restrict__ dw_part, float* __restrict__ dgamma_part,
                    float* __restrict__ dbeta_part, int group, int c0, int dout_stride, int pool) {
  extern __shared__ __align__(16) float xs[];
  __shared__ float red[STEM_THREADS * 9];
  __shared__ float xr[STEM_K + STEM_NR];  // X_t (7) then R (28)
  const int g = blockIdx.x;
  const int tid = threadIdx.x;
  const int c = tid % STEM_CS, rl = tid / STEM_CS;
  const int ch = blockIdx.y * STEM_CS + c;
  T* ds = reinterpret_cast<T*>(xs + (size_t)group * STEM_BS);  // [group*56][16] staged dout slice (STAGE only)
  if (STAGE) {
    constexpr int EPC = 16 / (int)sizeof(T);           // elements per 16-byte chunk
    constexpr int CPR = STEM_CS / EPC;                 // chunks per row: 2 (bf16) or 4 (fp32)
    const T* src = dout + (size_t)g * group * STEM_LP * dout_stride + blockIdx.y * STEM_CS;
    const int n_chunks = group * STEM_LP * CPR;
    for (int q = tid; q < n_chunks; q += STEM_THREADS) {
      const int row = q / CPR, part = q % CPR;
      stem_cp_async16(ds + row * STEM_CS + part * EPC, src + (size_t)row * dout_stride + part * EPC);
    }
  }
  stem_load_group(x + (size_t)g * group * STEM_L, xs, group);
  __syncthreads();
  const int n_conv = group * STEM_LC;

  // ---- A. channel-independent input moments: X_t = sum xin_t, R[a][b] = sum xin_a * xin_b --------------
  {
    float acc[STEM_K + STEM_NR];
#pragma unroll
    for (int i = 0; i < STEM_K + STEM_NR; ++i) acc[i] = 0.f;
    for (int i = tid; i < n_conv; i += STEM_THREADS) {
      const float* xb = xs + (i / STEM_LC) * STEM_BS + STEM_PAD + 2 * (i % STEM_LC) - 3;
      float xin[STEM_K];
#pragma unroll
      for (int t = 0; t < STEM_K; ++t) xin[t] = xb[t];
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
  // run = 14 pool outputs; the 11-element input window x[4lp-5 .. 4lp+5] of the three conv positions feeding
  // pool output lp (2lp-1, 2lp, 2lp+1) stays in registers and slides by 4 per pool output.
  float wr[STEM_K];
#pragma unroll
  for (int t = 0; t < STEM_K; ++t) wr[t] = w[ch * STEM_K + t];
  const float mean = save_mean[(size_t)g * c0 + ch], rstd = save_rstd[(size_t)g * c0 + ch];
  const float gm = gamma[ch];
  const float sc = rstd * gm, sh = beta[ch] - mean * sc;
  float s1 = 0.f, s2 = 0.f, gt[STEM_K];
#pragma unroll
  for (int t = 0; t < STEM_K; ++t) gt[t] = 0.f;
  const int n_runs = group * (STEM_LC / STEM_RUN);
  if (STAGE) {
    stem_cp_async_wait_all();
    __syncthreads();  // every thread's chunks have landed
  }
  const int dp_stride = STAGE ? STEM_CS : dout_stride;
  for (int run = rl; run < n_runs; run += STEM_LANES) {
    const int b = run >> 2, lp0 = (run & 3) * (STEM_RUN / 2);
    const float* xb = xs + b * STEM_BS + STEM_PAD + 4 * lp0 - 5;  // xw[i] = xb[i + 4j]
    const T* dp_ptr = STAGE ? ds + (b * STEM_LP + lp0) * STEM_CS + c
                            : dout + ((size_t)(g * group + b) * STEM_LP + lp0) * dout_stride + ch;
    float xw[11];
#pragma unroll
    for (int i = 0; i < 7; ++i) xw[i] = xb[i];
    float yl = stem_dot7(wr, xw[0], xw[1], xw[2], xw[3], xw[4], xw[5], xw[6]);
    bool okl = lp0 != 0;  // conv position -1 does not exist
#pragma unroll
    for (int j = 0; j < STEM_RUN / 2; ++j) {
      const float dp = Elem<T>::ld(dp_ptr + (size_t)j * dp_stride);
      const float2 n0 = *reinterpret_cast<const float2*>(xb + 7 + 4 * j);
      const float2 n1 = *reinterpret_cast<const float2*>(xb + 9 + 4 * j);
      xw[7] = n0.x; xw[8] = n0.y; xw[9] = n1.x; xw[10] = n1.y;
      const float ym = stem_dot7(wr, xw[2], xw[3], xw[4], xw[5], xw[6], xw[7], xw[8]);
      const float yr = stem_dot7(wr, xw[4], xw[5], xw[6], xw[7], xw[8], xw[9], xw[10]);
      const float zl = okl ? fmaxf(fmaf(yl, sc, sh), 0.f) : 0.f;
      const float zm = fmaxf(fmaf(ym, sc, sh), 0.f);
      const float zr = fmaxf(fmaf(yr, sc, sh), 0.f);
      float gl, gmid, gr;
      if (pool == 0) {
        // first strict maximum wins (ATen max_pool semantics); a zero maximum carries no gradient (ReLU)
        const bool wl = zl > 0.f;
        const bool wm = zm > (wl ? zl : 0.f);
        const float best = wm ? zm : (wl ? zl : 0.f);
        const bool wrt = zr > best;
        gl = (wl && !wm && !wrt) ? dp : 0.f;
        gmid = (wm && !wrt) ? dp : 0.f;
        gr = wrt ? dp : 0.f;
      } else {
        const float d3 = dp * (1.f / 3.f);
        gl = zl > 0.f ? d3 : 0.f;
        gmid = zm > 0.f ? d3 : 0.f;
        gr = zr > 0.f ? d3 : 0.f;
      }
      s1 += gl + gmid + gr;
      s2 = fmaf(gl, (yl - mean) * rstd, s2);
      s2 = fmaf(gmid, (ym - mean) * rstd, s2);
      s2 = fmaf(gr, (yr - mean) * rstd, s2);
#pragma unroll
      for (int t = 0; t < STEM_K; ++t) {
        gt[t] = fmaf(gl, xw[t], gt[t]);
        gt[t] = fmaf(gmid, xw[t + 2], gt[t]);
        gt[t] = fmaf(gr, xw[t + 4], gt[t]);
      }
      // slide by 4 inputs = 2 conv positions
#pragma unroll
      for (int i = 0; i < 7; ++i) xw[i] = xw[i + 4];
      yl = yr;
      okl = true;
    }
  }
  // reduce the 9 accumulators over the row lanes
  __syncthreads();
  red[tid * 9 + 0] = s1;
  red[tid * 9 + 1] = s2;
#pragma unroll
  for (int t = 0; t < STEM_K; ++t) red[tid * 9 + 2 + t] = gt[t];
  __syncthreads();
  if (rl == 0) {
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
}

static bool stem_c0_ok(int c0) { return c0 > 0 && c0 % STEM_CS == 0; }

template <typename K>
static int stem_smem_optin(K kernel, size_t smem, size_t* granted) {
  if (smem <= *granted) return DARDS_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("stem: cannot opt in to %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return DARDS_ERR_CUDA;
  }
  *granted = smem;
  return DARDS_OK;
}

int launch_stem_fwd(const float* x, const float* w, const float* gamma, const float* beta, void* out, float* save_mean,
                    float* save_rstd, int n_groups, int group, int c0, int out_stride, float eps, int pool, int dtype,
                    cudaStream_t st) {
  DARDS_CHECK_ARG(stem_c0_ok(c0), "stem: initial planes must be a multiple of %d (got %d)", STEM_CS, c0);
  DARDS_CHECK_ARG(group > 0 && group <= STEM_MAX_GROUP, "stem: BatchNorm group must be in [1, %d] breaths (got %d)",
                  STEM_MAX_GROUP, group);
  if (n_groups == 0) return DARDS_OK;
  DARDS_CHECK_ARG(c0 / STEM_CS <= 65535, "stem: grid too large");
  const size_t smem = (size_t)group * STEM_BS * sizeof(float);
  dim3 grid(n_groups, c0 / STEM_CS);
  static size_t granted[2] = {44 * 1024, 44 * 1024};  // 1 KB of static smem on top
  DARDS_DISPATCH_DTYPE(dtype, {
    int rc = stem_smem_optin(stem_fwd_kernel<T>, smem, &granted[dtype == DARDS_BF16 ? 1 : 0]);
    if (rc) return rc;
    stem_fwd_kernel<T><<<grid, STEM_THREADS, smem, st>>>(x, w, gamma, beta, static_cast<T*>(out), save_mean, save_rstd,
                                                         group, c0, out_stride, eps, pool);
  })
  DARDS_CHECK_LAUNCH("stem_fwd");
  return DARDS_OK;
}

int launch_stem_bwd(const void* dout, const float* x, const float* w, const float* gamma, const float* beta,
                    const float* save_mean, const float* save_rstd, float* dw_part, float* dgamma_part,
                    float* dbeta_part, int n_groups, int group, int c0, int dout_stride, int pool, int dtype,
                    cudaStream_t st) {
  DARDS_CHECK_ARG(stem_c0_ok(c0), "stem: initial planes must be a multiple of %d (got %d)", STEM_CS, c0);
  DARDS_CHECK_ARG(group > 0 && group <= STEM_MAX_GROUP - 10, "stem: BatchNorm group must be in [1, %d] breaths (got %d)",
                  STEM_MAX_GROUP - 10, group);
  if (n_groups == 0) return DARDS_OK;
  const size_t smem_x = (size_t)group * STEM_BS * sizeof(float);
  dim3 grid(n_groups, c0 / STEM_CS);
  static size_t granted[4] = {36 * 1024, 36 * 1024, 36 * 1024, 36 * 1024};  // 9.4 KB of static smem on top
  DARDS_DISPATCH_DTYPE(dtype, {
    // staged gradient slice: 16-byte aligned rows, and room for it next to the input (3 CTAs per SM at group 20)
    const size_t smem_d = (size_t)group * STEM_LP * STEM_CS * sizeof(T);
    const bool stage = (dout_stride * sizeof(T)) % 16 == 0 && (reinterpret_cast<uintptr_t>(dout) & 15) == 0 &&
                       smem_x + smem_d <= 200 * 1024;
    const int slot = (dtype == DARDS_BF16 ? 1 : 0) + (stage ? 2 : 0);
    if (stage) {
      int rc = stem_smem_optin(stem_bwd_kernel<T, true>, smem_x + smem_d, &granted[slot]);
      if (rc) return rc;
      stem_bwd_kernel<T, true><<<grid, STEM_THREADS, smem_x + smem_d, st>>>(
          static_cast<const T*>(dout), x, w, gamma, beta, save_mean, save_rstd, dw_part, dgamma_part, dbeta_part, group,
          c0, dout_stride, pool);
    } else {
      int rc = stem_smem_optin(stem_bwd_kernel<T, false>, smem_x, &granted[slot]);
      if (rc) return rc;
      stem_bwd_kernel<T, false><<<grid, STEM_THREADS, smem_x, st>>>(
          static_cast<const T*>(dout), x, w, gamma, beta, save_mean, save_rstd, dw_part, dgamma_part, dbeta_part, group,
          c0, dout_stride, pool);
    }
  })
  DARDS_CHECK_LAUNCH("stem_bwd");
  return DARDS_OK;
}

}  // namespace dards
