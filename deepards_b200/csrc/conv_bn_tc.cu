// Conv1d + grouped BatchNorm1d (+ residual) (+ ReLU) in ONE tcgen05 kernel (bf16 operands, fp32 accumulation in TMEM).
//
// Replaces conv -> bn -> relu [-> += residual -> relu] of resnet.py:27-38 and conv -> norm -> relu of
// densenet.py:25-29 without the separate BatchNorm pass over the convolution output.
//
// The BatchNorm statistics of a channel are sums over ALL positions of a group (= one sequence of `group` breaths),
// so the unit of work is a JOB = (group, tile of 128 output channels), never a tile that straddles two groups:
//   * an epilogue thread owns ONE output channel (its TMEM lane) and sees every position of the job's accumulators:
//     sum(y) and sum(y^2) are in-thread sums over the tcgen05.ld values, taken from the fp32 accumulators (shifted by
//     a sample of the data, so E[d^2] - E[d]^2 loses no digits), before the rounding to bf16;
//   * FUSED (the whole group's bf16 output fits the shared-memory "park": 20 breaths x L <= 14 positions x 128
//     channels <= 72 KB, ResNet-18 layers 3-4 / DenseNet blocks 3-4): the job runs its group/nb sub-tiles through the
//     2-deep TMEM ring into the park, the two epilogue threads of a channel merge their moments (Chan's formula),
//     mean / rstd / gamma / beta become one scale and shift per channel, the raw y leaves with ONE TMA store (the
//     backward needs it) and a second sweep over the park writes out = relu(y * scale + shift + res) with 16-byte
//     coalesced stores.  No BatchNorm launch, y is never re-read from memory;
//   * PARTIAL (long sequences, L >= 28: the group is several tiles of different CTAs): every tile writes its
//     (count, mean, M2) per channel; gbn_apply_fwd (bn.cu) merges them in a fixed order and does the elementwise
//     normalisation in one streaming pass -- the statistics sweep of the BatchNorm kernel is gone.
// Deterministic: no atomics, every merge has a fixed order.
//
// Main loop, two flavours (both: A = packed weights w[t][co][ci] 128 x 64 K-major SWIZZLE_128B, B = channels-last
// activations, rows = positions, accumulator = 128 channel lanes x N position columns):
//   * mode3 (k = 3, stride 1, pad 1): the activation tile is staged once per 64-channel chunk WITH its halo rows
//     (TMA box (64 ch, L + 2 positions from -1, nb breaths); out-of-range rows are zero-filled = the padding) and the
//     three taps read it through descriptors advanced by 0 / 1 / 2 rows.  Columns q >= L of a breath straddle two
//     breaths: computed and dropped.  Per chunk the SM receives 3 weight tiles + 1 activation tile instead of 3 + 3.
//   * per-tap (stride 2, 1x1): one activation load per tap, the tap shift / parity plane in the TMA coordinates.
// Rings: activation tiles (b_stages) and weight tiles (a_stages, as many as fit: a weight tile lasts only 4 MMAs, so
// the weight ring is what hides the TMA latency).
#include "tc_common.cuh"

namespace dards {

constexpr int CB_EPI_WARPS = 8;
constexpr int CB_EPI_THREADS = CB_EPI_WARPS * 32;
constexpr int CB_THREADS = 64 + CB_EPI_THREADS;
constexpr int CB_MAX_A = 8, CB_MAX_B = 4;
constexpr int CB_A_BYTES = 128 * 64 * 2;   // 16 KB weight tile
constexpr int CB_MAX_TAPS = 8;
constexpr int CB_MAX_COLS = 256;
constexpr int CB_TAIL_BYTES = 256 + 4096 + 1024;  // barriers + TMEM slot | moments [2][4][128] | scale/shift [2][128]
constexpr int CB_SMEM_LIMIT = 227 * 1024;

struct CbParams {
  int mode3;
  int n_taps;
  int w_tap[CB_MAX_TAPS], in_par[CB_MAX_TAPS], in_start[CB_MAX_TAPS];
  int k_chunks;
  int nb, l, lp;        // breaths per sub-tile, valid positions per breath, staged rows per breath (l + 2 | l)
  int n_cols;           // MMA N (multiple of 16)
  int nsub;             // sub-tiles per job (1 unless fused)
  int n_pos_jobs, n_co_tiles;
  int a_stages, b_stages, b_bytes, park_bytes;
  int c_out;
  int fuse, relu;
  int x_evict_first;
  float eps;
  const float* gamma;
  const float* beta;
  float* save_mean;
  float* save_rstd;
  float* part;          // PARTIAL: [pos job][2][3][c_out]
  __nv_bfloat16* out;
  const __nv_bfloat16* res;
  int out_stride, res_stride;
};

__device__ __forceinline__ void unpack8(const uint4& r, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return r;
}

__global__ void __launch_bounds__(CB_THREADS, 1)
    tc_conv_bn_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x,
                      const __grid_constant__ CUtensorMap tm_y, const CbParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = smem_base;
  const uint32_t a_base = b_base + p.b_stages * p.b_bytes;
  const uint32_t park_s = a_base + p.a_stages * CB_A_BYTES;
  const uint32_t bar_base = park_s + p.park_bytes;
  auto fullb = [&](int s) { return bar_base + 8u * s; };
  auto emptyb = [&](int s) { return bar_base + 8u * (CB_MAX_B + s); };
  auto fulla = [&](int s) { return bar_base + 8u * (2 * CB_MAX_B + s); };
  auto emptya = [&](int s) { return bar_base + 8u * (2 * CB_MAX_B + CB_MAX_A + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * CB_MAX_B + 2 * CB_MAX_A + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * CB_MAX_B + 2 * CB_MAX_A + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * CB_MAX_B + 2 * CB_MAX_A + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  __nv_bfloat16* park = reinterpret_cast<__nv_bfloat16*>(smem_gen + (park_s - smem_base));
  float* mom = reinterpret_cast<float*>(smem_gen + (bar_base - smem_base) + 256);   // [2][4][128]: cnt, s, ss, shift
  float* scsh = mom + 2 * 4 * 128;                                                  // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_jobs = p.n_pos_jobs * p.n_co_tiles;
  const int job_breaths = p.nb * p.nsub;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_y);
    for (int s = 0; s < p.b_stages; ++s) {
      mbar_init(fullb(s), 1);
      mbar_init(emptyb(s), 1);
    }
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(fulla(s), 1);
      mbar_init(emptya(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), CB_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      const uint32_t b_tx = (uint32_t)(p.nb * p.lp) * 128u;
      const uint64_t pol_x = l2_policy(p.x_evict_first != 0);
      for (int job = blockIdx.x; job < total_jobs; job += gridDim.x) {
        const int co0 = (job % p.n_co_tiles) * 128, nj = (job / p.n_co_tiles) * job_breaths;
        for (int sub = 0; sub < p.nsub; ++sub) {
          const int n0 = nj + sub * p.nb;
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            if (p.mode3) {
              mbar_wait(emptyb(sb), phb ^ 1u);
              mbar_arrive_expect_tx(fullb(sb), b_tx);
              tma_load_4d_pol(b_base + sb * p.b_bytes, &tm_x, fullb(sb), kc * 64, 0, -1, n0, pol_x);
              if (++sb == p.b_stages) {
                sb = 0;
                phb ^= 1u;
              }
            }
            for (int t = 0; t < p.n_taps; ++t) {
              if (!p.mode3) {
                mbar_wait(emptyb(sb), phb ^ 1u);
                mbar_arrive_expect_tx(fullb(sb), b_tx);
                tma_load_4d_pol(b_base + sb * p.b_bytes, &tm_x, fullb(sb), kc * 64, p.in_par[t], p.in_start[t], n0, pol_x);
                if (++sb == p.b_stages) {
                  sb = 0;
                  phb ^= 1u;
                }
              }
              mbar_wait(emptya(sa), pha ^ 1u);
              mbar_arrive_expect_tx(fulla(sa), CB_A_BYTES);
              tma_load_3d(a_base + sa * CB_A_BYTES, &tm_w, fulla(sa), kc * 64, co0, p.w_tap[t]);
              if (++sa == p.a_stages) {
                sa = 0;
                pha ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t tmpl = make_sw128_desc(0, 1, 1024 >> 4, 1, 0);
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      int it = 0;
      for (int job = blockIdx.x; job < total_jobs; job += gridDim.x) {
        for (int sub = 0; sub < p.nsub; ++sub, ++it) {
          const int buf = it & 1;
          mbar_wait(tempty_bar(buf), ((uint32_t)(it >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)buf * CB_MAX_COLS;
          uint32_t acc = 0u;
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            uint64_t b_desc = 0;
            if (p.mode3) {
              mbar_wait(fullb(sb), phb);
              b_desc = tmpl + (uint64_t)((b_base + sb * p.b_bytes) >> 4);
            }
            for (int t = 0; t < p.n_taps; ++t) {
              if (!p.mode3) {
                mbar_wait(fullb(sb), phb);
                b_desc = tmpl + (uint64_t)((b_base + sb * p.b_bytes) >> 4);
              }
              mbar_wait(fulla(sa), pha);
              tc_fence_after();
              const uint64_t a_desc = tmpl + (uint64_t)((a_base + sa * CB_A_BYTES) >> 4);
              const uint64_t b_t = b_desc + (uint64_t)(p.mode3 ? 8 * t : 0);  // tap t: start advanced by t rows of 128 B
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_t + (uint64_t)(2 * k), idesc, acc);
                acc = 1u;
              }
              umma_commit(emptya(sa));
              if (++sa == p.a_stages) {
                sa = 0;
                pha ^= 1u;
              }
              if (!p.mode3) {
                umma_commit(emptyb(sb));
                if (++sb == p.b_stages) {
                  sb = 0;
                  phb ^= 1u;
                }
              }
            }
            if (p.mode3) {
              umma_commit(emptyb(sb));
              if (++sb == p.b_stages) {
                sb = 0;
                phb ^= 1u;
              }
            }
          }
          umma_commit(tfull_bar(buf));
        }
      }
    }
  } else {
    // =========================== epilogue (warps 2..9) ===========================
    const int ew = warp - 2;
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32)
    const int sub2 = ew >> 2;      // the two warps of a quarter alternate over the 16-column chunks
    const int cl = quarter * 32 + lane;
    const int et = threadIdx.x - 64;
    const bool leader = (et == 0);
    const int n_chunks = p.n_cols >> 4;
    const int sub_rows = p.nb * p.l;
    int it = 0;
    for (int job = blockIdx.x; job < total_jobs; job += gridDim.x) {
      const int pos_job = job / p.n_co_tiles;
      const int co0 = (job % p.n_co_tiles) * 128, nj = pos_job * job_breaths;
      const int co = co0 + cl;
      float s = 0.f, ss = 0.f, shift = 0.f;
      int cnt = 0;
      bool have = false;
      for (int sub = 0; sub < p.nsub; ++sub, ++it) {
        const int buf = it & 1;
        mbar_wait(tfull_bar(buf), (uint32_t)(it >> 1) & 1u);
        tc_fence_after();
        if (sub == 0) {
          // the park is free once the previous job's TMA store has read it and every thread has left its last sweep
          if (leader) tma_store_wait_read();
          named_bar_sync(1, CB_EPI_THREADS);
        }
        const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)buf * CB_MAX_COLS;
        const int row_lo = sub * sub_rows, row_hi = row_lo + sub_rows;
        for (int ch = sub2; ch < n_chunks; ch += 2) {
          uint32_t v[16];
          tmem_ld16(t_row + (uint32_t)(ch << 4), v);
          tmem_ld_wait();
          const int col0 = ch << 4;
          int b = col0 / p.lp;
          int q = col0 - b * p.lp;
          int rowb = row_lo + b * p.l;
          if (!have) {
            shift = __uint_as_float(v[0]);  // any sample of the channel's data will do (columns < nb*lp - 2 are real sums)
            have = true;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x = __uint_as_float(v[j]);
            if (q < p.l && rowb < row_hi) {
              const float d = x - shift;
              s += d;
              ss = fmaf(d, d, ss);
              ++cnt;
              park[(rowb + q) * 128 + cl] = __float2bfloat16_rn(x);
            }
            if (++q == p.lp) {
              q = 0;
              rowb += p.l;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(buf));  // TMEM buffer may be overwritten by the next-but-one sub-tile
      }
      if (!p.fuse) {
        // ---- PARTIAL: this thread's moments of its columns -> part[pos job][sub2][cnt, mean, M2][channel] ----
        if (co < p.c_out) {
          const float n = (float)cnt;
          const float inv = cnt > 0 ? 1.f / n : 0.f;
          float* dst = p.part + ((size_t)(pos_job * 2 + sub2) * 3) * p.c_out + co;
          dst[0] = n;
          dst[p.c_out] = shift + s * inv;
          dst[2 * p.c_out] = fmaxf(ss - s * s * inv, 0.f);
        }
        fence_proxy_async();
        named_bar_sync(1, CB_EPI_THREADS);
        if (leader) {
          tma_store_4d(&tm_y, park_s, co0, 0, 0, nj);
          tma_store_commit();
        }
        continue;
      }
      // ---- FUSED: merge the two threads of every channel, scale / shift, raw y out, normalise out of the park ----
      mom[(sub2 * 4 + 0) * 128 + cl] = (float)cnt;
      mom[(sub2 * 4 + 1) * 128 + cl] = s;
      mom[(sub2 * 4 + 2) * 128 + cl] = ss;
      mom[(sub2 * 4 + 3) * 128 + cl] = shift;
      fence_proxy_async();  // park writes (generic proxy) -> visible to the TMA store below
      named_bar_sync(1, CB_EPI_THREADS);
      if (sub2 == 0) {
        const float n0 = mom[0 * 128 + cl], n1 = mom[4 * 128 + cl];
        const float i0 = n0 > 0.f ? 1.f / n0 : 0.f, i1 = n1 > 0.f ? 1.f / n1 : 0.f;
        const float s0 = mom[1 * 128 + cl], s1 = mom[5 * 128 + cl];
        const float m0 = mom[3 * 128 + cl] + s0 * i0, m1 = mom[7 * 128 + cl] + s1 * i1;
        const float q0 = fmaxf(mom[2 * 128 + cl] - s0 * s0 * i0, 0.f), q1 = fmaxf(mom[6 * 128 + cl] - s1 * s1 * i1, 0.f);
        const float n = n0 + n1, dlt = m1 - m0;
        const float mean = n1 > 0.f ? m0 + dlt * (n1 / n) : m0;
        const float m2 = n1 > 0.f ? q0 + q1 + dlt * dlt * (n0 * n1 / n) : q0;
        const float var = fmaxf(m2 / n, 0.f) + p.eps;
        float rstd = rsqrtf(var);
        rstd = rstd * (1.5f - 0.5f * var * rstd * rstd);  // one Newton step: the reference divides by sqrt()
        float sc = 0.f, sh = 0.f;
        if (co < p.c_out) {
          sc = rstd * p.gamma[co];
          sh = p.beta[co] - mean * sc;
          p.save_mean[(size_t)pos_job * p.c_out + co] = mean;
          p.save_rstd[(size_t)pos_job * p.c_out + co] = rstd;
        }
        scsh[cl] = sc;
        scsh[128 + cl] = sh;
      }
      named_bar_sync(1, CB_EPI_THREADS);
      if (leader) {
        tma_store_4d(&tm_y, park_s, co0, 0, 0, nj);
        tma_store_commit();
      }
      {
        const int vec = et & 15, rl = et >> 4;  // 16 vectors of 8 channels x 16 row lanes
        const int c = co0 + vec * 8;
        if (c < p.c_out) {
          float sc[8], sh[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            sc[j] = scsh[vec * 8 + j];
            sh[j] = scsh[128 + vec * 8 + j];
          }
          const int rows = p.nsub * sub_rows;
          const size_t grow0 = (size_t)nj * p.l;
          const uint4* pk = reinterpret_cast<const uint4*>(park) + vec;
          __nv_bfloat16* op = p.out + grow0 * p.out_stride + c;
          const __nv_bfloat16* rp = p.res ? p.res + grow0 * p.res_stride + c : nullptr;
          auto one = [&](int row, const uint4& rr) {
            float v[8];
            unpack8(pk[row * 16], v);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
            if (rp) {
              float e[8];
              unpack8(rr, e);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] += e[j];
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            *reinterpret_cast<uint4*>(op + (size_t)row * p.out_stride) = pack8(v);
          };
          int row = rl;
          const uint4 none = make_uint4(0u, 0u, 0u, 0u);
          for (; row + 48 < rows; row += 64) {  // 4 rows per thread in flight (the residual is the only global read)
            uint4 rr[4] = {none, none, none, none};
            if (rp) {
#pragma unroll
              for (int u = 0; u < 4; ++u) rr[u] = *reinterpret_cast<const uint4*>(rp + (size_t)(row + 16 * u) * p.res_stride);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) one(row + 16 * u, rr[u]);
          }
          for (; row < rows; row += 16) {
            uint4 rr = none;
            if (rp) rr = *reinterpret_cast<const uint4*>(rp + (size_t)row * p.res_stride);
            one(row, rr);
          }
        }
      }
    }
    if (leader) tma_store_wait_all();  // global writes complete before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct CbPlan {
  CbParams p;
  int mode;      // 0 unsupported, 1 partial statistics, 2 fused
  int smem;
  int l_in, stride;
};

static CbPlan cb_plan(int n_breaths, int group, int l_in, int l_out, int c_in, int c_out, int ktaps, int stride, int pad) {
  CbPlan w{};
  CbParams& p = w.p;
  w.mode = 0;
  w.l_in = l_in;
  w.stride = stride;
  if (group <= 0 || n_breaths <= 0 || n_breaths % group != 0) return w;
  if (c_in % 8 || c_out % 8 || ktaps > CB_MAX_TAPS) return w;
  if (stride != 1 && stride != 2) return w;
  if (l_in % stride != 0 || l_out != (l_in + 2 * pad - ktaps) / stride + 1) return w;
  p.mode3 = (ktaps == 3 && stride == 1 && pad == 1 && l_out >= 2) ? 1 : 0;
  p.n_taps = ktaps;
  for (int t = 0; t < ktaps; ++t) {
    // source position = q*stride + (t - pad) = stride*(q + floor((t-pad)/stride)) + ((t-pad) mod stride)
    const int d = t - pad;
    const int fl = d >= 0 ? d / stride : -((-d + stride - 1) / stride);
    p.w_tap[t] = t;
    p.in_par[t] = d - fl * stride;
    p.in_start[t] = fl;
  }
  p.l = l_out;
  p.lp = p.mode3 ? l_out + 2 : l_out;
  // breaths per sub-tile: the largest divisor of the group whose columns fit one accumulator
  int nb = 0;
  for (int d = 1; d <= group; ++d) {
    if (group % d) continue;
    const int cols = d * p.lp - (p.mode3 ? 2 : 0);
    if (cols <= CB_MAX_COLS) nb = d;
  }
  if (nb == 0) return w;
  p.nb = nb;
  p.n_cols = ((nb * p.lp - (p.mode3 ? 2 : 0)) + 15) / 16 * 16;
  if (p.n_cols < 64) return w;  // tiny tiles: the plain kernels + the BatchNorm kernel are the better path
  p.k_chunks = ceil_div(c_in, 64);
  p.n_co_tiles = ceil_div(c_out, 128);
  p.c_out = c_out;
  // staged rows read by the MMAs: n_cols (+ 2 for the shifted taps)
  p.b_bytes = ((p.n_cols + (p.mode3 ? 2 : 0)) * 128 + 1023) / 1024 * 1024;
  const int group_park = group * l_out * 256;
  const int tile_park = nb * l_out * 256;
  const int min_rings = 3 * CB_A_BYTES + 2 * p.b_bytes;
  p.fuse = (group_park + min_rings + 48 * 1024 + CB_TAIL_BYTES + 1024 <= CB_SMEM_LIMIT) ? 1 : 0;
  p.nsub = p.fuse ? group / nb : 1;
  p.park_bytes = ((p.fuse ? group_park : tile_park) + 1023) / 1024 * 1024;
  p.n_pos_jobs = n_breaths / (nb * p.nsub);
  int left = CB_SMEM_LIMIT - 1024 - CB_TAIL_BYTES - p.park_bytes;
  p.b_stages = p.mode3 ? 3 : 4;
  while (p.b_stages > 2 && left - p.b_stages * p.b_bytes < 3 * CB_A_BYTES) --p.b_stages;
  left -= p.b_stages * p.b_bytes;
  p.a_stages = left / CB_A_BYTES;
  if (p.a_stages > CB_MAX_A) p.a_stages = CB_MAX_A;
  if (!p.mode3 && p.a_stages > p.b_stages) {
    // per-tap loads consume one weight and one activation tile together: balance the two rings
    left += p.b_stages * p.b_bytes;
    int st = left / (CB_A_BYTES + p.b_bytes);
    if (st > CB_MAX_B) st = CB_MAX_B;
    p.b_stages = st;
    p.a_stages = st;
  }
  if (p.a_stages < 2 || p.b_stages < 2) return w;
  w.smem = p.b_stages * p.b_bytes + p.a_stages * CB_A_BYTES + p.park_bytes + CB_TAIL_BYTES + 1024;
  w.mode = p.fuse ? 2 : 1;
  return w;
}

int tc_conv_bn_mode(int n_breaths, int group, int l_in, int l_out, int c_in, int c_out, int ktaps, int stride, int pad) {
  return cb_plan(n_breaths, group, l_in, l_out, c_in, c_out, ktaps, stride, pad).mode;
}

// PARTIAL mode: number of (count, mean, M2) records per group and channel, and the size of `part` in floats
int tc_conv_bn_part_entries(int n_breaths, int group, int l_in, int l_out, int c_in, int c_out, int ktaps, int stride, int pad) {
  CbPlan w = cb_plan(n_breaths, group, l_in, l_out, c_in, c_out, ktaps, stride, pad);
  if (w.mode != 1) return 0;
  return 2 * (group / w.p.nb);
}

int tc_conv_bn_fwd(const void* in, const void* w_koi, void* y, void* out, const void* res, const float* gamma,
                   const float* beta, float* save_mean, float* save_rstd, float* part, int n_breaths, int group, int l_in,
                   int l_out, int c_in, int c_out, int in_stride, int y_stride, int out_stride, int res_stride, int ktaps,
                   int stride, int pad, float eps, int relu, int src_last_use, cudaStream_t st) {
  CbPlan w = cb_plan(n_breaths, group, l_in, l_out, c_in, c_out, ktaps, stride, pad);
  if (w.mode == 0) {
    set_error("conv+bn: unsupported shape (n=%d group=%d l=%d->%d cin=%d cout=%d k=%d s=%d p=%d)", n_breaths, group, l_in,
              l_out, c_in, c_out, ktaps, stride, pad);
    return DARDS_ERR_UNSUPPORTED;
  }
  DARDS_CHECK_ARG(in_stride % 8 == 0 && y_stride % 8 == 0, "conv+bn: row strides must be multiples of 8");
  DARDS_CHECK_ARG((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_koi) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(y) & 15) == 0,
                  "conv+bn: operands must be 16-byte aligned");
  CbParams& p = w.p;
  if (w.mode == 2) {
    DARDS_CHECK_ARG(out && gamma && beta && save_mean && save_rstd, "conv+bn (fused): null pointer");
    DARDS_CHECK_ARG(out_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "conv+bn: `out` must be 16-byte aligned rows");
    DARDS_CHECK_ARG(!res || (res_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(res) & 15) == 0),
                    "conv+bn: `res` must be 16-byte aligned rows");
  } else {
    DARDS_CHECK_ARG(part, "conv+bn (partial statistics): null `part`");
  }
  p.eps = eps;
  p.relu = relu ? 1 : 0;
  p.gamma = gamma;
  p.beta = beta;
  p.save_mean = save_mean;
  p.save_rstd = save_rstd;
  p.part = part;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.res = static_cast<const __nv_bfloat16*>(res);
  p.out_stride = out_stride;
  p.res_stride = res_stride;
  p.x_evict_first = (src_last_use && p.n_co_tiles == 1 && p.mode3 && g_dbg_l2_hint != 0) ? 1 : 0;
  CUtensorMap tm_w, tm_x, tm_y;
  {
    cuuint64_t dims[3] = {(cuuint64_t)c_in, (cuuint64_t)c_out, (cuuint64_t)ktaps};
    cuuint64_t str[2] = {(cuuint64_t)c_in * 2, (cuuint64_t)c_in * c_out * 2};
    cuuint32_t box[3] = {64, 128, 1};
    int rc = make_bf16_map(&tm_w, w_koi, 3, dims, str, box, true);
    if (rc) return rc;
  }
  {
    const int planes = stride, l_plane = l_in / planes;
    cuuint64_t dims[4] = {(cuuint64_t)c_in, (cuuint64_t)planes, (cuuint64_t)l_plane, (cuuint64_t)n_breaths};
    cuuint64_t str[3] = {(cuuint64_t)in_stride * 2, (cuuint64_t)in_stride * planes * 2, (cuuint64_t)in_stride * l_in * 2};
    cuuint32_t box[4] = {64, 1, (cuuint32_t)p.lp, (cuuint32_t)p.nb};
    int rc = make_bf16_map(&tm_x, in, 4, dims, str, box, true);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)c_out, 1, (cuuint64_t)l_out, (cuuint64_t)n_breaths};
    cuuint64_t str[3] = {(cuuint64_t)y_stride * 2, (cuuint64_t)y_stride * 2, (cuuint64_t)y_stride * l_out * 2};
    cuuint32_t box[4] = {128, 1, (cuuint32_t)l_out, (cuuint32_t)(p.nb * p.nsub)};
    int rc = make_bf16_map(&tm_y, y, 4, dims, str, box, false);
    if (rc) return rc;
  }
  static int attr_smem = 0;
  if (w.smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_bn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, w.smem);
    if (e != cudaSuccess) {
      set_error("conv+bn: cannot opt in to %d bytes of shared memory: %s", w.smem, cudaGetErrorString(e));
      return DARDS_ERR_CUDA;
    }
    attr_smem = w.smem;
  }
  const int jobs = p.n_pos_jobs * p.n_co_tiles;
  const int grid = jobs < sm_count() ? jobs : sm_count();
  tc_conv_bn_kernel<<<grid, CB_THREADS, w.smem, st>>>(tm_w, tm_x, tm_y, p);
  DARDS_CHECK_LAUNCH("tc_conv_bn");
  return DARDS_OK;
}

}  // namespace dards
