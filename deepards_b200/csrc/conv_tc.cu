// tcgen05 / TMEM / TMA implicit-GEMM Conv1d for sm_100a (bf16 operands, fp32 accumulation in TMEM).
//
// GEMM view (forward; dgrad is the same kernel with the roles of the channel dims swapped):
//     D[co, pos] = sum_{tap t} sum_{ci} W_t[co, ci] * X[pos shifted by t, ci]
//   A operand  = packed weights  w[t][co][ci]  -> 128(co) x 64(ci) K-major tile, SWIZZLE_128B
//   B operand  = channels-last activations     -> N_TILE(pos) x 64(ci) K-major tile, SWIZZLE_128B
//   accumulator D in TMEM: 128 lanes (output channels) x N_TILE fp32 columns (positions)
// The convolution is im2col-free: for each tap the SAME activation tensor is fetched by TMA with the
// position coordinate shifted by (t - pad); rows that fall outside [0, L) are zero-filled by the TMA
// unit, which is exactly the conv padding.  Stride-2 convolutions address the input through a 4-D view
// (C, 2, L/2, N) and pick the parity plane, so no strided gather is needed either.
// A position tile is NB whole breaths (NB*L = 224 columns for L = 56/28/14/7), so tiles never straddle a
// breath and there is no wasted MMA column.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc/dealloc),
// warps 2..5 = epilogue (TMEM -> registers -> global; one output channel per thread).
// Pipelines: smem ring (full/empty mbarriers, TMA <-> MMA) and a 2-deep TMEM accumulator ring
// (tmem_full/tmem_empty, MMA <-> epilogue) so the epilogue of tile i overlaps the MMAs of tile i+1.
// The kernel is persistent: grid = min(#tiles, #SMs), static round-robin tile schedule.
#include <cuda.h>

#include "common.cuh"

namespace dards {

constexpr int TC_THREADS = 192;
constexpr int TC_STAGES = 4;
constexpr int TC_BLOCK_M = 128;          // output channels per tile (TMEM lanes)
constexpr int TC_BLOCK_K = 64;           // reduction channels per stage = one 128-byte swizzle row
constexpr int TC_MAX_N = 256;            // positions per tile (TMEM columns per accumulator)
constexpr int TC_A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;   // 16 KB
constexpr int TC_B_BYTES = TC_MAX_N * TC_BLOCK_K * 2;     // 32 KB
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int TC_MAX_TAPS = 8;

struct TcConvParams {
  int n_taps;                 // taps issued by this launch
  int w_tap[TC_MAX_TAPS];     // tap coordinate in the weight map
  int in_par[TC_MAX_TAPS];    // parity-plane coordinate in the activation map
  int in_start[TC_MAX_TAPS];  // first position coordinate in the activation map (may be negative)
  int k_chunks;               // ceil(reduction channels / 64)
  int c_cols;                 // valid output channels
  int nb, l_tile;             // breaths per tile, positions per breath in the tile; N_TILE = nb*l_tile
  int n_breaths;
  int out_l, out_mul, out_off;  // output row = n*out_l + m*out_mul + out_off
  int out_stride, addend_stride;
  int n_pos_tiles, n_co_tiles;
};

static int g_dbg_lbo = -1, g_dbg_version = -1, g_dbg_sbo = -1;

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (an error the host sees), never as a hang.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s at 2 GHz
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread i of the warp = lane base+i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major, canonical value 1)
//   [32,46) SBO >> 4 = 1024 B between 8-row core-matrix groups | [46,48) version = 1 | [61,64) layout = 2
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16,
                                                           uint32_t version) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(lbo16 & 0x3FFF) << 16;
  d |= (uint64_t)(sbo16 & 0x3FFF) << 32;
  d |= (uint64_t)(version & 0x3) << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// ---------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
    tc_conv_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x,
                   __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ addend, const TcConvParams p,
                   uint32_t desc_lbo16, uint32_t desc_sbo16, uint32_t desc_version) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + TC_STAGES * TC_STAGE_BYTES;
  // barrier layout (8 bytes each): full[S], empty[S], tmem_full[2], tmem_empty[2], then the TMEM base word
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TC_STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * TC_STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * TC_STAGES + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * TC_STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = p.nb * p.l_tile;
  const int total_tiles = p.n_pos_tiles * p.n_co_tiles;
  const int k_steps = p.n_taps * p.k_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_x);
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);  // 2 accumulators x 256 columns
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t stage_tx = TC_A_BYTES + (uint32_t)n_tile * (TC_BLOCK_K * 2);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int co_tile = tile % p.n_co_tiles, pos_tile = tile / p.n_co_tiles;
        const int co0 = co_tile * TC_BLOCK_M, n0 = pos_tile * p.nb;
        for (int ti = 0; ti < p.n_taps; ++ti) {
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = smem_base + stage * TC_STAGE_BYTES, sb = sa + TC_A_BYTES;
            mbar_arrive_expect_tx(full_bar(stage), stage_tx);
            tma_load_3d(sa, &tm_w, full_bar(stage), kc * TC_BLOCK_K, co0, p.w_tap[ti]);
            tma_load_4d(sb, &tm_x, full_bar(stage), kc * TC_BLOCK_K, p.in_par[ti], p.in_start[ti], n0);
            if (++stage == TC_STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
      // A,B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_tile >> 3) << 17) |
                             ((uint32_t)(TC_BLOCK_M >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty_bar(buf), acc_phase ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * TC_MAX_N;
        for (int ks = 0; ks < k_steps; ++ks) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * TC_STAGE_BYTES, sb = sa + TC_A_BYTES;
          const uint64_t a_desc = make_kmajor_sw128_desc(sa, desc_lbo16, desc_sbo16, desc_version);
          const uint64_t b_desc = make_kmajor_sw128_desc(sb, desc_lbo16, desc_sbo16, desc_version);
#pragma unroll
          for (int k = 0; k < TC_BLOCK_K / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr>>4) field
            umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (ks | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs are done
          if (++stage == TC_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(tfull_bar(buf));  // accumulator complete -> epilogue
      }
    }
  } else {
    // =========================== epilogue (warps 2..5) ===========================
    const int quarter = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const int co_tile = tile % p.n_co_tiles, pos_tile = tile / p.n_co_tiles;
      const int co = co_tile * TC_BLOCK_M + quarter * 32 + lane;
      const int n0 = pos_tile * p.nb;
      mbar_wait(tfull_bar(buf), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)buf * TC_MAX_N;
      const bool co_ok = co < p.c_cols;
      for (int c0 = 0; c0 < n_tile; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(t_row + (uint32_t)c0, v);
        tmem_ld_wait();
        if (co_ok) {
          int b = c0 / p.l_tile, m = c0 % p.l_tile;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + b;
            if (n < p.n_breaths) {
              const size_t row = (size_t)n * p.out_l + (size_t)m * p.out_mul + p.out_off;
              float val = __uint_as_float(v[j]);
              if (addend) val += __bfloat162float(addend[row * p.addend_stride + co]);
              out[row * p.out_stride + co] = __float2bfloat16_rn(val);
            }
            if (++m == p.l_tile) {
              m = 0;
              ++b;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

// bf16 tensor map, SWIZZLE_128B, zero fill out of bounds.  dims/strides innermost first; strides in BYTES for
// dims 1..rank-1.
static int make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                    const cuuint32_t* box) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return DARDS_ERR_CUDA;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_b, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u]", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
              (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return DARDS_ERR_CUDA;
  }
  return DARDS_OK;
}

static int sm_count() {
  static int n = 0;
  if (n) return n;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (n <= 0) n = 148;
  return n;
}

// breaths per position tile: the largest nb with nb*l <= 256 and (nb*l) % 16 == 0
static int pick_nb(int l) {
  int best = 0;
  for (int nb = 1; nb * l <= TC_MAX_N && nb <= 256; ++nb)
    if ((nb * l) % 16 == 0) best = nb;
  return best;
}

struct TcProblem {
  const void* src;     // activations read by the MMA (x for fwd, dout for dgrad)
  const void* w;       // packed weights [ktaps][c_cols][c_red]
  void* dst;
  const void* addend;
  int n_breaths, l_src, src_planes /*1 or 2: parity planes of the source view*/;
  int c_red, c_cols, src_stride, dst_stride, addend_stride, ktaps_total;
  TcConvParams p;
};

static int tc_launch(const TcProblem& q, cudaStream_t st) {
  DARDS_CHECK_ARG(q.c_red % 8 == 0 && q.src_stride % 8 == 0, "tcgen05 conv: channels/strides must be multiples of 8");
  DARDS_CHECK_ARG((reinterpret_cast<uintptr_t>(q.src) & 15) == 0 && (reinterpret_cast<uintptr_t>(q.w) & 15) == 0,
                  "tcgen05 conv: operands must be 16-byte aligned");
  DARDS_CHECK_ARG(q.l_src % q.src_planes == 0, "tcgen05 conv: source length not divisible by the stride");
  if (q.n_breaths == 0) return DARDS_OK;
  TcConvParams p = q.p;
  const int n_tile = p.nb * p.l_tile;
  DARDS_CHECK_ARG(p.nb > 0 && n_tile % 16 == 0 && n_tile <= TC_MAX_N && n_tile >= 16,
                  "tcgen05 conv: unsupported position tile (%d breaths x %d)", p.nb, p.l_tile);
  CUtensorMap tm_w, tm_x;
  {
    cuuint64_t dims[3] = {(cuuint64_t)q.c_red, (cuuint64_t)q.c_cols, (cuuint64_t)q.ktaps_total};
    cuuint64_t str[2] = {(cuuint64_t)q.c_red * 2, (cuuint64_t)q.c_red * q.c_cols * 2};
    cuuint32_t box[3] = {TC_BLOCK_K, TC_BLOCK_M, 1};
    int rc = make_map(&tm_w, q.w, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    const int planes = q.src_planes, l_plane = q.l_src / planes;
    cuuint64_t dims[4] = {(cuuint64_t)q.c_red, (cuuint64_t)planes, (cuuint64_t)l_plane, (cuuint64_t)q.n_breaths};
    cuuint64_t str[3] = {(cuuint64_t)q.src_stride * 2, (cuuint64_t)q.src_stride * planes * 2,
                         (cuuint64_t)q.src_stride * q.l_src * 2};
    cuuint32_t box[4] = {TC_BLOCK_K, 1, (cuuint32_t)p.l_tile, (cuuint32_t)p.nb};
    int rc = make_map(&tm_x, q.src, 4, dims, str, box);
    if (rc) return rc;
  }
  p.k_chunks = ceil_div(q.c_red, TC_BLOCK_K);
  p.c_cols = q.c_cols;
  p.n_breaths = q.n_breaths;
  p.out_stride = q.dst_stride;
  p.addend_stride = q.addend_stride;
  p.n_pos_tiles = ceil_div(q.n_breaths, p.nb);
  p.n_co_tiles = ceil_div(q.c_cols, TC_BLOCK_M);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("tcgen05 conv: cannot opt in to %d bytes of shared memory: %s", TC_SMEM_BYTES, cudaGetErrorString(e));
      return DARDS_ERR_CUDA;
    }
    attr_set = true;
  }
  const int tiles = p.n_pos_tiles * p.n_co_tiles;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  const uint32_t lbo = g_dbg_lbo >= 0 ? (uint32_t)g_dbg_lbo : 1u;
  const uint32_t sbo = g_dbg_sbo >= 0 ? (uint32_t)g_dbg_sbo : (1024u >> 4);
  const uint32_t ver = g_dbg_version >= 0 ? (uint32_t)g_dbg_version : 1u;
  tc_conv_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(tm_w, tm_x, static_cast<__nv_bfloat16*>(q.dst),
                                                          static_cast<const __nv_bfloat16*>(q.addend), p, lbo, sbo, ver);
  DARDS_CHECK_LAUNCH("tc_conv");
  return DARDS_OK;
}

int tc_conv_fwd(const void* in, const void* w_koi, void* out, const void* addend, int n_breaths, int l_in, int l_out,
                int c_in, int c_out, int in_stride, int out_stride, int addend_stride, int ktaps, int stride, int pad,
                cudaStream_t st) {
  DARDS_CHECK_ARG(stride == 1 || stride == 2, "tcgen05 conv: stride must be 1 or 2");
  DARDS_CHECK_ARG(ktaps <= TC_MAX_TAPS, "tcgen05 conv: at most %d taps", TC_MAX_TAPS);
  TcProblem q{};
  q.src = in; q.w = w_koi; q.dst = out; q.addend = addend;
  q.n_breaths = n_breaths; q.l_src = l_in; q.src_planes = stride;
  q.c_red = c_in; q.c_cols = c_out; q.src_stride = in_stride; q.dst_stride = out_stride; q.addend_stride = addend_stride;
  q.ktaps_total = ktaps;
  TcConvParams& p = q.p;
  p.n_taps = ktaps;
  for (int t = 0; t < ktaps; ++t) {
    // source position = q*stride + (t - pad) = stride*(q + floor((t-pad)/stride)) + ((t-pad) mod stride)
    const int d = t - pad;
    const int fl = d >= 0 ? d / stride : -((-d + stride - 1) / stride);
    p.w_tap[t] = t;
    p.in_par[t] = d - fl * stride;
    p.in_start[t] = fl;
  }
  p.l_tile = l_out;
  p.nb = pick_nb(l_out);
  p.out_l = l_out; p.out_mul = 1; p.out_off = 0;
  return tc_launch(q, st);
}

int tc_conv_dgrad(const void* dout, const void* w_kio, void* din, const void* addend, int n_breaths, int l_in, int l_out,
                  int c_in, int c_out, int dout_stride, int din_stride, int addend_stride, int ktaps, int stride, int pad,
                  cudaStream_t st) {
  DARDS_CHECK_ARG(ktaps <= TC_MAX_TAPS, "tcgen05 conv: at most %d taps", TC_MAX_TAPS);
  if (stride != 1) {
    set_error("tcgen05 dgrad: stride %d is not implemented on the tensor-core path", stride);
    return DARDS_ERR_UNSUPPORTED;
  }
  // din[p] = sum_t dout[p + pad - t] * W_t  -> the forward kernel with start = pad - t, reduction over c_out
  TcProblem q{};
  q.src = dout; q.w = w_kio; q.dst = din; q.addend = addend;
  q.n_breaths = n_breaths; q.l_src = l_out; q.src_planes = 1;
  q.c_red = c_out; q.c_cols = c_in; q.src_stride = dout_stride; q.dst_stride = din_stride; q.addend_stride = addend_stride;
  q.ktaps_total = ktaps;
  TcConvParams& p = q.p;
  p.n_taps = ktaps;
  for (int t = 0; t < ktaps; ++t) {
    p.w_tap[t] = t;
    p.in_par[t] = 0;
    p.in_start[t] = pad - t;
  }
  p.l_tile = l_in;
  p.nb = pick_nb(l_in);
  p.out_l = l_in; p.out_mul = 1; p.out_off = 0;
  return tc_launch(q, st);
}

int tc_conv_wgrad(const void*, const void*, float*, int, void*, long long, int, int, int, int, int, int, int, int, int,
                  int, cudaStream_t) {
  set_error("tcgen05 wgrad is not implemented yet");
  return DARDS_ERR_UNSUPPORTED;
}

long long tc_wgrad_workspace_bytes(int, int, int, int, int) { return 0; }

int tc_debug_set(int key, int value) {
  if (key == 0) g_dbg_lbo = value;
  else if (key == 1) g_dbg_version = value;
  else if (key == 2) g_dbg_sbo = value;
  else {
    set_error("tc_debug_set: unknown key %d", key);
    return DARDS_ERR_INVALID_ARGUMENT;
  }
  return DARDS_OK;
}

}  // namespace dards
