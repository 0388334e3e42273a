// Micro-benchmark of tcgen05.mma issue cost on sm_100a: cycles per MMA (M = 128, K = 16, bf16) as a function of N, for
// SS mode (A and B from shared memory), B descriptors advanced by whole rows (the 3-tap trick of conv_tc.cu), and A from
// TMEM.  One thread per CTA issues `reps` batches of `per_batch` MMAs, commits to an mbarrier and waits; clock64 around it.
// No data dependence on the operand contents (shared memory is zeroed).  Build: nvcc -gencode arch=compute_100a,code=sm_100a
//   -O3 -o tools/mma_probe tools/mma_probe.cu ; run on the GPU box: tools/mma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
               "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// mode 0: SS, same descriptors every MMA; 1: SS, B start advanced by (i % 3) rows; 2: A from TMEM; 3: SS with the K-advance
// pattern of a real main loop (4 k-steps over a 64-wide tile, ring of `stages` distinct tiles)
__global__ void __launch_bounds__(128, 1) probe(int n, int mode, int reps, int per_batch, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int stages = 4;
  const uint32_t a_bytes = 16384, b_bytes = 34 * 1024;
  const uint32_t bar = base + stages * (a_bytes + b_bytes);
  const uint32_t slot = bar + 32;
  for (uint32_t i = threadIdx.x; i < stages * (a_bytes + b_bytes) / 16; i += blockDim.x) reinterpret_cast<uint4*>(gen)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (slot - base));
  if (threadIdx.x == 0) {
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // modes 12..14: MN-major operands (the weight-gradient kernel's layout: rows = K index, 128-byte row = 64 elements of M/N)
    if (mode == 12 || mode == 13) idesc |= 1u << 15;
    if (mode == 12 || mode == 14) idesc |= 1u << 16;
    uint32_t phase = 0;
    long long best = 1ll << 60;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      if (mode >= 4) {
        // lean loop: the 16 (stage, k-step) descriptor pairs are loop invariant, 16 MMAs per iteration, nothing else
        uint64_t ad[16], bd[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t sa = base + (j >> 2) * (a_bytes + b_bytes);
          ad[j] = make_desc(sa) + 2 * (j & 3);
          bd[j] = make_desc(sa + a_bytes) + 2 * (j & 3);
          if (mode == 5) bd[j] += 8 * ((j >> 2) % 3);      // taps: B start advanced by 0 / 1 / 2 rows
          if (mode == 8) bd[j] += 8 * 1;                   // every MMA one row off the swizzle atom
          if (mode >= 12) {
            // MN-major: a K step of 16 = 16 rows = 2048 bytes; LBO = distance between 64-column chunks (8 KB here)
            const uint64_t lbo = (uint64_t)(8192 >> 4) << 16;
            ad[j] = (make_desc(sa) & ~((uint64_t)0x3FFF << 16)) | lbo;
            bd[j] = (make_desc(sa + a_bytes) & ~((uint64_t)0x3FFF << 16)) | lbo;
            ad[j] += 128 * (j & 3);
            bd[j] += 128 * (j & 3);
          }
        }
        const uint32_t t1 = tmem + (uint32_t)((n + 31) / 32 * 32 > 256 ? 0 : n);  // second accumulator right behind the first
        for (int i = 0; i < per_batch; i += 16) {
          if (mode == 6) {
#pragma unroll
            for (int j = 0; j < 16; ++j) umma_ss((j & 4) ? t1 : tmem, ad[j], bd[j], idesc, 1u);   // alternate accumulators every 4 MMAs
          } else if (mode >= 9) {
            // a commit (to a barrier nobody waits on) after every 4 / 8 / 16 MMAs, like a real main loop's "stage free" signal
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              umma_ss(tmem, ad[j], bd[j], idesc, 1u);
              if (mode == 9 && (j & 3) == 3) umma_commit(bar + 8);
              if (mode == 10 && (j & 7) == 7) umma_commit(bar + 8);
              if (mode == 11 && j == 15) umma_commit(bar + 8);
            }
          } else if (mode == 7) {
#pragma unroll
            for (int j = 0; j < 16; ++j) umma_ts(tmem, tmem + 448 + 8 * (j & 3), bd[j], idesc, 1u);  // A from TMEM
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) umma_ss(tmem, ad[j], bd[j], idesc, 1u);
          }
        }
      } else
      for (int i = 0; i < per_batch; ++i) {
        const int st = (i >> 2) % stages;
        const uint32_t sa = base + st * (a_bytes + b_bytes), sb = sa + a_bytes;
        uint64_t ad = make_desc(sa), bd = make_desc(sb);
        if (mode == 1) bd += (uint64_t)(8 * (i % 3));
        if (mode == 1 || mode == 3) { ad += 2 * (i & 3); bd += 2 * (i & 3); }
        if (mode == 2) umma_ts(tmem, tmem + 256 + 8 * (i & 3), bd, idesc, 1u);
        else umma_ss(tmem, ad, bd, idesc, 1u);
      }
      umma_commit(bar);
      while (!mbar_try_wait(bar, phase)) {}
      phase ^= 1u;
      const long long dt = clock64() - t0;
      if (dt < best) best = dt;
    }
    out[blockIdx.x] = best;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

int main() {
  const int smem = 4 * (16384 + 34 * 1024) + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  const int per_batch = 256;
  const char* names[15] = {"SS same tile", "SS row-shifted B (taps)", "A from TMEM", "SS k-advance, 4-tile ring", "lean: SS k-advance", "lean: SS taps (B +0/1/2 rows)", "lean: SS two accumulators", "lean: A from TMEM", "lean: SS B +1 row always", "lean: commit every 4 MMAs", "lean: commit every 8 MMAs", "lean: commit every 16 MMAs", "lean: A and B MN-major", "lean: A MN-major", "lean: B MN-major"};
  for (int grid : {148}) {
    for (int mode : {4, 12, 13, 14}) {
      for (int n : {64, 128, 144, 160, 192, 224, 256}) {
        probe<<<grid, 128, smem>>>(n, mode, 5, per_batch, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("grid %d mode %d n %d: %s\n", grid, mode, n, cudaGetErrorString(e));
          return 1;
        }
        long long h[148];
        cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("grid %3d  %-28s N=%3d: %7.1f clk/MMA (nominal %5.1f)  -> %4.1f %% of the tensor rate\n", grid, names[mode], n,
               (double)mx / per_batch, n / 2.0, 100.0 * (n / 2.0) / ((double)mx / per_batch));
      }
    }
  }
  return 0;
}
