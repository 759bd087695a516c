// umma_probe.cu -- how long does one tcgen05.mma.kind::i8 (M = 128, K = 32) take as a function of N, of the
// operand source (A from shared memory or from TMEM) and of the accumulator pattern?  One CTA per SM, one
// issuing thread, operands resident in shared memory (no loads in the timed region), random digit bytes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe tools/umma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n"
               ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc_of(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)8 << 16) | ((uint64_t)16 << 32) | (1ull << 46);
}

// pattern 0: the 21 digit pairs of the Ozaki kernel (6 accumulators of N columns, N <= 80)
// pattern 1: every MMA into accumulator 0, planes cycling
// pattern 2: 3 accumulators of N columns, 6 pairs each round robin (N <= 160)
__global__ void __launch_bounds__(64, 1) probe(int N, int pattern, int ts, int iters, unsigned long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  unsigned char* sA = smem;               // 6 planes x 4096
  unsigned char* sB = smem + 6 * 4096;    // 6 planes x N*32
  for (int i = threadIdx.x; i < 6 * 4096 + 6 * N * 32; i += blockDim.x)
    smem[i] = (unsigned char)((i * 2654435761u + blockIdx.x * 40503u) >> 13);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> async proxy (MMA)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
  if (warp == 1) {
    uint32_t e;
    asm volatile("{\n.reg .pred px;\n.reg .b32 rx;\nelect.sync rx|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}\n" : "=r"(e));
    long long t0 = 0, t1 = 0;
    if (e) {
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
      const uint32_t a_tm = tm + 480;  // 6 planes x 8 columns when A comes from TMEM (pattern 0 uses cols < 480)
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        if (pattern == 0) {
#pragma unroll
          for (int g = 0; g < 6; ++g)
#pragma unroll
            for (int s = 0; s <= g; ++s) {
              if (ts) umma_ts(tm + g * N, a_tm + (s & 3) * 8, desc_of(b0 + (g - s) * N * 32), idesc, (it | s) ? 1u : 0u);
              else umma_ss(tm + g * N, desc_of(a0 + s * 4096), desc_of(b0 + (g - s) * N * 32), idesc, (it | s) ? 1u : 0u);
            }
        } else if (pattern == 1) {
#pragma unroll
          for (int j = 0; j < 21; ++j) {
            if (ts) umma_ts(tm, a_tm + (j & 3) * 8, desc_of(b0 + (j % 6) * N * 32), idesc, (it | j) ? 1u : 0u);
            else umma_ss(tm, desc_of(a0 + (j % 6) * 4096), desc_of(b0 + ((j + 1) % 6) * N * 32), idesc, (it | j) ? 1u : 0u);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 21; ++j) {
            const int g = j % 3;
            if (ts) umma_ts(tm + g * N, a_tm + (j & 3) * 8, desc_of(b0 + (j % 6) * N * 32), idesc, (it | (j / 3)) ? 1u : 0u);
            else umma_ss(tm + g * N, desc_of(a0 + (j % 6) * 4096), desc_of(b0 + ((j + 1) % 6) * N * 32), idesc, (it | (j / 3)) ? 1u : 0u);
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
      t1 = clock64();
      out[blockIdx.x] = (unsigned long long)(t1 - t0);
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* out;
  cudaMalloc(&out, sizeof(unsigned long long) * sms);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 4096 + 6 * 256 * 32);
  unsigned long long* h = (unsigned long long*)malloc(sizeof(unsigned long long) * sms);
  const int iters = 2000;
  struct Cfg { int N, pattern, ts; } cfgs[] = {
      {16, 0, 0}, {32, 0, 0}, {48, 0, 0}, {64, 0, 0}, {80, 0, 0}, {80, 0, 1}, {64, 0, 1},
      {80, 1, 0}, {128, 1, 0}, {160, 1, 0}, {256, 1, 0}, {80, 1, 1}, {128, 1, 1}, {160, 1, 1}, {256, 1, 1},
      {160, 2, 0}, {160, 2, 1}, {128, 2, 0}, {128, 2, 1}};
  for (auto c : cfgs) {
    for (int grid : {1, sms}) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      const size_t smem = 6 * 4096 + 6 * c.N * 32;
      probe<<<grid, 64, smem>>>(c.N, c.pattern, c.ts, 50, out);  // warm-up
      cudaEventRecord(e0);
      probe<<<grid, 64, smem>>>(c.N, c.pattern, c.ts, iters, out);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("CUDA error: %s (N=%d pattern=%d ts=%d)\n", cudaGetErrorString(err), c.N, c.pattern, c.ts); return 1; }
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      cudaMemcpy(h, out, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost);
      double cyc = 0; for (int i = 0; i < grid; ++i) cyc += (double)h[i]; cyc /= grid;
      const double per = cyc / (iters * 21.0);
      const double tops = (double)grid * iters * 21.0 * 128.0 * c.N * 32.0 * 2.0 / (ms * 1e-3) / 1e12;
      printf("{\"N\": %d, \"pattern\": %d, \"a_from_tmem\": %d, \"ctas\": %d, \"clk_per_mma\": %.2f, \"ideal_clk\": %.1f, \"ms\": %.3f, \"int8_tops\": %.1f, \"sm_ghz\": %.3f}\n",
             c.N, c.pattern, c.ts, grid, per, c.N / 2.0, ms, tops, cyc / (ms * 1e6));
    }
  }
  return 0;
}
