// umma_probe2.cu -- A operand from TMEM (tcgen05.mma ... [a_tmem], b_desc) fed by tcgen05.cp.128x256b:
// (1) exactness of cp + TS-mode kind::i8 MMA against the CPU on random digits, (2) cycles per k-step of the
// Ozaki pattern (6 cp + 21 MMA, A slots double-buffered) and of the copies alone.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe2 tools/umma_probe2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n"
               ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void utccp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ uint64_t desc_of(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)8 << 16) | ((uint64_t)16 << 32) | (1ull << 46);
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// mode 0: verify (A, B given in gmem as canonical images; out = D (128 x N) via TS and via SS)
// mode 1: time 6 cp + 21 TS MMA per iteration;  mode 2: time 6 cp only;  mode 3: 21 TS MMA only
__global__ void __launch_bounds__(128, 1) probe(int N, int mode, int iters, const unsigned char* gA, const unsigned char* gB,
                                                int* outTS, int* outSS, unsigned long long* cyc) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar, bar2, bar3;
  __shared__ uint32_t slot;
  unsigned char* sA = smem;             // 6 planes x 4096
  unsigned char* sB = smem + 6 * 4096;  // 6 planes x N*32
  for (int i = threadIdx.x; i < 6 * 4096 + 6 * N * 32; i += blockDim.x) {
    unsigned char v = (unsigned char)((i * 2654435761u + blockIdx.x * 40503u) >> 13);
    if (mode == 0) v = i < 6 * 4096 ? (i < 4096 ? gA[i] : 0) : (i - 6 * 4096 < N * 32 ? gB[i - 6 * 4096] : 0);
    smem[i] = v;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(smem_u32(&bar3)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
  const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
  const uint32_t a_tm = tm + 6 * N;  // two slots of 48 columns behind the accumulators
  if (warp == 1) {
    uint32_t e;
    asm volatile("{\n.reg .pred px;\n.reg .b32 rx;\nelect.sync rx|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}\n" : "=r"(e));
    if (e) {
      if (mode == 0) {
        utccp_128x256b(a_tm, desc_of(a0));
        umma_ts(tm, a_tm, desc_of(b0), idesc, 0u);          // D0 = A(tmem) B
        umma_ss(tm + N, desc_of(a0), desc_of(b0), idesc, 0u);  // D1 = A(smem) B
        commit(&bar);
      } else if (mode >= 4) {
        // per-iteration overhead candidates around 21 TS MMAs: 4 = + commit, 5 = + fence::after_thread_sync,
        // 6 = + try_wait on an already completed barrier phase, 7 = 4+5+6
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
          const uint32_t as = a_tm + (it & 1) * 48;
          if (mode == 6 || mode == 7) wait_bar(&bar2, 1);  // fresh barrier: parity 1 is "already complete"
          if (mode == 5 || mode == 7) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int g = 0; g < 6; ++g)
#pragma unroll
            for (int s = 0; s <= g; ++s)
              umma_ts(tm + g * N, as + s * 8, desc_of(b0 + (g - s) * N * 32), idesc, (it | s) ? 1u : 0u);
          if (mode == 4 || mode == 7) commit(&bar3);
        }
        commit(&bar);
        wait_bar(&bar, 0);
        cyc[blockIdx.x] = (unsigned long long)(clock64() - t0);
      } else {
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
          const uint32_t as = a_tm + (it & 1) * 48;
          if (mode == 1 || mode == 2) {
#pragma unroll
            for (int p = 0; p < 6; ++p) utccp_128x256b(as + p * 8, desc_of(a0 + p * 4096));
          }
          if (mode == 1 || mode == 3) {
#pragma unroll
            for (int g = 0; g < 6; ++g)
#pragma unroll
              for (int s = 0; s <= g; ++s)
                umma_ts(tm + g * N, as + s * 8, desc_of(b0 + (g - s) * N * 32), idesc, (it | s) ? 1u : 0u);
          }
        }
        commit(&bar);
        wait_bar(&bar, 0);
        cyc[blockIdx.x] = (unsigned long long)(clock64() - t0);
      }
    }
    __syncwarp();
  }
  if (mode == 0) {
    wait_bar(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < N; ++c) {
      int v0, v1;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v0) : "r"(tm + ((uint32_t)(warp * 32) << 16) + c));
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v1) : "r"(tm + ((uint32_t)(warp * 32) << 16) + N + c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      outTS[(warp * 32 + lane) * N + c] = v0;
      outSS[(warp * 32 + lane) * N + c] = v1;
    }
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
  const int N = 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 4096 + 6 * 80 * 32);
  // ---- verification
  signed char A[128][32], B[N][32];
  srand(1);
  for (int i = 0; i < 128; ++i) for (int k = 0; k < 32; ++k) A[i][k] = (signed char)(rand() % 256 - 128);
  for (int j = 0; j < N; ++j) for (int k = 0; k < 32; ++k) B[j][k] = (signed char)(rand() % 256 - 128);
  unsigned char imgA[4096], imgB[N * 32];
  for (int i = 0; i < 128; ++i) for (int k = 0; k < 32; ++k) imgA[(i / 8) * 256 + (k / 16) * 128 + (i % 8) * 16 + k % 16] = (unsigned char)A[i][k];
  for (int j = 0; j < N; ++j) for (int k = 0; k < 32; ++k) imgB[(j / 8) * 256 + (k / 16) * 128 + (j % 8) * 16 + k % 16] = (unsigned char)B[j][k];
  unsigned char *dA, *dB; int *dTS, *dSS; unsigned long long* dC;
  cudaMalloc(&dA, 4096); cudaMalloc(&dB, N * 32); cudaMalloc(&dTS, 128 * N * 4); cudaMalloc(&dSS, 128 * N * 4);
  cudaMalloc(&dC, 8 * sms);
  cudaMemcpy(dA, imgA, 4096, cudaMemcpyHostToDevice); cudaMemcpy(dB, imgB, N * 32, cudaMemcpyHostToDevice);
  cudaMemset(dTS, 0xff, 128 * N * 4); cudaMemset(dSS, 0xff, 128 * N * 4);
  probe<<<1, 128, 6 * 4096 + 6 * N * 32>>>(N, 0, 1, dA, dB, dTS, dSS, dC);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("CUDA error in verify: %s\n", cudaGetErrorString(err)); return 1; }
  static int hTS[128 * N], hSS[128 * N];
  cudaMemcpy(hTS, dTS, sizeof(hTS), cudaMemcpyDeviceToHost); cudaMemcpy(hSS, dSS, sizeof(hSS), cudaMemcpyDeviceToHost);
  int badTS = 0, badSS = 0;
  for (int i = 0; i < 128; ++i) for (int j = 0; j < N; ++j) {
    int want = 0; for (int k = 0; k < 32; ++k) want += (int)A[i][k] * (int)B[j][k];
    badTS += hTS[i * N + j] != want; badSS += hSS[i * N + j] != want;
  }
  printf("{\"verify\": \"cp.128x256b + TS mma\", \"mismatch_ts\": %d, \"mismatch_ss\": %d, \"of\": %d}\n", badTS, badSS, 128 * N);
  if (badTS) {  // diagnose: does TS equal the product with some permutation of k within a row?
    printf("first rows TS vs want: ");
    for (int j = 0; j < 4; ++j) { int want = 0; for (int k = 0; k < 32; ++k) want += (int)A[0][k] * (int)B[j][k]; printf("%d/%d ", hTS[j], want); }
    printf("\n");
  }
  // ---- timing
  unsigned long long* h = (unsigned long long*)malloc(8 * sms);
  const int iters = 2000;
  for (int n : {64}) for (int mode : {1, 2, 3, 4, 5, 6, 7}) for (int grid : {sms}) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t smem = 6 * 4096 + 6 * n * 32;
    probe<<<grid, 128, smem>>>(n, mode, 50, dA, dB, dTS, dSS, dC);
    cudaEventRecord(e0);
    probe<<<grid, 128, smem>>>(n, mode, iters, dA, dB, dTS, dSS, dC);
    cudaEventRecord(e1);
    err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error: %s (N=%d mode=%d)\n", cudaGetErrorString(err), n, mode); return 1; }
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(h, dC, 8 * grid, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < grid; ++i) c += (double)h[i]; c /= grid;
    printf("{\"N\": %d, \"mode\": \"%s\", \"ctas\": %d, \"clk_per_kstep\": %.1f, \"ideal_mma_clk_per_kstep\": %.1f, \"ms\": %.3f}\n", n,
           mode == 1 ? "6cp+21mma_ts" : mode == 2 ? "6cp" : mode == 3 ? "21mma_ts" : mode == 4 ? "21mma_ts+commit" : mode == 5 ? "21mma_ts+fence_after" : mode == 6 ? "21mma_ts+try_wait" : "21mma_ts+wait+fence+commit", grid, c / iters, 21 * n / 2.0, ms);
  }
  return 0;
}
