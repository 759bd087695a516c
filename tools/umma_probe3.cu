// umma_probe3.cu -- prototype of the "register path" for the TMEM A operand: 8 feeder warps copy the six
// 128 x 32-byte A planes shared memory -> registers -> TMEM (ld.shared.v4 + tcgen05.st.32x32b.x8) into one of two
// slots while a single thread issues the 21 TS-mode kind::i8 MMAs of the Ozaki k-step on the other slot.
// Handshake: a_ready[slot] (8 warp arrivals) and slot_free[slot] (tcgen05.commit).  Checks exactness of the
// accumulators against the CPU and reports cycles per k-step; also try_wait vs test_wait latency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe3 tools/umma_probe3.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define TN 64
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n"
               ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc_of(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)8 << 16) | ((uint64_t)16 << 32) | (1ull << 46);
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <bool TEST>
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    if (TEST)
      asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    else
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void sttm8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

// mode 0: feeders + MMAs (the prototype); mode 1: same with test_wait polling; mode 2: MMAs only (no handshake)
__global__ void __launch_bounds__(352, 1) probe(int mode, int iters, const unsigned char* gA, const unsigned char* gB, int* outD,
                                                unsigned long long* cyc) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t done, done2, a_ready[2], slot_free[2], turn[2];
  __shared__ uint32_t slot;
  unsigned char* sA = smem;             // 6 planes x 4096 (canonical K-major image)
  unsigned char* sB = smem + 6 * 4096;  // 6 planes x 2048
  for (int i = threadIdx.x; i < 6 * 4096; i += blockDim.x) sA[i] = gA[i];
  for (int i = threadIdx.x; i < 6 * 2048; i += blockDim.x) sB[i] = gB[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done2)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&turn[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&turn[1])));
    for (int s = 0; s < 2; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(smem_u32(&a_ready[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&slot_free[s])));
    }
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
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | (8u << 24);
  const uint32_t b0 = smem_u32(sB);
  const uint32_t a_tm = tm + 6 * TN;
  if (warp == 1 || (warp == 10 && mode == 3)) {
    uint32_t e;
    asm volatile("{\n.reg .pred px;\n.reg .b32 rx;\nelect.sync rx|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}\n" : "=r"(e));
    if (e) {
      const int me = warp == 1 ? 0 : 1;  // issuer 0: even k-steps (slot 0), issuer 1: odd k-steps (slot 1)
      const long long t0 = clock64();
      for (int it = (mode == 3 ? me : 0); it < iters; it += (mode == 3 ? 2 : 1)) {
        const int sl = it & 1;
        const uint32_t as = a_tm + sl * 48;
        if (mode != 2) {
          if (mode == 1) wait_bar<true>(&a_ready[sl], (it >> 1) & 1); else wait_bar<false>(&a_ready[sl], (it >> 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        if (mode == 3 && it > 0) wait_bar<false>(&turn[me], ((it - 1) >> 1) & 1);  // the other issuer handed over
        if (mode == 3) {
          // the six overwrite-capable MMAs (s = 0) first, the token after 17 of 21
          int cnt = 0;
#pragma unroll
          for (int g = 0; g < 6; ++g) { umma_ts(tm + g * TN, as, desc_of(b0 + g * 2048), idesc, it ? 1u : 0u); ++cnt; }
#pragma unroll
          for (int g = 1; g < 6; ++g)
#pragma unroll
            for (int s = 1; s <= g; ++s) {
              umma_ts(tm + g * TN, as + s * 8, desc_of(b0 + (g - s) * 2048), idesc, 1u);
              if (++cnt == 17) arrive(&turn[me ^ 1]);
            }
        } else {
#pragma unroll
          for (int g = 0; g < 6; ++g)
#pragma unroll
            for (int s = 0; s <= g; ++s)
              umma_ts(tm + g * TN, as + s * 8, desc_of(b0 + (g - s) * 2048), idesc, (it | s) ? 1u : 0u);
        }
        if (mode != 2) commit(&slot_free[sl]);
      }
      commit(me ? &done2 : &done);
      wait_bar<false>(me ? &done2 : &done, 0);
      if (me == 0) cyc[blockIdx.x] = (unsigned long long)(clock64() - t0);
    }
    __syncwarp();
  } else if (warp >= 2 && warp < 10) {
    // feeders: quarter = warp % 4 (TMEM lanes), half = (warp - 2) / 4 (planes 3*half .. 3*half + 2)
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const unsigned char* src = sA + (row >> 3) * 256 + (row & 7) * 16;
    const uint32_t lane_addr = ((uint32_t)(quarter * 32) << 16);
    for (int it = 0; it < iters; ++it) {
      const int sl = it & 1;
      if (mode == 2) {
        if (it >= 2) break;  // fill both slots once, no handshake
      } else if (it >= 2) {
        if (mode == 1) wait_bar<true>(&slot_free[sl], ((it >> 1) - 1) & 1); else wait_bar<false>(&slot_free[sl], ((it >> 1) - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        const int plane = half * 3 + p;
        const uint4 lo = *reinterpret_cast<const uint4*>(src + plane * 4096);        // k 0..15
        const uint4 hi = *reinterpret_cast<const uint4*>(src + plane * 4096 + 128);  // k 16..31
        const uint32_t v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        sttm8(tm + lane_addr + 6 * TN + sl * 48 + plane * 8, v);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0 && mode != 2) arrive(&a_ready[sl]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (outD && warp >= 2 && warp < 6) {
    const int quarter = warp & 3;
    for (int c = 0; c < 6 * TN; ++c) {
      int v;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tm + ((uint32_t)(quarter * 32) << 16) + c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      outD[(quarter * 32 + lane) * 6 * TN + c] = v;
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
  const size_t smem = 6 * 4096 + 6 * 2048;
  static signed char A[6][128][32], B[6][TN][32];
  srand(3);
  for (int p = 0; p < 6; ++p) {
    for (int i = 0; i < 128; ++i) for (int k = 0; k < 32; ++k) A[p][i][k] = (signed char)(rand() % 256 - 128);
    for (int j = 0; j < TN; ++j) for (int k = 0; k < 32; ++k) B[p][j][k] = (signed char)(rand() % 256 - 128);
  }
  static unsigned char imgA[6 * 4096], imgB[6 * 2048];
  for (int p = 0; p < 6; ++p) {
    for (int i = 0; i < 128; ++i) for (int k = 0; k < 32; ++k) imgA[p * 4096 + (i / 8) * 256 + (k / 16) * 128 + (i % 8) * 16 + k % 16] = (unsigned char)A[p][i][k];
    for (int j = 0; j < TN; ++j) for (int k = 0; k < 32; ++k) imgB[p * 2048 + (j / 8) * 256 + (k / 16) * 128 + (j % 8) * 16 + k % 16] = (unsigned char)B[p][j][k];
  }
  unsigned char *dA, *dB; int* dD; unsigned long long* dC;
  cudaMalloc(&dA, sizeof(imgA)); cudaMalloc(&dB, sizeof(imgB)); cudaMalloc(&dD, 128 * 6 * TN * 4); cudaMalloc(&dC, 8 * sms);
  cudaMemcpy(dA, imgA, sizeof(imgA), cudaMemcpyHostToDevice); cudaMemcpy(dB, imgB, sizeof(imgB), cudaMemcpyHostToDevice);
  // ---- exactness: 7 k-steps with the handshake
  cudaError_t err;
  for (int vmode : {0, 3}) {
  const int vit = vmode == 3 ? 400 : 7;
  probe<<<1, 352, smem>>>(vmode, vit, dA, dB, dD, dC);
  err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("CUDA error in verify: %s\n", cudaGetErrorString(err)); return 1; }
  static int hD[128 * 6 * TN];
  cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int g = 0; g < 6; ++g) for (int i = 0; i < 128; ++i) for (int j = 0; j < TN; ++j) {
    long want = 0;
    for (int s = 0; s <= g; ++s) for (int k = 0; k < 32; ++k) want += (long)A[s][i][k] * (long)B[g - s][j][k];
    bad += hD[i * 6 * TN + g * TN + j] != (int)(want * vit);
  }
  printf("{\"verify\": \"lds + tcgen05.st feeders + TS mma, %d k-steps, %s\", \"mismatch\": %ld, \"of\": %d}\n", vit, vmode == 3 ? "two issuers" : "one issuer", bad, 128 * 6 * TN);
  }
  // ---- timing
  unsigned long long* h = (unsigned long long*)malloc(8 * sms);
  const int iters = 4000;
  for (int mode : {0, 3, 2}) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<<<sms, 352, smem>>>(mode, 100, dA, dB, nullptr, dC);
    cudaEventRecord(e0);
    probe<<<sms, 352, smem>>>(mode, iters, dA, dB, nullptr, dC);
    cudaEventRecord(e1);
    err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error: %s (mode=%d)\n", cudaGetErrorString(err), mode); return 1; }
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(h, dC, 8 * sms, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < sms; ++i) c += (double)h[i]; c /= sms;
    printf("{\"mode\": \"%s\", \"clk_per_kstep\": %.1f, \"ideal\": 672, \"ms\": %.3f}\n",
           mode == 0 ? "feeders+mma, one issuer" : mode == 3 ? "feeders+mma, two issuers alternating" : "mma only", c / iters, ms);
  }
  return 0;
}
