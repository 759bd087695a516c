// pipe_probe.cu -- do DMMA.8x8x4 and DFMA share an execution pipe on B200?
// One CTA per SM, 8 warps.  Mode bits: 1 = warps 0-3 run DMMA chains, 2 = warps 4-7 run DFMA chains.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 1) probe(double* out, int iters, int mode, int dmma_warps) {
  const int warp = threadIdx.x >> 5;
  double acc[16][2];
  for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = threadIdx.x * 1e-9;
  double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  if (warp < dmma_warps) {
    if (mode & 1) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
      }
    }
  } else {
    if (mode & 2) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          acc[i][0] = fma(acc[i][0], a, b);
          acc[i][1] = fma(acc[i][1], a, b);
        }
      }
    }
  }
  double s = 0;
  for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}

static float run(double* out, int iters, int mode, int dw) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<<<148, 256>>>(out, iters, mode, dw);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe<<<148, 256>>>(out, iters, mode, dw);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  double* out; cudaMalloc(&out, 148 * 256 * 8);
  const int iters = 20000;
  for (int dw : {4, 8}) {
    const double dmma_flops = 148.0 * dw * iters * 16 * 512;          // 256 FMA = 512 flop per DMMA
    const double dfma_flops = 148.0 * (8 - dw) * 32.0 * iters * 32 * 2;  // 32 DFMA per thread-iter
    float t1 = run(out, iters, 1, dw), t2 = dw < 8 ? run(out, iters, 2, dw) : 0.f, t3 = dw < 8 ? run(out, iters, 3, dw) : 0.f;
    printf("dmma_warps=%d  DMMA alone %.3f ms (%.2f TF/s)", dw, t1, dmma_flops / t1 / 1e9);
    if (dw < 8) printf(" | DFMA alone %.3f ms (%.2f TF/s) | both %.3f ms (DMMA %.2f + DFMA %.2f TF/s)", t2,
                       dfma_flops / t2 / 1e9, t3, dmma_flops / t3 / 1e9, dfma_flops / t3 / 1e9);
    printf("\n");
  }
  return 0;
}
