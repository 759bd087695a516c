// gemm.cuh -- host interface of the general DMMA GEMM (gemm.cu).
#pragma once
#include "common.cuh"

namespace bo {

struct GemmArgs {
  int M = 0, N = 0, K = 0;
  double alpha = 1.0, beta = 0.0;
  const double* A = nullptr;
  long long lda = 0, strideA = 0;
  const double* B = nullptr;
  long long ldb = 0, strideB = 0;
  double* C = nullptr;
  long long ldc = 0, strideC = 0;
  int batch = 1;
  int lower_only = 0;    // skip tiles strictly above the diagonal (SYRK-style updates)
  int k_limit_rows = 0;  // opA is lower triangular (amode 0): stop the k loop at the tile's last row
  int k_start_cols = 0;  // opB is lower triangular (bmode 1): start the k loop at the tile's first column
  int fast = 0;          // set by gemm()
};

// amode 0: A(i,k) = A[i*lda+k]; amode 1: A(i,k) = A[k*lda+i]
// bmode 0: B(k,j) = B[j*ldb+k]; bmode 1: B(k,j) = B[k*ldb+j]
int gemm(const GemmArgs& args, int amode, int bmode, cudaStream_t stream);

}  // namespace bo
