// factor.cuh -- host drivers of the training-side factorisation (factor.cu).
#pragma once
#include "common.cuh"

namespace bo {

// K (m, ldk, ldk): RBF Gram over rows/cols [last_eval, npad_rows); indices >= n get identity padding.
int gram(double* K, long long ldk, long long strideK, const double* x, int ldx, int last_eval, int n, int npad_rows,
         int d, int m, const ObjParams& hp, double diag_add, cudaStream_t stream, bool lower_only = false);

// In-place lower Cholesky of `batch` npad x npad matrices (npad % 64 == 0).  D receives the inverted 64x64
// diagonal blocks (npad/64 blocks of 4096 doubles per matrix).  info holds 2*batch ints (zeroed by the caller):
// info[b] = 1-based pivot where the matrix proved indefinite (0 = fine), info[batch+b] = pivots clamped to the
// floor.  pol: 2*batch doubles of scratch.  The pivot floor of matrix b is jit_dev[b / per_setting] when jit_dev
// is given, else jit_scalar (see potf2_kernel).
int cholesky_blocked(double* A, long long lda, long long strideA, int npad, int batch, double* D, long long strideD,
                     int* info, double* pol, const double* jit_dev, double jit_scalar, int per_setting,
                     cudaStream_t stream);

// W = L^-1 (lower), T is scratch of npad*npad/2 doubles per matrix.
int tri_inverse(double* W, long long ldw, long long strideW, const double* L, long long ldl, long long strideL,
                const double* D, long long strideD, double* T, long long strideT, int npad, int batch,
                cudaStream_t stream);

// alpha[o] = W_o^T (W_o (y[:,o] - mu0_o)), (m, npad)
int compute_alpha(double* alpha, const double* W, long long ldw, long long strideW, const double* y, int ldy, int n,
                  int npad, int m, const ObjParams& hp, double* scratch, cudaStream_t stream);
size_t alpha_scratch_doubles(int npad, int m);

// Incremental update: rows [n_old, n_new) of the padded dense L and W = L^-1 (both npad x npad, the
// state bo_gp_fit_f64 leaves in its workspace) are computed from the existing factor of the first n_old points.
// b = n_new - n_old <= BO_MAX_APPEND, and n_new <= npad.  info: 2*m ints (failed pivot, clamped pivots).
size_t append_scratch_doubles(int npad, int m);
int append_rows(double* L, double* W, long long ld, long long stride, const double* x, int ldx, int n_old, int n_new,
                int npad, int d, int m, const ObjParams& hp, double jitter, double* scratch, int* info,
                cudaStream_t stream);

// W -> fragment-ordered 16 KB tiles (see common.cuh)
int pack_w(double* Wp, long long strideWp, const double* W, long long ldw, long long strideW, int npad, int n, int m,
           cudaStream_t stream);

}  // namespace bo
