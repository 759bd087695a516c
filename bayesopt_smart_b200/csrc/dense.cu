// dense.cu -- function-level drop-ins that exchange the reference's materialised arrays:
//   update_k_star  (numba_kernels.py:406-442)  -> kstar_dense
//   update_mean    (numba_kernels.py:450-488)  -> mean_dense     (two mat-vecs)
//   update_variance(numba_kernels.py:491-535)  -> variance_dense (DMMA GEMM Kinv @ K*, then column dots)
// The fused path in score.cu never materialises K*; these exist so that callers who use the reference's
// functions one by one keep working.
#include "dense.cuh"
#include "gemm.cuh"
#include "rbf.cuh"

namespace bo {

namespace {

template <typename CT>
__global__ void kstar_dense_kernel(double* __restrict__ ks, long long ld_row, long long ld_obj,
                                   const double* __restrict__ x, int ldx, const CT* __restrict__ cand, int ldc,
                                   long long n_cand, int last_eval, int current_eval, int d, int m, ObjParams hp) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int e = last_eval + blockIdx.y;
  if (c >= n_cand || e >= current_eval) return;
  double sq = 0.0;
  for (int k = 0; k < d; ++k) {
    const double diff = x[(long long)e * ldx + k] - (double)cand[c * ldc + k];
    sq = fma(diff, diff, sq);
  }
  for (int o = 0; o < m; ++o)
    ks[o * ld_obj + (long long)e * ld_row + c] = hp.prior_var[o] * rbf_exp(sq * hp.neg_half_inv_ls2[o], kExp2Tab);
}

// t[o][i] = sum_k Kinv[o][i][k] * (y[k][o] - mu0[o])      (one warp per row)
__global__ void kinv_delta_kernel(double* __restrict__ t, const double* __restrict__ kinv, int ld_kinv,
                                  long long ld_kinv_obj, const double* __restrict__ y, int ldy, int n, ObjParams hp) {
  const int o = blockIdx.y;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const double* r = kinv + o * ld_kinv_obj + (long long)row * ld_kinv;
  double s = 0.0;
  for (int k = lane; k < n; k += 32) s = fma(r[k], y[(long long)k * ldy + o] - hp.prior_mean[o], s);
#pragma unroll
  for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) t[(long long)o * n + row] = s;
}

// mu[o][c] = mu0[o] + sum_e K*[o][e][c] * t[o][e]       (thread per candidate, coalesced over c)
__global__ void mean_apply_kernel(double* __restrict__ mu, long long ld_mu, const double* __restrict__ ks,
                                  long long ld_row, long long ld_obj, const double* __restrict__ t, int n,
                                  long long n_cand, ObjParams hp) {
  const int o = blockIdx.y;
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cand) return;
  const double* k = ks + o * ld_obj + c;
  const double* to = t + (long long)o * n;
  double s = 0.0;
  for (int e = 0; e < n; ++e) s = fma(k[(long long)e * ld_row], to[e], s);
  mu[o * ld_mu + c] = hp.prior_mean[o] + s;
}

// var[c] = max(var0 - sum_e K*[e][c] * T[e][c], min_var)
__global__ void coldot_variance_kernel(double* __restrict__ var, const double* __restrict__ ks, long long ld_row,
                                       const double* __restrict__ T, long long ld_t, int n, long long n_chunk,
                                       double var0, double min_variance) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunk) return;
  double s = 0.0;
  for (int e = 0; e < n; ++e) s = fma(ks[(long long)e * ld_row + c], T[(long long)e * ld_t + c], s);
  var[c] = fmax(var0 - s, min_variance);
}

constexpr long long DENSE_CHUNK = 1 << 15;

}  // namespace

int kstar_dense(double* ks, long long ld_row, long long ld_obj, const double* x, int ldx, const void* cand,
                int cand_kind, int ldc, long long n_cand, int last_eval, int current_eval, int d, int m,
                const ObjParams& hp, cudaStream_t stream) {
  const int rows = current_eval - last_eval;
  if (rows <= 0 || n_cand <= 0) return BO_OK;
  // gridDim.y is limited to 65535 rows per launch
  for (int r0 = 0; r0 < rows; r0 += 65535) {
    const int nr = rows - r0 < 65535 ? rows - r0 : 65535;
    dim3 grid((unsigned)((n_cand + 255) / 256), nr);
    if (cand_kind == BO_CAND_I64)
      kstar_dense_kernel<long long><<<grid, 256, 0, stream>>>(ks, ld_row, ld_obj, x, ldx,
                                                              static_cast<const long long*>(cand), ldc, n_cand,
                                                              last_eval + r0, current_eval, d, m, hp);
    else
      kstar_dense_kernel<double><<<grid, 256, 0, stream>>>(ks, ld_row, ld_obj, x, ldx,
                                                           static_cast<const double*>(cand), ldc, n_cand,
                                                           last_eval + r0, current_eval, d, m, hp);
    BO_LAUNCH_CHECK("kstar_dense_kernel");
  }
  return BO_OK;
}

size_t dense_workspace_bytes(int n, long long n_cand) {
  const long long ch = n_cand < DENSE_CHUNK ? n_cand : DENSE_CHUNK;
  return align256((size_t)n * (size_t)(ch > 0 ? ch : 1) * sizeof(double)) +
         align256((size_t)BO_MAX_OBJECTIVES * n * sizeof(double));
}

int mean_dense(double* mu, long long ld_mu, const double* ks, long long ld_row, long long ld_obj, const double* kinv,
               int ld_kinv, long long ld_kinv_obj, const double* y, int ldy, const ObjParams& hp, int n,
               long long n_cand, int m, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (workspace_bytes < align256((size_t)m * n * sizeof(double))) {
    set_error("dense workspace too small");
    return BO_ERR_WORKSPACE;
  }
  double* t = static_cast<double*>(workspace);
  kinv_delta_kernel<<<dim3((n + 7) / 8, m), 256, 0, stream>>>(t, kinv, ld_kinv, ld_kinv_obj, y, ldy, n, hp);
  BO_LAUNCH_CHECK("kinv_delta_kernel");
  mean_apply_kernel<<<dim3((unsigned)((n_cand + 255) / 256), m), 256, 0, stream>>>(mu, ld_mu, ks, ld_row, ld_obj, t, n,
                                                                                   n_cand, hp);
  BO_LAUNCH_CHECK("mean_apply_kernel");
  return BO_OK;
}

int variance_dense(double* var, long long ld_var, const double* ks, long long ld_row, long long ld_obj,
                   const double* kinv, int ld_kinv, long long ld_kinv_obj, const ObjParams& hp, double min_variance,
                   int n, long long n_cand, int m, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (workspace_bytes < dense_workspace_bytes(n, n_cand)) {
    set_error("dense workspace too small");
    return BO_ERR_WORKSPACE;
  }
  double* T = static_cast<double*>(workspace);
  for (int o = 0; o < m; ++o) {
    for (long long c0 = 0; c0 < n_cand; c0 += DENSE_CHUNK) {
      const long long nc = n_cand - c0 < DENSE_CHUNK ? n_cand - c0 : DENSE_CHUNK;
      GemmArgs g;  // T = Kinv_o @ K*_o[:, c0:c0+nc]     (numba_kernels.py:521)
      g.M = n; g.N = (int)nc; g.K = n;
      g.A = kinv + o * ld_kinv_obj; g.lda = ld_kinv;
      g.B = ks + o * ld_obj + c0; g.ldb = ld_row;
      g.C = T; g.ldc = nc;
      int rc = gemm(g, 0, 1, stream);
      if (rc) return rc;
      coldot_variance_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, stream>>>(
          var + o * ld_var + c0, ks + o * ld_obj + c0, ld_row, T, nc, n, nc, hp.prior_var[o], min_variance);
      BO_LAUNCH_CHECK("coldot_variance_kernel");
    }
  }
  return BO_OK;
}

}  // namespace bo
