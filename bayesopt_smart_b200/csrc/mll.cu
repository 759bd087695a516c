// mll.cu -- batched log marginal likelihood over many (length-scale, jitter) settings.
// Reference: compute_mll, numba_kernels.py:152-235.  Per setting s and objective o:
//   R = exp(-0.5 |xi-xj|^2 / ls^2) + jit I          (the reference divides K by prior_variance: :195-197)
//   yt = (y - mu0) / std(y - mu0)                    (:201-208, population std, skipped when 0)
//   L = chol(R)                                      (:211-214)      -> cholesky_blocked (DMMA)
//   mll = -0.5 |L^-1 yt|^2 - sum(log diag L) - 0.5 n log(2 pi)     (:216-232; yt.alpha == |L^-1 yt|^2)
#include "factor.cuh"
#include "mll.cuh"
#include "rbf.cuh"

namespace bo {

namespace {

__device__ __forceinline__ double block_sum(double v, double* scratch) {
  // deterministic tree: warp shuffles then the first warp over the 8 partials
#pragma unroll
  for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += scratch[w];
  return s;
}

// one CTA per objective: standardised targets, zero padded to npad
__global__ void __launch_bounds__(256) mll_prepare_y_kernel(double* __restrict__ yt, const double* __restrict__ y,
                                                            int ldy, int n, int npad, ObjParams hp) {
  __shared__ double scratch[8];
  const int o = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += y[(long long)i * ldy + o] - hp.prior_mean[o];
  const double mean = block_sum(s, scratch) / n;
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double c = (y[(long long)i * ldy + o] - hp.prior_mean[o]) - mean;
    v = fma(c, c, v);
  }
  const double sd = sqrt(block_sum(v, scratch) / n);
  for (int i = threadIdx.x; i < npad; i += blockDim.x) {
    double c = 0.0;
    if (i < n) {
      c = y[(long long)i * ldy + o] - hp.prior_mean[o];
      if (sd > 0.0) c /= sd;
    }
    yt[(long long)o * npad + i] = c;
  }
}

// one CTA per matrix: forward substitution z = L^-1 yt using the inverted 64x64 diagonal blocks,
// then the three MLL terms.  z lives in shared memory.
// 32 warps per matrix: the substitution streams the whole lower triangle of L once (64 MB at n = 4096), and with one
// CTA per matrix its speed is set by how many loads that CTA keeps in flight (8 warps: 4.5 ms per group of matrices).
constexpr int SOLVE_THREADS = 1024;
__global__ void __launch_bounds__(SOLVE_THREADS)
    mll_solve_kernel(double* __restrict__ mll_obj, const double* __restrict__ L, long long ldl, long long strideL,
                     const double* __restrict__ D, long long strideD, const int* __restrict__ info,
                     const double* __restrict__ yt, int n, int npad, int m) {
  extern __shared__ double zsm[];  // npad (z) + 64 (rhs) + 32 (scratch: one partial per warp)
  double* z = zsm;
  double* rhs = zsm + npad;
  double* scratch = rhs + 64;
  const int b = blockIdx.x;
  const int o = b % m;
  const double* Lb = L + (long long)b * strideL;
  const double* Db = D + (long long)b * strideD;
  const double* yo = yt + (long long)o * npad;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = npad / 64;
  for (int jb = 0; jb < nblk; ++jb) {
    const int r0 = jb * 64;
    // rhs[r] = yt[r0+r] - sum_{k<r0} L[r0+r][k] z[k]; warp w owns rows 2w, 2w+1
    constexpr int ROWS_PER_WARP = 64 / (SOLVE_THREADS / 32);
    for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
      const int r = warp * ROWS_PER_WARP + rr;
      const double* Lr = Lb + (long long)(r0 + r) * ldl;
      double s = 0.0;
      for (int k = lane; k < r0; k += 32) s = fma(Lr[k], z[k], s);
#pragma unroll
      for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (lane == 0) rhs[r] = yo[r0 + r] - s;
    }
    __syncthreads();
    // z[r0 + r] = sum_{c<=r} Dinv[r][c] rhs[c]; 4 threads per row (the first 256 threads)
    if (threadIdx.x < 256) {
      const int r = threadIdx.x >> 2, q = threadIdx.x & 3;
      const double* Dr = Db + (long long)jb * 4096 + r * 64;
      double s = 0.0;
      for (int c = q; c <= r; c += 4) s = fma(Dr[c], rhs[c], s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (q == 0) z[r0 + r] = s;
    }
    __syncthreads();
  }
  double fit = 0.0, logdiag = 0.0;
  for (int i = threadIdx.x; i < npad; i += blockDim.x) {
    fit = fma(z[i], z[i], fit);
    logdiag += log(Lb[(long long)i * (ldl + 1)]);
  }
  fit = block_sum(fit, scratch);
  logdiag = block_sum(logdiag, scratch);
  if (threadIdx.x == 0) {
    double v = -0.5 * fit - logdiag - 0.5 * n * log(2.0 * 3.14159265358979323846);
    if (info[b] != 0) v = __longlong_as_double(0x7ff8000000000000ll);
    mll_obj[b] = v;
  }
}

__global__ void mll_sum_kernel(double* __restrict__ out, const double* __restrict__ mll_obj, int n_settings, int m) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_settings) return;
  double t = 0.0;
  for (int o = 0; o < m; ++o) t += mll_obj[(long long)s * m + o];  // np.sum over objectives (:235)
  out[s] = t;
}

// ------------------------------------------------------------------------------------------- small n
// n <= 128: the whole evaluation of one setting (all objectives, in sequence) in ONE CTA -- Gram matrix,
// Cholesky, forward substitution and the three MLL terms live in shared memory.  This is the shape of the
// reference's own use (tens of evaluated points, hundreds of sequential Powell evaluations per iteration:
// numba_kernels.py:305-315), where launch latency, not arithmetic, is the cost.
constexpr int SMALL_ROWS = 128;        // rows of the shared-memory matrix: n training rows + the target row
constexpr int SMALL_N = SMALL_ROWS - 1;

// One CTA per (setting, objective); two threads per matrix row (columns of equal parity).  The standardised
// targets ride along as row n of the matrix: a right-looking Cholesky leaves z = L^-1 yt in that row, so there is
// no separate forward substitution.  Column j costs ONE barrier: the trailing update uses the unscaled column
// and the pivot, a_ik -= a_ij a_kj / piv (the scaled column itself is never needed -- only its diagonal, for the
// log-determinant, and its target-row entry, for the fit term).  The last CTA of a setting adds the per-objective
// values in objective order (deterministic), so the whole evaluation is one launch.
// Hyper-parameters of up to SMALL_INLINE settings travel in the kernel parameters (a Powell evaluation is one
// setting: no host-to-device copies on that path).
constexpr int SMALL_INLINE = 4;
struct SmallHyper {
  double ls[SMALL_INLINE * BO_MAX_OBJECTIVES];
  double jit[SMALL_INLINE];
};

__global__ void __launch_bounds__(2 * SMALL_ROWS)
    mll_small_kernel(double* __restrict__ out, double* __restrict__ vals, unsigned int* __restrict__ done,
                     const double* __restrict__ x, int ldx, const double* __restrict__ y, int ldy, int n, int d, int m,
                     ObjParams hp0, const double* __restrict__ ls_dev, const double* __restrict__ jit_dev,
                     const __grid_constant__ SmallHyper inl) {
  const double* ls_all = ls_dev ? ls_dev : inl.ls;
  const double* jit_all = jit_dev ? jit_dev : inl.jit;
  extern __shared__ double sm[];
  double(*A)[SMALL_ROWS + 1] = reinterpret_cast<double(*)[SMALL_ROWS + 1]>(sm);  // rows 0..n-1: Gram; row n: targets
  double* xs = sm + SMALL_ROWS * (SMALL_ROWS + 1);                                // n x d coordinates
  double* rinv = xs + SMALL_ROWS * BO_MAX_DIMS;                                   // 1 / l_jj per column
  double* scratch = rinv + SMALL_ROWS;                                            // 8 doubles
  __shared__ int nan_flag;
  __shared__ unsigned int ticket;
  const int tid = threadIdx.x;
  const int row = tid >> 1, q = tid & 1;
  const int s = blockIdx.x, o = blockIdx.y;
  const double jit = jit_all[s];
  if (tid == 0) nan_flag = 0;
  for (int e = tid; e < n * d; e += 2 * SMALL_ROWS) xs[e] = x[(long long)(e / d) * ldx + (e % d)];
  // ---- standardised targets (population std of y - mu0, skipped when 0: numba_kernels.py:201-208)
  double acc = 0.0;
  for (int i = tid; i < n; i += 2 * SMALL_ROWS) acc += y[(long long)i * ldy + o] - hp0.prior_mean[o];
  const double mean = block_sum(acc, scratch) / n;  // (block_sum synchronises: xs is visible afterwards)
  acc = 0.0;
  for (int i = tid; i < n; i += 2 * SMALL_ROWS) {
    const double c = (y[(long long)i * ldy + o] - hp0.prior_mean[o]) - mean;
    acc = fma(c, c, acc);
  }
  const double sd = sqrt(block_sum(acc, scratch) / n);
  // ---- lower triangle of the correlation matrix + jitter; the target row
  const double ls = ls_all[(long long)s * m + o];
  const double coef = -0.5 / (ls * ls);
  if (row < n) {
    for (int k = q; k <= row; k += 2) {
      double sq = 0.0;
      for (int c = 0; c < d; ++c) {
        const double diff = xs[row * d + c] - xs[k * d + c];
        sq = fma(diff, diff, sq);
      }
      A[row][k] = rbf_exp(sq * coef, kExp2Tab) + (row == k ? jit : 0.0);
    }
  } else if (row == n) {
    for (int k = q; k < n; k += 2) {
      double c = y[(long long)k * ldy + o] - hp0.prior_mean[o];
      if (sd > 0.0) c /= sd;
      A[n][k] = c;
    }
  }
  // ---- right-looking Cholesky, one barrier per column; same pivot policy as potf2_kernel
  const double neg_tol = 1.4901161193847656e-08 * (1.0 + jit);
  for (int j = 0; j < n; ++j) {
    __syncthreads();  // the update of column j (by the previous step) is complete
    double piv = A[j][j];
    if (!(piv >= jit)) {
      if (!(piv > -neg_tol) && tid == 0) nan_flag = 1;
      piv = fmax(jit, 2.220446049250313e-16);
    }
    const double r = rsqrt(piv);  // 1 / l_jj, computed redundantly by every thread (no second barrier)
    if (tid == 0) rinv[j] = r;
    if (row > j && row <= n) {
      const double l = A[row][j] * (r * r);  // a_ij / piv
      const int kend = row < n ? row : n - 1;  // the target row has no diagonal entry
      for (int k = j + 1 + ((j + 1 + q) & 1); k <= kend; k += 2) A[row][k] = fma(-l, A[k][j], A[row][k]);
    }
  }
  __syncthreads();
  // z_j = a_nj / l_jj;  log det L = -sum log(1 / l_jj)
  double fit = 0.0, logdiag = 0.0;
  for (int j = tid; j < n; j += 2 * SMALL_ROWS) {
    const double z = A[n][j] * rinv[j];
    fit = fma(z, z, fit);
    logdiag -= log(rinv[j]);
  }
  fit = block_sum(fit, scratch);
  logdiag = block_sum(logdiag, scratch);
  if (tid == 0) {
    const double v = -0.5 * fit - logdiag - 0.5 * n * log(2.0 * 3.14159265358979323846);
    vals[(long long)s * m + o] = nan_flag ? __longlong_as_double(0x7ff8000000000000ll) : v;
    __threadfence();
    ticket = atomicInc(&done[s], (unsigned int)(m - 1));  // wraps to 0 after the m-th arrival: self-resetting
  }
  __syncthreads();
  if (tid == 0 && ticket == (unsigned int)(m - 1)) {
    __threadfence();
    double total = 0.0;
    for (int oo = 0; oo < m; ++oo) total += reinterpret_cast<volatile double*>(vals)[(long long)s * m + oo];
    out[s] = total;  // NaN if any objective was not positive definite
  }
}

}  // namespace

static int group_settings(int npad, int m, int n_settings) {
  // settings factored together: the matrices of one group take up to 1/8 of the device memory (22 GB on a
  // B200) -- the 64-column steps of the blocked Cholesky and the forward substitution are latency chains whose
  // cost per group does not depend on how many matrices ride along
  const size_t per = (size_t)m * npad * npad * sizeof(double);
  // per device ordinal (a process may drive GPUs of different sizes)
  static size_t budgets[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  size_t budget = budgets[dev];
  if (budget == 0) {
    size_t free_b = 0, total_b = 0;
    budget = (size_t)4 << 30;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b / 8 > budget) budget = total_b / 8;
    budgets[dev] = budget;
  }
  size_t g = budget / per;
  if (g < 1) g = 1;
  if (g > (size_t)n_settings) g = n_settings;
  return (int)g;
}

size_t mll_workspace_bytes(int n, int m, int n_settings) {
  const int npad = round_up(n, TM);
  const int gs = group_settings(npad, m, n_settings);
  size_t b = 0;
  b += align256((size_t)gs * m * npad * npad * sizeof(double));   // matrices
  b += align256((size_t)gs * m * npad * 64 * sizeof(double));     // inverted diagonal blocks
  b += align256((size_t)2 * gs * m * sizeof(int));                // info (failed pivot, clamped count)
  b += align256((size_t)2 * gs * m * sizeof(double));             // pivot policy
  b += align256((size_t)n_settings * sizeof(double));             // jitters on the device
  b += align256((size_t)m * npad * sizeof(double));               // standardised targets
  b += align256((size_t)n_settings * m * sizeof(double));         // per-objective values
  return b;
}

int mll_batched(double* out, const double* x, int ldx, const double* y, int ldy, int n, int d, int m,
                const double* prior_mean, const double* length_scales, const double* jitter, int n_settings,
                void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (workspace_bytes < mll_workspace_bytes(n, m, n_settings)) {
    set_error("mll workspace too small");
    return BO_ERR_WORKSPACE;
  }
  const int npad = round_up(n, TM);
  const int gs = group_settings(npad, m, n_settings);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  if (n <= SMALL_N) {
    // one CTA per (setting, objective); hyper-parameters, per-objective values and the arrival counters travel
    // through the head of the workspace
    double* ls_dev = reinterpret_cast<double*>(ws);
    double* jit_small = ls_dev + (size_t)n_settings * m;
    double* vals_small = jit_small + n_settings;
    unsigned int* done = reinterpret_cast<unsigned int*>(vals_small + (size_t)n_settings * m);
    SmallHyper inl;
    memset(&inl, 0, sizeof(inl));
    const bool inline_hyper = n_settings <= SMALL_INLINE;
    if (inline_hyper) {
      memcpy(inl.ls, length_scales, sizeof(double) * n_settings * m);
      memcpy(inl.jit, jitter, sizeof(double) * n_settings);
    } else {
      BO_CUDA(cudaMemcpyAsync(ls_dev, length_scales, sizeof(double) * n_settings * m, cudaMemcpyHostToDevice, stream));
      BO_CUDA(cudaMemcpyAsync(jit_small, jitter, sizeof(double) * n_settings, cudaMemcpyHostToDevice, stream));
    }
    BO_CUDA(cudaMemsetAsync(done, 0, sizeof(unsigned int) * n_settings, stream));
    ObjParams hps;
    memset(&hps, 0, sizeof(hps));
    for (int o = 0; o < m; ++o) hps.prior_mean[o] = prior_mean[o];
    const size_t smem =
        (size_t)(SMALL_ROWS * (SMALL_ROWS + 1) + SMALL_ROWS * BO_MAX_DIMS + SMALL_ROWS + 8) * sizeof(double);
    {
      const int rc_attr = ensure_dynamic_smem(mll_small_kernel, smem);
      if (rc_attr) return rc_attr;
    }
    mll_small_kernel<<<dim3(n_settings, m), 2 * SMALL_ROWS, smem, stream>>>(
        out, vals_small, done, x, ldx, y, ldy, n, d, m, hps, inline_hyper ? nullptr : ls_dev,
        inline_hyper ? nullptr : jit_small, inl);
    BO_LAUNCH_CHECK("mll_small_kernel");
    BO_CUDA(cudaStreamSynchronize(stream));
    return BO_OK;
  }
  size_t off = 0;
  double* A = reinterpret_cast<double*>(ws + off);    off += align256((size_t)gs * m * npad * npad * sizeof(double));
  double* D = reinterpret_cast<double*>(ws + off);    off += align256((size_t)gs * m * npad * 64 * sizeof(double));
  int* info = reinterpret_cast<int*>(ws + off);       off += align256((size_t)2 * gs * m * sizeof(int));
  double* pol = reinterpret_cast<double*>(ws + off);  off += align256((size_t)2 * gs * m * sizeof(double));
  double* jit_dev = reinterpret_cast<double*>(ws + off); off += align256((size_t)n_settings * sizeof(double));
  double* yt = reinterpret_cast<double*>(ws + off);   off += align256((size_t)m * npad * sizeof(double));
  double* vals = reinterpret_cast<double*>(ws + off);

  BO_CUDA(cudaMemcpyAsync(jit_dev, jitter, sizeof(double) * n_settings, cudaMemcpyHostToDevice, stream));
  ObjParams hp0;
  memset(&hp0, 0, sizeof(hp0));
  for (int o = 0; o < m; ++o) hp0.prior_mean[o] = prior_mean[o];
  mll_prepare_y_kernel<<<m, 256, 0, stream>>>(yt, y, ldy, n, npad, hp0);
  BO_LAUNCH_CHECK("mll_prepare_y_kernel");

  const size_t solve_smem = (size_t)(npad + 64 + 32) * sizeof(double);
  {
    const int rc_attr = ensure_dynamic_smem(mll_solve_kernel, solve_smem);
    if (rc_attr) return rc_attr;
  }
  const long long strideA = (long long)npad * npad, strideD = (long long)npad * 64;
  for (int s0 = 0; s0 < n_settings; s0 += gs) {
    const int g = (n_settings - s0 < gs) ? (n_settings - s0) : gs;
    for (int s = 0; s < g; ++s) {
      ObjParams hp = hp0;
      for (int o = 0; o < m; ++o) {
        const double ls = length_scales[(long long)(s0 + s) * m + o];
        hp.prior_var[o] = 1.0;
        hp.neg_half_inv_ls2[o] = -0.5 / (ls * ls);
      }
      int rc = gram(A + (long long)s * m * strideA, npad, strideA, x, ldx, 0, n, npad, d, m, hp, jitter[s0 + s],
                    stream, /*lower_only=*/true);
      if (rc) return rc;
    }
    BO_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * 2 * g * m, stream));
    int rc = cholesky_blocked(A, npad, strideA, npad, g * m, D, strideD, info, pol, jit_dev + s0, 0.0, m, stream);
    if (rc) return rc;
    mll_solve_kernel<<<g * m, SOLVE_THREADS, solve_smem, stream>>>(vals + (long long)s0 * m, A, npad, strideA, D, strideD, info,
                                                         yt, n, npad, m);
    BO_LAUNCH_CHECK("mll_solve_kernel");
  }
  mll_sum_kernel<<<(n_settings + 127) / 128, 128, 0, stream>>>(out, vals, n_settings, m);
  BO_LAUNCH_CHECK("mll_sum_kernel");
  BO_CUDA(cudaStreamSynchronize(stream));
  return BO_OK;
}

}  // namespace bo
