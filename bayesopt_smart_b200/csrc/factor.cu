// factor.cu -- training-side factorisation, once per BO iteration (reference: update_k + invert_k,
// numba_kernels.py:329-403).  All matrices here are npad x npad (npad = n rounded up to 128) with an
// identity block in the padding, so every GEMM below runs on whole tiles.
//
//   gram_kernel        K = var * exp(-0.5 |xi-xj|^2 / ls^2) (+ diag_add), identity padding
//   potf2_kernel       64x64 diagonal block: rows in registers (4 threads per row), column broadcast through smem
//   trsm_panel_kernel  panel <- panel * L_jj^-T by substitution (backward stable, like dtrsm), row per thread
//   block_inverse      inverses of all diagonal blocks in parallel (base of W = L^-1 and of the MLL solve)
//   cholesky_blocked   right-looking: potf2 -> panel TRSM -> trailing SYRK (DMMA)
//   tri_inverse        W = L^-1 by recursive doubling: W21 = -W22 (L21 W11), two batched GEMMs per level
//   alpha kernels      alpha = W^T (W (y - mu0))
//   pack_w_kernel      W -> 16 KB tiles in DMMA fragment order for trmm.cu
#include "factor.cuh"

#include <stdlib.h>
#include "gemm.cuh"
#include "rbf.cuh"

namespace bo {

namespace {

constexpr int NB = 64;

// ----------------------------------------------------------------------------------------- gram
// 64x64 tile of pairs per CTA (256 threads, 16 pairs each); the tile's two sets of training rows and the exp table
// are staged in shared memory, every value is written with coalesced 128-byte row segments.
// LOWER = false (bo_gram_f64, the reference's update_k): tiles with bj >= bi; the mirror (j, i) is written too.
// LOWER = true (the factorisations only read the lower triangle): tiles with bj <= bi, one write per value; the
// mirror is kept only in the diagonal tiles so that every 64x64 tile the blocked Cholesky touches is initialised.
// The grid is the triangular list of tiles (no empty CTAs).
constexpr int GT = 64;
template <bool LOWER>
__global__ void __launch_bounds__(256)
    gram_kernel(double* __restrict__ K, long long ldk, long long strideK, const double* __restrict__ x, int ldx,
                int last_eval, int n, int npad_rows, int d, int m, ObjParams hp, double diag_add) {
  __shared__ double exp_tab[64];
  __shared__ double xi[GT][BO_MAX_DIMS + 1];   // rows of the tile (read as broadcasts)
  __shared__ double xjT[BO_MAX_DIMS][GT];      // columns of the tile, transposed (lanes read consecutive words)
  const int tid = threadIdx.x;
  if (tid < 64) exp_tab[tid] = kExp2Tab[tid];
  // triangular tile index -> (big, small) with small <= big
  int big = (int)((sqrt(8.0 * (double)blockIdx.x + 1.0) - 1.0) * 0.5);
  while ((long long)big * (big + 1) / 2 > (long long)blockIdx.x) --big;
  while ((long long)(big + 1) * (big + 2) / 2 <= (long long)blockIdx.x) ++big;
  const int small = blockIdx.x - big * (big + 1) / 2;
  const int bi = LOWER ? big : small, bj = LOWER ? small : big;  // tile row / column
  const int i0 = last_eval + bi * GT, j0 = last_eval + bj * GT;
  for (int e = tid; e < GT * d; e += 256) {
    const int r = e / d, k = e - r * d;
    const int gi = i0 + r, gj = j0 + r;
    xi[r][k] = gi < n ? x[(long long)gi * ldx + k] : 0.0;
    xjT[k][r] = gj < n ? x[(long long)gj * ldx + k] : 0.0;
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  const bool mirror = !LOWER || bi == bj;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int r = ty * 4 + a;
    const int i = i0 + r;
    if (i >= npad_rows) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int cc = tx + 16 * c;
      const int j = j0 + cc;
      if (j >= npad_rows || (bi == bj && (LOWER ? j > i : j < i))) continue;
      if (i >= n || j >= n) {
        // identity padding keeps the padded factor trivially L = W = I
        const double v = (i == j) ? 1.0 : 0.0;
        for (int o = 0; o < m; ++o) {
          K[o * strideK + (long long)i * ldk + j] = v;
          if (mirror) K[o * strideK + (long long)j * ldk + i] = v;
        }
        continue;
      }
      double sq = 0.0;
      for (int k = 0; k < d; ++k) {
        const double diff = xi[r][k] - xjT[k][cc];
        sq = fma(diff, diff, sq);
      }
      for (int o = 0; o < m; ++o) {
        double v = hp.prior_var[o] * rbf_exp(sq * hp.neg_half_inv_ls2[o], exp_tab);
        if (i == j) v += diag_add;
        K[o * strideK + (long long)i * ldk + j] = v;
        if (mirror) K[o * strideK + (long long)j * ldk + i] = v;
      }
    }
  }
}

// ----------------------------------------------------------------------------------------- potf2
// One CTA (256 threads) per matrix in the batch: lower Cholesky factor of the 64x64 diagonal block.
//
// Pivot policy (pol[2b] = floor, pol[2b+1] = negative tolerance): in exact arithmetic every pivot of
// K + jitter*I is >= jitter, so a pivot that rounding pushed below the floor -- but not below -tolerance --
// is clamped to the floor (counted in info[batch + b]); a pivot below -tolerance or NaN means the input is
// genuinely indefinite and is reported in info[b] (numpy LinAlgError at the API).
// Four threads per row: thread (t, q) keeps the columns k = q, q+4, ... of row t in registers (16 values, every
// loop fully unrolled so they are statically indexed).  Per column j: the owner of the pivot publishes
// r = 1/sqrt(pivot), the owners of column j scale their entry and publish it, then every thread updates its
// trailing entries with 16 - j/4 FMAs: two block barriers per column and a quarter of the serial FMA chain of a
// row-per-thread layout.
__global__ void __launch_bounds__(4 * NB) potf2_kernel(double* __restrict__ A, long long lda, long long strideA,
                                                       int* __restrict__ info, int j0, const double* __restrict__ pol,
                                                       int batch) {
  __shared__ double col[NB];
  __shared__ double rinv_sm;
  const int t = threadIdx.x >> 2, q = threadIdx.x & 3;
  const double floor_piv = pol[2 * blockIdx.x], neg_tol = pol[2 * blockIdx.x + 1];
  double* row = A + (long long)blockIdx.x * strideA + (long long)(j0 + t) * lda + j0;
  double s[NB / 4];
#pragma unroll
  for (int i = 0; i < NB / 4; ++i) {
    const int k = q + 4 * i;
    s[i] = (k <= t) ? row[k] : 0.0;
  }
  int bad = 0, nclamp = 0;  // only meaningful in the thread that owns the pivot
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const int ij = j >> 2, qj = j & 3;  // register index / owner lane of column j
    if (t == j && q == qj) {
      double piv = s[ij];
      if (!(piv >= floor_piv)) {  // also catches NaN
        if (piv > -neg_tol) ++nclamp;
        else bad = j0 + j + 1;
        piv = floor_piv;
      }
      // one reciprocal square root per column on the critical path (instead of a square root followed by a
      // division in every row): l_jj = piv * r, l_tj = s_tj * r, each within 1-2 ulp of the divided values
      const double r = rsqrt(piv);
      s[ij] = piv * r;
      rinv_sm = r;
    }
    __syncthreads();
    if (t > j && q == qj) {
      const double l = s[ij] * rinv_sm;
      s[ij] = l;
      col[t] = l;
    }
    __syncthreads();
    if (t > j) {
      const double l = col[t];
      if (q > qj) s[ij] = fma(-l, col[q + 4 * ij], s[ij]);  // the entries of this register right of column j
#pragma unroll
      for (int i = ij + 1; i < NB / 4; ++i) s[i] = fma(-l, col[q + 4 * i], s[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < NB / 4; ++i) {
    const int k = q + 4 * i;
    row[k] = (k <= t) ? s[i] : 0.0;  // strict upper part of the block becomes exact zeros
  }
  // smallest failing pivot index wins (0 = none); earlier blocks were factored by earlier launches
  if (bad != 0) {
    int old = atomicCAS(&info[blockIdx.x], 0, bad);
    while (old != 0 && old > bad) {
      const int prev = atomicCAS(&info[blockIdx.x], old, bad);
      if (prev == old) break;
      old = prev;
    }
  }
  if (nclamp != 0) atomicAdd(&info[batch + blockIdx.x], nclamp);
}

// ----------------------------------------------------------------------------------------- block inverses
// D[jb] = L[jb,jb]^-1 for every 64x64 diagonal block, all blocks in parallel (off the Cholesky critical path).
// Column c of the inverse by forward substitution, 4 threads per column splitting the dot product; the i
// loop is uniform across the warp so the shuffles are convergent.
__global__ void __launch_bounds__(256) block_inverse_kernel(double* __restrict__ D, long long strideD,
                                                            const double* __restrict__ A, long long lda,
                                                            long long strideA) {
  extern __shared__ double sm[];
  double(*S)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(sm);
  double(*X)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(sm + NB * (NB + 1));
  const int tid = threadIdx.x;
  const int jb = blockIdx.x;
  const double* Ab = A + (long long)blockIdx.y * strideA + (long long)jb * NB * (lda + 1);
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e >> 6, c = e & 63;
    S[r][c] = (c <= r) ? Ab[(long long)r * lda + c] : 0.0;
    X[r][c] = 0.0;
  }
  __syncthreads();
  {
    const int c = tid >> 2, q = tid & 3;
    if (q == 0) X[c][c] = 1.0 / S[c][c];
    __syncwarp();
    for (int i = 1; i < NB; ++i) {
      double s = 0.0;
      if (i > c)
        for (int k = c + q; k < i; k += 4) s = fma(S[i][k], X[k][c], s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (q == 0 && i > c) X[i][c] = -s / S[i][i];
      __syncwarp();
    }
  }
  __syncthreads();
  double* Db = D + (long long)blockIdx.y * strideD + (long long)jb * NB * NB;
  for (int e = tid; e < NB * NB; e += 256) Db[e] = X[e >> 6][e & 63];
}

// pol[2b] = pivot floor (jitter of the matrix), pol[2b+1] = sqrt(eps) * max diagonal entry
__global__ void __launch_bounds__(256) chol_policy_kernel(double* __restrict__ pol, const double* __restrict__ A,
                                                          long long lda, long long strideA, int npad,
                                                          const double* __restrict__ jit_dev, double jit_scalar,
                                                          int per_setting) {
  __shared__ double red[8];
  const double* Ab = A + (long long)blockIdx.x * strideA;
  double mx = 0.0;
  for (int i = threadIdx.x; i < npad; i += blockDim.x) mx = fmax(mx, fabs(Ab[(long long)i * (lda + 1)]));
#pragma unroll
  for (int off = 16; off; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) mx = fmax(mx, red[w]);
    const double jit = jit_dev ? jit_dev[blockIdx.x / per_setting] : jit_scalar;
    pol[2 * blockIdx.x] = fmax(jit, 2.220446049250313e-16 * mx);  // never a zero floor
    pol[2 * blockIdx.x + 1] = 1.4901161193847656e-08 * mx;
  }
}

// ----------------------------------------------------------------------------------------- panel TRSM
// P <- P * L_jj^-T by forward substitution (row i of P solves x L_jj^T = p_i), like LAPACK's dtrsm.  An
// explicit inverse of the diagonal block is NOT used here: it is not backward stable when the block is
// ill conditioned (cfg1 reaches cond(K) ~ 1e15) and the trailing update then cancels catastrophically.
// One CTA = 64 rows of the panel; thread r owns row r; L_jj is read from shared memory (broadcast).
// One thread = one row of the panel, held in registers (64 doubles, loops fully unrolled); right-looking: as soon
// as x_c is known the rest of the row is updated with column c of L_jj, read as 16-byte broadcasts from a
// transposed copy in shared memory.  TRSM_ROWS rows per CTA.
constexpr int TRSM_ROWS = 128;
__global__ void __launch_bounds__(TRSM_ROWS) trsm_panel_kernel(double* __restrict__ A, long long lda,
                                                               long long strideA, int j0, int rows) {
  __shared__ __align__(16) double Lt[NB][NB + 2];  // Lt[c][k] = L_jj[k][c]; +2 keeps rows 16-byte aligned and spreads banks
  double* Ab = A + (long long)blockIdx.y * strideA;
  const double* Ljj = Ab + (long long)j0 * lda + j0;
  const int tid = threadIdx.x;
  for (int e = tid; e < NB * NB; e += TRSM_ROWS) {
    const int k = e >> 6, c = e & 63;  // coalesced read of row k
    Lt[c][k] = Ljj[(long long)k * lda + c];
  }
  __syncthreads();
  const int r = blockIdx.x * TRSM_ROWS + tid;
  if (r >= rows) return;
  double* Pg = Ab + (long long)(j0 + NB + r) * lda + j0;
  double p[NB];
  {
    const double2* row = reinterpret_cast<const double2*>(Pg);
#pragma unroll
    for (int k = 0; k < NB; k += 2) {
      const double2 v = row[k >> 1];
      p[k] = v.x;
      p[k + 1] = v.y;
    }
  }
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    const double x = p[c] / Lt[c][c];
    p[c] = x;
    // p[k] -= x * L[k][c] for k > c; pairs (k, k+1) with even k come from one 16-byte load
    if ((c & 1) == 0) p[c + 1] = fma(-x, Lt[c][c + 1], p[c + 1]);
#pragma unroll
    for (int k = (c + 2) & ~1; k < NB; k += 2) {
      const double2 lv = *reinterpret_cast<const double2*>(&Lt[c][k]);
      p[k] = fma(-x, lv.x, p[k]);
      p[k + 1] = fma(-x, lv.y, p[k + 1]);
    }
  }
  {
    double2* row = reinterpret_cast<double2*>(Pg);
#pragma unroll
    for (int k = 0; k < NB; k += 2) row[k >> 1] = make_double2(p[k], p[k + 1]);
  }
}

// ----------------------------------------------------------------------------------------- small helpers
__global__ void copy_diag_blocks_kernel(double* __restrict__ W, long long ldw, long long strideW,
                                        const double* __restrict__ D, long long strideD) {
  // grid (nblk, batch): W[jb*64.., jb*64..] = D[jb]
  const int jb = blockIdx.x;
  const double* Db = D + (long long)blockIdx.y * strideD + (long long)jb * NB * NB;
  double* Wb = W + (long long)blockIdx.y * strideW + (long long)jb * NB * (ldw + 1);
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) Wb[(long long)(e >> 6) * ldw + (e & 63)] = Db[e];
}

// u[i] = sum_{k<=i} W[i][k] * (y[k*ldy + o] - mu0)     (one warp per row)
__global__ void gemv_lower_delta_kernel(double* __restrict__ u, const double* __restrict__ W, long long ldw,
                                        long long strideW, const double* __restrict__ y, int ldy, int n, int npad,
                                        ObjParams hp) {
  const int o = blockIdx.y;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= npad) return;
  const double* Wr = W + o * strideW + (long long)row * ldw;
  double s = 0.0;
  const int kend = min(row + 1, n);
  for (int k = lane; k < kend; k += 32) s = fma(Wr[k], y[(long long)k * ldy + o] - hp.prior_mean[o], s);
#pragma unroll
  for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) u[(long long)o * npad + row] = s;
}

// part[rs][j] = sum over rows i in split rs of W[i][j] * u[i]   (thread per column, coalesced)
__global__ void gemvT_lower_partial_kernel(double* __restrict__ part, const double* __restrict__ W, long long ldw,
                                           long long strideW, const double* __restrict__ u, int npad, int rows_per) {
  const int o = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= npad) return;
  const int i0 = max(blockIdx.y * rows_per, j), i1 = min((blockIdx.y + 1) * rows_per, npad);
  const double* Wo = W + o * strideW;
  const double* uo = u + (long long)o * npad;
  double s = 0.0;
  for (int i = i0; i < i1; ++i) s = fma(Wo[(long long)i * ldw + j], uo[i], s);
  part[((long long)o * gridDim.y + blockIdx.y) * npad + j] = s;
}
__global__ void sum_partials_kernel(double* __restrict__ out, const double* __restrict__ part, int npad, int nsplit) {
  const int o = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= npad) return;
  double s = 0.0;
  for (int r = 0; r < nsplit; ++r) s += part[((long long)o * nsplit + r) * npad + j];
  out[(long long)o * npad + j] = s;
}

// ----------------------------------------------------------------------------------------- pack W
// tile (ib, kt) of W (128 rows x 16 k) -> [wm(2)][i(8)][sp(2)][lane(32)][q(2)]
//   element = W[ib*128 + wm*64 + i*8 + g][kt*16 + (sp*2+q)*4 + t],  lane = g*4 + t
// Rows/columns >= n (the identity padding) are written as zeros, so padded K* rows never contribute.
__global__ void __launch_bounds__(256) pack_w_kernel(double* __restrict__ Wp, long long strideWp,
                                                     const double* __restrict__ W, long long ldw, long long strideW,
                                                     int nb, int n) {
  __shared__ double s[TM][TK + 1];
  const int o = blockIdx.y;
  // decode tile index -> (ib, kt)
  long long tile = blockIdx.x;
  int ib = 0;
  while (wpack_tile_offset(ib + 1) <= tile) ++ib;
  const int kt = (int)(tile - wpack_tile_offset(ib));
  const double* src = W + o * strideW + (long long)ib * TM * ldw + (long long)kt * TK;
  for (int e = threadIdx.x; e < TM * TK / 2; e += 256) {
    const int r = e >> 3, c2 = (e & 7) * 2;
    const double2 v = *reinterpret_cast<const double2*>(src + (long long)r * ldw + c2);
    s[r][c2] = v.x;
    s[r][c2 + 1] = v.y;
  }
  __syncthreads();
  double* dst = Wp + o * strideWp + tile * TILE_DOUBLES;
  for (int e = threadIdx.x; e < TILE_DOUBLES; e += 256) {
    const int q = e & 1, lane = (e >> 1) & 31, sp = (e >> 6) & 1, i = (e >> 7) & 7, wm = e >> 10;
    const int g = lane >> 2, t = lane & 3;
    const int lr = wm * 64 + i * 8 + g, lk = (sp * 2 + q) * 4 + t;
    const bool pad = (ib * TM + lr >= n) || (kt * TK + lk >= n);
    dst[e] = pad ? 0.0 : s[lr][lk];
  }
  (void)nb;
}

// ----------------------------------------------------------------------------------------- incremental append
// (f)2 of SURVEY 8: the reference rebuilds K and its inverse from scratch every iteration
// (bayesian_optimization.py:129-142, last_eval = 0).  When the hyper-parameters did not change, the factor of the
// first n_old points is still valid and the b = n_new - n_old new points only add b rows:
//   K_new = [K11 K12; K21 K22],  L21 = (W11 K12)^T,  L22 L22^T = K22 + jitter I - L21 L21^T,
//   W_new = [W11 0; W21 W22],    W22 = L22^-1,       W21 = -W22 (L21 W11)
// O(b N^2) instead of O(N^3).  The new rows replace identity-padding rows of the padded work matrices, so the
// update is in place as long as n_new stays inside the same 128-row padding (otherwise the caller refits).

// Kc[o][j][i] = K_o(x_i, x_{n_old+j}) (+ jitter on the diagonal entry i = n_old + j), i < n_new
__global__ void __launch_bounds__(256)
    append_cross_kernel(double* __restrict__ Kc, int ldkc, const double* __restrict__ x, int ldx, int n_old, int n_new,
                        int d, ObjParams hp, double jitter) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y, o = blockIdx.z, b = n_new - n_old;
  if (i >= n_new) return;
  double sq = 0.0;
  for (int k = 0; k < d; ++k) {
    const double diff = x[(long long)i * ldx + k] - x[(long long)(n_old + j) * ldx + k];
    sq = fma(diff, diff, sq);
  }
  double v = hp.prior_var[o] * rbf_exp(sq * hp.neg_half_inv_ls2[o], kExp2Tab);
  if (i == n_old + j) v += jitter;
  Kc[((long long)o * b + j) * ldkc + i] = v;
}

// S[o][j][i] = sum_{k<=i} W[i][k] Kc[o][j][k]   (= L21[j][i]), one warp per row i < n_old, all b columns
__global__ void __launch_bounds__(256)
    append_forward_kernel(double* __restrict__ S, const double* __restrict__ Kc, int ldkc, const double* __restrict__ W,
                          long long ldw, long long strideW, int n_old, int b) {
  const int o = blockIdx.y;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n_old) return;
  const double* Wr = W + o * strideW + (long long)row * ldw;
  for (int j = 0; j < b; ++j) {
    const double* kc = Kc + ((long long)o * b + j) * ldkc;
    double s = 0.0;
    for (int k = lane; k <= row; k += 32) s = fma(Wr[k], kc[k], s);
#pragma unroll
    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) S[((long long)o * b + j) * ldkc + row] = s;
  }
}

// One CTA per objective: Schur complement C = K22 + jitter I - L21 L21^T (b x b), its Cholesky factor L22 (same
// pivot policy as potf2_kernel) and W22 = L22^-1.  small[o] = [L22 (b*b) | W22 (b*b)], row-major.
__global__ void __launch_bounds__(256)
    append_schur_kernel(double* __restrict__ small, int* __restrict__ info, const double* __restrict__ S,
                        const double* __restrict__ Kc, int ldkc, int n_old, int b, int m, double jitter) {
  __shared__ double C[BO_MAX_APPEND][BO_MAX_APPEND + 1];
  __shared__ double X[BO_MAX_APPEND][BO_MAX_APPEND + 1];
  const int o = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double* So = S + (long long)o * b * ldkc;
  const double* Ko = Kc + (long long)o * b * ldkc;
  for (int p = warp; p < b * b; p += 8) {
    const int j = p / b, l = p % b;
    if (l > j) continue;
    double s = 0.0;
    for (int i = lane; i < n_old; i += 32) s = fma(So[(long long)j * ldkc + i], So[(long long)l * ldkc + i], s);
#pragma unroll
    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) C[j][l] = Ko[(long long)j * ldkc + n_old + l] - s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double mx = 0.0;
    for (int j = 0; j < b; ++j) mx = fmax(mx, fabs(Ko[(long long)j * ldkc + n_old + j]));
    const double floor_piv = fmax(jitter, 2.220446049250313e-16 * mx), neg_tol = 1.4901161193847656e-08 * mx;
    int bad = 0, nclamp = 0;
    for (int j = 0; j < b; ++j) {
      double piv = C[j][j];
      for (int k = 0; k < j; ++k) piv = fma(-C[j][k], C[j][k], piv);
      if (!(piv >= floor_piv)) {
        if (piv > -neg_tol) ++nclamp;
        else if (bad == 0) bad = n_old + j + 1;
        piv = floor_piv;
      }
      const double ljj = sqrt(piv);
      C[j][j] = ljj;
      for (int i = j + 1; i < b; ++i) {
        double v = C[i][j];
        for (int k = 0; k < j; ++k) v = fma(-C[i][k], C[j][k], v);
        C[i][j] = v / ljj;
      }
    }
    // W22 = L22^-1 by forward substitution, column by column
    for (int c = 0; c < b; ++c) {
      for (int i = 0; i < b; ++i) {
        if (i < c) { X[i][c] = 0.0; continue; }
        double v = (i == c) ? 1.0 : 0.0;
        for (int k = c; k < i; ++k) v = fma(-C[i][k], X[k][c], v);
        X[i][c] = v / C[i][i];
      }
    }
    info[o] = bad;
    info[m + o] = nclamp;
  }
  __syncthreads();
  double* out = small + (long long)o * 2 * b * b;
  for (int p = threadIdx.x; p < b * b; p += blockDim.x) {
    const int j = p / b, l = p % b;
    out[p] = (l <= j) ? C[j][l] : 0.0;
    out[b * b + p] = X[j][l];
  }
}

// part[o][rs][j][c] = sum over rows i of split rs (i >= c) of S[o][j][i] W[i][c]      (T = L21 W11, thread per column)
constexpr int APPEND_SPLIT = 8;
__global__ void __launch_bounds__(128)
    append_backward_partial_kernel(double* __restrict__ part, const double* __restrict__ S, int ldkc,
                                   const double* __restrict__ W, long long ldw, long long strideW, int n_old, int b,
                                   int rows_per) {
  const int o = blockIdx.z, rs = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int i_lo = rs * rows_per, i_hi = min((rs + 1) * rows_per, n_old);
  const double* Wo = W + o * strideW;
  const double* So = S + (long long)o * b * ldkc;
  for (int j0 = 0; j0 < b; j0 += 8) {
    double acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.0;
    if (c < n_old) {
      for (int i = max(i_lo, c); i < i_hi; ++i) {
        const double w = Wo[(long long)i * ldw + c];
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (j0 + q < b) acc[q] = fma(So[(long long)(j0 + q) * ldkc + i], w, acc[q]);  // warp-uniform address
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (j0 + q < b) part[(((long long)o * APPEND_SPLIT + rs) * b + j0 + q) * ldkc + c] = acc[q];
    }
  }
}

// rows n_old + j of L and W:  L = [S | L22 | 0],  W = [-W22 T | W22 | 0]   (T = sum of the partials, fixed order)
__global__ void __launch_bounds__(256)
    append_write_rows_kernel(double* __restrict__ L, double* __restrict__ W, long long ld, long long stride,
                             const double* __restrict__ S, const double* __restrict__ part,
                             const double* __restrict__ small, int ldkc, int n_old, int b, int npad) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y, o = blockIdx.z;
  if (c >= npad) return;
  const double* L22 = small + (long long)o * 2 * b * b;
  const double* W22 = L22 + b * b;
  double lv = 0.0, wv = 0.0;
  if (c < n_old) {
    lv = S[((long long)o * b + j) * ldkc + c];
    double t = 0.0;
    for (int l = 0; l <= j; ++l) {  // W22 is lower triangular
      double tl = 0.0;
      for (int rs = 0; rs < APPEND_SPLIT; ++rs) tl += part[(((long long)o * APPEND_SPLIT + rs) * b + l) * ldkc + c];
      t = fma(W22[j * b + l], tl, t);
    }
    wv = -t;
  } else if (c < n_old + b) {
    const int l = c - n_old;
    lv = (l <= j) ? L22[j * b + l] : 0.0;
    wv = (l <= j) ? W22[j * b + l] : 0.0;
  }
  L[o * stride + (long long)(n_old + j) * ld + c] = lv;
  W[o * stride + (long long)(n_old + j) * ld + c] = wv;
}

}  // namespace

// =========================================================================================== host drivers
size_t append_scratch_doubles(int npad, int m) {
  // Kc + S: 2 * b * npad; partials: APPEND_SPLIT * b * npad; small: 2 b^2   (b <= BO_MAX_APPEND), per objective
  return (size_t)m * ((size_t)(2 + APPEND_SPLIT) * BO_MAX_APPEND * npad + 2 * BO_MAX_APPEND * BO_MAX_APPEND);
}

int append_rows(double* L, double* W, long long ld, long long stride, const double* x, int ldx, int n_old, int n_new,
                int npad, int d, int m, const ObjParams& hp, double jitter, double* scratch, int* info,
                cudaStream_t stream) {
  const int b = n_new - n_old;
  double* Kc = scratch;
  double* S = Kc + (size_t)m * b * npad;
  double* part = S + (size_t)m * b * npad;
  double* small = part + (size_t)m * APPEND_SPLIT * b * npad;
  append_cross_kernel<<<dim3((n_new + 255) / 256, b, m), 256, 0, stream>>>(Kc, npad, x, ldx, n_old, n_new, d, hp, jitter);
  BO_LAUNCH_CHECK("append_cross_kernel");
  append_forward_kernel<<<dim3((n_old + 7) / 8, m), 256, 0, stream>>>(S, Kc, npad, W, ld, stride, n_old, b);
  BO_LAUNCH_CHECK("append_forward_kernel");
  append_schur_kernel<<<m, 256, 0, stream>>>(small, info, S, Kc, npad, n_old, b, m, jitter);
  BO_LAUNCH_CHECK("append_schur_kernel");
  const int rows_per = (n_old + APPEND_SPLIT - 1) / APPEND_SPLIT;
  append_backward_partial_kernel<<<dim3((n_old + 127) / 128, APPEND_SPLIT, m), 128, 0, stream>>>(part, S, npad, W, ld,
                                                                                                stride, n_old, b, rows_per);
  BO_LAUNCH_CHECK("append_backward_partial_kernel");
  append_write_rows_kernel<<<dim3((npad + 255) / 256, b, m), 256, 0, stream>>>(L, W, ld, stride, S, part, small, npad,
                                                                              n_old, b, npad);
  BO_LAUNCH_CHECK("append_write_rows_kernel");
  return BO_OK;
}

int gram(double* K, long long ldk, long long strideK, const double* x, int ldx, int last_eval, int n, int npad_rows,
         int d, int m, const ObjParams& hp, double diag_add, cudaStream_t stream, bool lower_only) {
  const int span = npad_rows - last_eval;
  if (span <= 0) return BO_OK;
  const int nt = (span + GT - 1) / GT;
  const unsigned tiles = (unsigned)((long long)nt * (nt + 1) / 2);
  if (lower_only)
    gram_kernel<true><<<tiles, 256, 0, stream>>>(K, ldk, strideK, x, ldx, last_eval, n, npad_rows, d, m, hp, diag_add);
  else
    gram_kernel<false><<<tiles, 256, 0, stream>>>(K, ldk, strideK, x, ldx, last_eval, n, npad_rows, d, m, hp, diag_add);
  BO_LAUNCH_CHECK("gram_kernel");
  return BO_OK;
}

int cholesky_blocked(double* A, long long lda, long long strideA, int npad, int batch, double* D, long long strideD,
                     int* info, double* pol, const double* jit_dev, double jit_scalar, int per_setting,
                     cudaStream_t stream) {
  chol_policy_kernel<<<batch, 256, 0, stream>>>(pol, A, lda, strideA, npad, jit_dev, jit_scalar, per_setting);
  BO_LAUNCH_CHECK("chol_policy_kernel");
  const int smem = 2 * NB * (NB + 1) * (int)sizeof(double);
  {
    const int rc_attr = ensure_dynamic_smem(block_inverse_kernel, (size_t)smem);
    if (rc_attr) return rc_attr;
  }
  // Two-level blocking: inside an outer panel of NBO columns the 64-column steps update only the rest of that
  // panel (rank-64 updates of a narrow strip); the trailing matrix is updated once per outer panel with K = NBO.
  // A plain right-looking sweep with K = 64 reads and writes the whole trailing matrix 64 times per 4096 columns
  // and is HBM-bound for large batches (cfg5: 512 matrices of 4096^2).
  const int NBO = 4 * NB;  // 128 / 512 / 1024 measured within 6 % of each other on the cfg5 sweep; 256 is the best
  for (int J0 = 0; J0 < npad; J0 += NBO) {
    const int Jend = (J0 + NBO < npad) ? J0 + NBO : npad;
    for (int j0 = J0; j0 < Jend; j0 += NB) {
      potf2_kernel<<<batch, 4 * NB, 0, stream>>>(A, lda, strideA, info, j0, pol, batch);
      BO_LAUNCH_CHECK("potf2_kernel");
      const int r = npad - j0 - NB;
      if (r <= 0) break;
      trsm_panel_kernel<<<dim3((r + TRSM_ROWS - 1) / TRSM_ROWS, batch), TRSM_ROWS, 0, stream>>>(A, lda, strideA, j0, r);
      BO_LAUNCH_CHECK("trsm_panel_kernel");
      const int nin = Jend - (j0 + NB);  // columns of the outer panel still to be factored
      if (nin > 0) {
        const double* P = A + (long long)(j0 + NB) * lda + j0;
        GemmArgs t;  // strip [j0+NB, npad) x [j0+NB, Jend) -= P P^T (tiles above the diagonal skipped)
        t.M = r; t.N = nin; t.K = NB; t.alpha = -1.0; t.beta = 1.0;
        t.A = P; t.lda = lda; t.strideA = strideA;
        t.B = P; t.ldb = lda; t.strideB = strideA;
        t.C = A + (long long)(j0 + NB) * (lda + 1); t.ldc = lda; t.strideC = strideA;
        t.batch = batch; t.lower_only = 1;
        const int rc = gemm(t, 0, 0, stream);
        if (rc) return rc;
      }
    }
    const int ro = npad - Jend;
    if (ro > 0) {
      const double* P = A + (long long)Jend * lda + J0;
      GemmArgs t;  // trailing [Jend, npad)^2 -= P P^T with the whole outer panel, K = Jend - J0
      t.M = ro; t.N = ro; t.K = Jend - J0; t.alpha = -1.0; t.beta = 1.0;
      t.A = P; t.lda = lda; t.strideA = strideA;
      t.B = P; t.ldb = lda; t.strideB = strideA;
      t.C = A + (long long)Jend * (lda + 1); t.ldc = lda; t.strideC = strideA;
      t.batch = batch; t.lower_only = 1;
      const int rc = gemm(t, 0, 0, stream);
      if (rc) return rc;
    }
  }
  block_inverse_kernel<<<dim3(npad / NB, batch), 256, smem, stream>>>(D, strideD, A, lda, strideA);
  BO_LAUNCH_CHECK("block_inverse_kernel");
  return BO_OK;
}

int tri_inverse(double* W, long long ldw, long long strideW, const double* L, long long ldl, long long strideL,
                const double* D, long long strideD, double* T, long long strideT, int npad, int batch,
                cudaStream_t stream) {
  BO_CUDA(cudaMemsetAsync(W, 0, sizeof(double) * (size_t)strideW * batch, stream));
  copy_diag_blocks_kernel<<<dim3(npad / NB, batch), 256, 0, stream>>>(W, ldw, strideW, D, strideD);
  BO_LAUNCH_CHECK("copy_diag_blocks_kernel");
  for (int s = NB; s < npad; s *= 2) {
    const int full_pairs = npad / (2 * s);
    const int rem = npad % (2 * s);
    const int r2_tail = rem > s ? rem - s : 0;
    for (int pass = 0; pass < 2; ++pass) {
      // pass 0: all full pairs, batched per matrix; pass 1: the ragged last pair
      const int npairs = pass == 0 ? full_pairs : (r2_tail > 0 ? 1 : 0);
      if (npairs == 0) continue;
      const int r2 = pass == 0 ? s : r2_tail;
      const long long o0 = pass == 0 ? 0 : (long long)full_pairs * 2 * s;
      for (int b = 0; b < batch; ++b) {
        const double* Lb = L + b * strideL;
        double* Wb = W + b * strideW;
        double* Tb = T + b * strideT;
        GemmArgs a;  // T = L21 * W11
        a.M = r2; a.N = s; a.K = s;
        a.A = Lb + (o0 + s) * ldl + o0; a.lda = ldl; a.strideA = 2LL * s * (ldl + 1);
        a.B = Wb + o0 * (ldw + 1); a.ldb = ldw; a.strideB = 2LL * s * (ldw + 1);
        a.C = Tb; a.ldc = s; a.strideC = (long long)s * s;
        a.batch = npairs; a.k_start_cols = 1;
        int rc = gemm(a, 0, 1, stream);
        if (rc) return rc;
        GemmArgs c;  // W21 = -W22 * T
        c.M = r2; c.N = s; c.K = r2; c.alpha = -1.0;
        c.A = Wb + (o0 + s) * (ldw + 1); c.lda = ldw; c.strideA = 2LL * s * (ldw + 1);
        c.B = Tb; c.ldb = s; c.strideB = (long long)s * s;
        c.C = Wb + (o0 + s) * ldw + o0; c.ldc = ldw; c.strideC = 2LL * s * (ldw + 1);
        c.batch = npairs; c.k_limit_rows = 1;
        rc = gemm(c, 0, 1, stream);
        if (rc) return rc;
      }
    }
  }
  return BO_OK;
}

int compute_alpha(double* alpha, const double* W, long long ldw, long long strideW, const double* y, int ldy, int n,
                  int npad, int m, const ObjParams& hp, double* scratch, cudaStream_t stream) {
  // scratch: m*npad (u) + m*NSPLIT*npad (partials)
  const int NSPLIT = 16;
  double* u = scratch;
  double* part = scratch + (size_t)m * npad;
  gemv_lower_delta_kernel<<<dim3((npad + 7) / 8, m), 256, 0, stream>>>(u, W, ldw, strideW, y, ldy, n, npad, hp);
  BO_LAUNCH_CHECK("gemv_lower_delta_kernel");
  const int rows_per = (npad + NSPLIT - 1) / NSPLIT;
  gemvT_lower_partial_kernel<<<dim3((npad + 127) / 128, NSPLIT, m), 128, 0, stream>>>(part, W, ldw, strideW, u, npad,
                                                                                     rows_per);
  BO_LAUNCH_CHECK("gemvT_lower_partial_kernel");
  sum_partials_kernel<<<dim3((npad + 127) / 128, m), 128, 0, stream>>>(alpha, part, npad, NSPLIT);
  BO_LAUNCH_CHECK("sum_partials_kernel");
  return BO_OK;
}
size_t alpha_scratch_doubles(int npad, int m) { return (size_t)m * npad * 17; }

int pack_w(double* Wp, long long strideWp, const double* W, long long ldw, long long strideW, int npad, int n, int m,
           cudaStream_t stream) {
  const int nb = npad / TM;
  const long long ntiles = wpack_tile_offset(nb);
  pack_w_kernel<<<dim3((unsigned)ntiles, m), 256, 0, stream>>>(Wp, strideWp, W, ldw, strideW, nb, n);
  BO_LAUNCH_CHECK("pack_w_kernel");
  return BO_OK;
}

}  // namespace bo
