/*
 * bo_b200.h -- C ABI of libbo_b200.so, the B200 (sm_100a) implementation of the
 * BayesOpt_smart acquisition hot path.
 *
 * The reference (alebal123bal/BayesOpt_smart) has no FFI of its own: its hot path is a
 * set of Python free functions (bayesopt/numba_kernels.py, bayesopt/acquisition.py,
 * bayesopt/pareto.py) called from bayesopt/bayesian_optimization.py:115-207.  Each entry
 * point below names the reference function(s) (file:line) it replaces; the Python
 * package bayesopt_smart_b200 binds them with ctypes behind the
 * reference's own signatures.
 *
 * Conventions
 *  - plain C: raw pointers, sizes, no torch / C++ types in any signature;
 *  - every pointer named *_dev is DEVICE memory (FP64 unless stated), every pointer
 *    named *_host is host memory read synchronously during the call (tiny arrays:
 *    per-objective hyper-parameters);
 *  - matrices are row-major with an explicit leading dimension ("ld", in elements);
 *  - the caller owns all memory; scratch space is passed in as (workspace_dev,
 *    workspace_bytes) and sized by the matching *_workspace_bytes() query;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all
 *    work is enqueued on it.  Functions documented as "synchronising" wait for the
 *    stream before returning because they report a device-side status;
 *  - return value: 0 = BO_OK, otherwise a BO_ERR_* code; bo_last_error() gives a
 *    thread-local message.  No exception crosses the boundary;
 *  - no hidden global state: re-entrant for distinct streams + workspaces.
 */
#ifndef BO_B200_H
#define BO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BO_ABI_VERSION 1

enum {
  BO_OK = 0,
  BO_ERR_INVALID = 1,   /* bad argument (null pointer, size out of range, misaligned) */
  BO_ERR_CUDA = 2,      /* a CUDA runtime call or kernel launch failed               */
  BO_ERR_NOT_PD = 3,    /* Cholesky met a non-positive pivot (numpy LinAlgError)     */
  BO_ERR_WORKSPACE = 4, /* workspace_bytes too small                                 */
  BO_ERR_GUARD = 5      /* the INT8 engine's sampled cross-check against FP64 failed */
};

/* element type of the candidate matrix (`input_space`): the reference builds an int64
 * Cartesian grid (bayesian_optimization.py:338-340) but its kernels accept floats too */
enum { BO_CAND_F64 = 0, BO_CAND_I64 = 1 };

#define BO_MAX_OBJECTIVES 4 /* m: number of objectives                          */
#define BO_MAX_DIMS 16      /* d: input dimensions                               */
#define BO_MAX_TOPK 1024    /* k of bo_topk_f64                                  */
#define BO_TILE 128         /* training rows are padded to a multiple of this    */
#define BO_MAX_APPEND 32    /* rows bo_gp_append_f64 adds in one call            */

int bo_abi_version(void);
const char* bo_last_error(void);
/* device facts used for grid sizing / roofline reporting */
int bo_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* l2_bytes, size_t* hbm_bytes);

/* ------------------------------------------------------------------ a1: update_k
 * RBF Gram matrix K[o,i,j] = var[o] * exp(-0.5*|x_i-x_j|^2 / ls[o]^2) for rows/cols in
 * [last_eval, current_eval), written symmetric into K_dev (m, ldk, ldk).
 * Replaces update_k, numba_kernels.py:329-367.                                    */
int bo_gram_f64(double* K_dev, int ldk, const double* x_dev, int ldx, int last_eval, int current_eval, int d,
                int m, const double* prior_variance_host, const double* length_scales_host, void* stream);

/* ------------------------------------------------------------------ a2: invert_k
 * Kinv[o] = (K[o][:n,:n] + jitter*I)^-1, dense (m, n, n) with ld = n.  Computed as
 * W^T W with W = chol(K + jitter I)^-1 (blocked Cholesky + triangular inverse on DMMA).
 * Synchronising.  BO_ERR_NOT_PD mirrors numpy.linalg.LinAlgError.
 * Replaces invert_k, numba_kernels.py:370-403 (jitter = KERNEL_JITTER = 1e-6).     */
size_t bo_inverse_workspace_bytes(int n, int m);
int bo_inverse_f64(double* Kinv_dev, const double* K_dev, int ldk, int n, int m, double jitter,
                   void* workspace_dev, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------- a1+a2 fused: the GP factor
 * Builds everything the scoring pass needs from the training set and keeps it on the
 * device:  K = gram(x) + jitter I = L L^T,  W = L^-1 (stored tile-packed in the DMMA
 * fragment order of bo_score_f64),  alpha[o] = K^-1 (y[:,o] - prior_mean[o]).
 *   wpack_dev : m * bo_wpack_doubles(n) doubles     alpha_dev : m * bo_npad(n) doubles
 * Synchronising.  Pivot policy: in exact arithmetic every Cholesky pivot of K + jitter*I is >= jitter;
 * a pivot that rounding pushed below that (but above -sqrt(eps)*max diag) is clamped to `jitter`
 * (count via bo_last_clamped_pivots()); anything more negative, or NaN, returns BO_ERR_NOT_PD.
 * Replaces update_k + invert_k + the "Kinv @ delta_y" half of update_mean
 * (numba_kernels.py:329-403, :477-483) as called at bayesian_optimization.py:129-142. */
int bo_npad(int n);
int bo_last_clamped_pivots(void); /* pivots clamped by the last bo_gp_fit_f64 / bo_inverse_f64 on this thread */
size_t bo_wpack_doubles(int n);
size_t bo_fit_workspace_bytes(int n, int m);
int bo_gp_fit_f64(double* wpack_dev, double* alpha_dev, const double* x_dev, int ldx, const double* y_dev, int ldy,
                  int n, int d, int m, const double* prior_mean_host, const double* prior_variance_host,
                  const double* length_scales_host, double jitter, void* workspace_dev, size_t workspace_bytes,
                  void* stream);

/* ------------------------------------------- incremental factor update
 * The reference rebuilds K and its inverse from scratch every iteration (update_k / invert_k are called with
 * last_eval = 0, bayesian_optimization.py:129-142).  When the hyper-parameters are unchanged the factor of the
 * first n_old points stays valid; this call adds the rows of points [n_old, n_new):
 *   L21 = (W11 K12)^T,  L22 L22^T = K22 + jitter I - L21 L21^T,  W22 = L22^-1,  W21 = -W22 L21 W11,
 * then recomputes alpha and re-packs W: O((n_new - n_old) n^2) instead of O(n^3).
 * `workspace_dev` MUST be the workspace of the preceding bo_gp_fit_f64 / bo_gp_append_f64 call for the same
 * training prefix, untouched since (it holds the dense L and W); x_dev / y_dev hold all n_new rows; the
 * hyper-parameters and jitter must be the ones of that fit.  Requirements: n_new - n_old <= BO_MAX_APPEND and
 * bo_npad(n_new) == bo_npad(n_old) (otherwise BO_ERR_INVALID: refit).  Same pivot policy and error codes as
 * bo_gp_fit_f64.  Synchronising.                                                                      */
int bo_gp_append_f64(double* wpack_dev, double* alpha_dev, const double* x_dev, int ldx, const double* y_dev, int ldy,
                     int n_old, int n_new, int d, int m, const double* prior_mean_host,
                     const double* prior_variance_host, const double* length_scales_host, double jitter,
                     void* workspace_dev, size_t workspace_bytes, void* stream);

/* ------------------------------------------- a3..a8 fused: predict + UCB + sum-UCB
 * For candidates c in [0, n_cand):  k* (RBF cross kernel, never materialised in API
 * memory), mu = mu0 + k*.alpha, var = max(var0 - |W k*|^2, min_variance), standardise,
 * UCB, acq = sum_o ucb[o].  Output arrays are (m, ld_out) / (ld_out,) and any of them
 * may be NULL (not written).  cand_kind selects f64 / i64 candidates (row-major, ldc).
 * Asynchronous on `stream`.
 * Replaces update_k_star, update_mean, update_variance, standardize_objectives
 * (numba_kernels.py:406-570), update_ucb, update_hypervolume_improvement
 * (acquisition.py:55-108) as called at bayesian_optimization.py:145-199.           */
size_t bo_score_workspace_bytes(int n, int m, long long n_cand);
int bo_score_f64(double* mu_dev, double* var_dev, double* std_mu_dev, double* std_var_dev, double* ucb_dev,
                 double* acq_dev, long long ld_out, const void* cand_dev, int cand_kind, int ldc, long long n_cand,
                 const double* x_dev, int ldx, int n, int d, int m, const double* wpack_dev,
                 const double* alpha_dev, const double* prior_mean_host, const double* prior_variance_host,
                 const double* length_scales_host, const double* betas_host, double min_variance,
                 void* workspace_dev, size_t workspace_bytes, void* stream);

/* ------------------------------------- a3..a8 fused, INT8 tensor-core variance engine
 * Same contract and outputs as bo_score_f64; the triangular product |W k*|^2 is computed by error-free
 * splitting on tcgen05.mma.kind::i8: W rows (per-row power-of-two scale) and K* entries are written as six
 * balanced base-256 digits, the 21 digit-pair products with s + t <= 5 are accumulated exactly in int32 (TMEM)
 * and recombined in int64 / FP64.  Truncation error of var / prior_variance: ~3e-12 at cond(K) ~ 1e7
 * (DESIGN.md section 9), i.e. far inside the 1e-9 parity tolerance but not bit-identical to bo_score_f64.
 *   bo_i8_quantize_w : wpack (from bo_gp_fit_f64) -> digit planes wq (m * bo_i8_wq_bytes(n) bytes) and
 *                      row scales wscale (bo_i8_wscale_doubles(n, m) doubles).  Asynchronous on `stream`.
 *   n <= 16384 (int32 accumulator bound).
 * Replaces update_variance's contraction (numba_kernels.py:510-535) inside the same fused pass.     */
size_t bo_i8_wq_bytes(int n);
size_t bo_i8_wscale_doubles(int n, int m);
int bo_i8_quantize_w(uint8_t* wq_dev, double* wscale_dev, const double* wpack_dev, int n, int m, void* stream);
size_t bo_score_i8_workspace_bytes(int n, int m, long long n_cand);
int bo_score_i8(double* mu_dev, double* var_dev, double* std_mu_dev, double* std_var_dev, double* ucb_dev,
                double* acq_dev, long long ld_out, const void* cand_dev, int cand_kind, int ldc, long long n_cand,
                const double* x_dev, int ldx, int n, int d, int m, const uint8_t* wq_dev, const double* wscale_dev,
                const double* alpha_dev, const double* prior_mean_host, const double* prior_variance_host,
                const double* length_scales_host, const double* betas_host, double min_variance,
                void* workspace_dev, size_t workspace_bytes, void* stream);
/* Sampled guard of the INT8 engine.  The engine's a-priori worst-case error (every rounding and every dropped
 * digit product conspiring) is ~2e-7 of the prior variance at cfg2 -- above the 1e-9 parity
 * target -- while the measured error is ~1e-11 because balanced digits make those terms zero-mean.  The guard turns
 * that statistical statement into a checked one: one candidate per window of `stride` candidates (hashed offset
 * inside the window) is scored with BOTH engines from the same factor, and the call returns BO_ERR_GUARD -- no
 * silent fallback -- if max |var_int8 - var_fp64| / prior_variance exceeds tau = max(tol, 10 eps cond_upper), the
 * parity tolerance of SURVEY 8(c) with the rigorous upper bound cond(K + jitter I) <= |K + jitter I|_inf *
 * trace((K + jitter I)^-1) = (largest row sum of the Gram matrix) * |W|_F^2 evaluated on the device (for well-conditioned fits tau is
 * `tol`, i.e. 1e-9; it widens only where the FP64 result itself is no better).  The sampled candidates' INT8
 * results are bit for bit those of the main pass (a candidate's numbers do not depend on chunking).  If a fraction
 * p of all candidates violated `tol`, a sample of S = ceil(n_cand / stride) misses them with probability
 * (1 - p)^S (stride 4096, 10^6 candidates: S = 245, p = 2 %  ->  0.7 %); errors of this engine come from the
 * quantisation of W and K*, which every candidate shares, so a real failure shows up in essentially every sample.
 * wpack_dev is the FP64 factor of the same fit (jitter: the one it was built with).  *worst_host receives the
 * largest sampled difference, *tau_host the tolerance it was held to.  Synchronising.                      */
size_t bo_i8_guard_workspace_bytes(int n, int m, int d, long long n_cand, long long stride);
int bo_i8_guard_f64(double* worst_host, double* tau_host, const void* cand_dev, int cand_kind, int ldc,
                    long long n_cand, long long stride, const double* x_dev, int ldx, int n, int d, int m,
                    const uint8_t* wq_dev, const double* wscale_dev, const double* wpack_dev,
                    const double* alpha_dev, const double* prior_mean_host, const double* prior_variance_host,
                    const double* length_scales_host, double jitter, double min_variance, double tol,
                    void* workspace_dev, size_t workspace_bytes, void* stream);
/* roofline denominator of the INT8 engine: the rate (TOP/s, multiply + add) at which this GPU executes the
 * kernel's own MMA batch (tcgen05.mma.kind::i8 128x64x32, A from TMEM) on resident operands for about `seconds`
 * seconds, one CTA per SM.  Synchronising.                                                              */
int bo_i8_peak_tops(double* tops_host, double seconds, void* stream);
/* test hooks: the two halves of the INT8 pass on caller-owned buffers (tiles = ceil(n_cand / 64) rounded up to a multiple of 4).
 *   kq_dev  : m * tiles * bo_npad(n) * 384 bytes of K* digit planes;  meandot_dev: (m, tiles*64) k*.alpha
 *   q_dev   : (m, nsplit, tiles*64) partial sums of |W k*|^2 (sum over nsplit = the full quadratic form)  */
int bo_i8_kstar_digits(uint8_t* kq_dev, double* meandot_dev, const void* cand_dev, int cand_kind, int ldc,
                       long long n_cand, const double* x_dev, int ldx, int n, int d, int m,
                       const double* alpha_dev, const double* prior_variance_host,
                       const double* length_scales_host, void* stream);
int bo_i8_sumsq(double* q_dev, const uint8_t* wq_dev, const double* wscale_dev, const uint8_t* kq_dev, int n, int m,
                long long n_cand, int nsplit, const double* prior_variance_host, void* stream);

/* ----------------------------------------------------- a6..a8 stand-alone (HBM bound)
 * standardize_objectives + update_ucb + update_hypervolume_improvement on existing
 * (m, ld) mu / var arrays.  Outputs may be NULL.  numba_kernels.py:538-570,
 * acquisition.py:55-108.                                                          */
int bo_acquisition_f64(double* std_mu_dev, double* std_var_dev, double* ucb_dev, double* acq_dev,
                       const double* mu_dev, const double* var_dev, long long ld, long long n_cand, int m,
                       const double* prior_mean_host, const double* prior_variance_host, const double* betas_host,
                       void* stream);

/* ------------------------------------------------------------ a9: select_next_batch
 * Top-k of acq (value descending, index ascending on ties, NaN last).  Writes k
 * (value, index + index_base) pairs sorted best-first.  k <= BO_MAX_TOPK.  For n_cand >= 2^18 and
 * k <= 256 the scan is: exact top-k of a hashed 1/stride sample -> its k-th element is a lower bound of
 * the true k-th best -> one streaming filter pass -> exact top-k of the survivors.  The survivor count stays
 * on the device: the survivor pass and the plain multi-level scan (the fallback when the survivors overflow or
 * are too few) are both enqueued and gated by it, so the call is asynchronous on `stream`.
 * bo_match_rows_f64 flags, for each listed candidate row, whether it equals (==, all
 * d coordinates) some evaluated row x[0:n): the exclusion test of acquisition.py:139.
 * Replaces the argsort + walk of select_next_batch, acquisition.py:116-144.        */
size_t bo_topk_workspace_bytes(long long n_cand, int k);
int bo_topk_f64(double* out_val_dev, long long* out_idx_dev, const double* acq_dev, long long n_cand, int k,
                long long index_base, void* workspace_dev, size_t workspace_bytes, void* stream);
int bo_match_rows_f64(uint8_t* out_flag_dev, const long long* idx_dev, int n_idx, long long index_base,
                      const void* cand_dev, int cand_kind, int ldc, const double* x_dev, int ldx, int n, int d,
                      void* stream);
/* exhaustive form of the same exclusion test over the WHOLE candidate set: out[i] = NaN (ranked last by
 * bo_topk_f64) where candidate i equals some evaluated row, acq[i] elsewhere.  O(n_cand * n) compares; the
 * selection falls back to it only when every listed top-BO_MAX_TOPK row turned out to be an evaluated point. */
int bo_mask_evaluated_f64(double* out_dev, const double* acq_dev, const void* cand_dev, int cand_kind, int ldc,
                          long long n_cand, const double* x_dev, int ldx, int n, int d, void* stream);
/* merge of several sorted-or-not (value, index) lists into the global top-k with the
 * same comparator (used after the all-gather of per-rank top-k lists).             */
int bo_topk_merge_f64(double* out_val_dev, long long* out_idx_dev, const double* val_dev, const long long* idx_dev,
                      int n_pairs, int k, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ----------------------------------------------------------- a10: is_pareto_efficient
 * mask[i] = 1 unless some j has all(y_j >= y_i) and any(y_j > y_i)  (maximisation,
 * duplicates kept, NaN rows kept).  y is (n, ldy) row-major, m <= BO_MAX_OBJECTIVES.
 * `against` variant: rows of y are tested against a separate set z (used for the
 * sample-front prefilter and for the cross-rank union pass).
 * Replaces is_pareto_efficient, pareto.py:12-45.                                   */
int bo_pareto_mask_f64(uint8_t* mask_dev, const double* y_dev, long long ldy, long long n, int m, void* stream);
int bo_pareto_mask_against_f64(uint8_t* mask_dev, const double* y_dev, long long ldy, long long n,
                               const double* z_dev, long long ldz, long long nz, int m, void* stream);
/* The same mask for LARGE n (BASELINE config 4 filters the 8 M UCB vectors) entirely on the device: four rounds of
 * [strided sample -> its exact front, strongest points first -> drop every row it dominates -> stream compaction],
 * then the plain n_s x n_s test among the survivors and a scatter back to the input order.  Exact: a row dominated
 * by a sample-front member is dominated, efficient rows always survive.  All counts stay on the device (no host
 * synchronisation, no sort/compaction outside the library); asynchronous on `stream`.                     */
size_t bo_pareto_workspace_bytes(long long n, int m);
int bo_pareto_mask_filtered_f64(uint8_t* mask_dev, const double* y_dev, long long ldy, long long n, int m,
                                void* workspace_dev, size_t workspace_bytes, void* stream);

/* ----------------------------------------------------------------- a11: compute_mll
 * Batched log marginal likelihood.  For setting s in [0, S): every objective o uses
 * length scale ls[s*m + o] and jitter jit[s]; the Gram matrix is the pure correlation
 * matrix (prior_variance cancels, numba_kernels.py:195-197).  out[s] = sum_o mll_o.
 * Non-PD settings give NaN (no error).  Synchronising.
 * Replaces compute_mll, numba_kernels.py:152-235 (jit = CHOLESKY_JITTER = 1e-8).    */
size_t bo_mll_workspace_bytes(int n, int m, int n_settings);
int bo_mll_batched_f64(double* out_dev, const double* x_dev, int ldx, const double* y_dev, int ldy, int n, int d,
                       int m, const double* prior_mean_host, const double* length_scales_host,
                       const double* jitter_host, int n_settings, void* workspace_dev, size_t workspace_bytes,
                       void* stream);

/* ------------------------------------------------ opt-in: exact hypervolume improvement
 * hvi[i] = HV(front U {u_i}) - HV(front), u_i = (ucb[0][i], .., ucb[m-1][i]), m = 2 or 3.
 * `front_dev` is (n_front, m) row-major, sorted by objective 0 DESCENDING (points below `ref` are clipped to
 * it; dominated points are harmless); fronts of more than 1024 points are read from global memory instead of
 * shared memory (no cap).  Second pass over an existing ucb array: the fused forms are below.  No reference
 * counterpart (the reference's "HVI" is sum-UCB, acquisition.py:104-108).           */
int bo_hvi_f64(double* hvi_dev, const double* ucb_dev, long long ld, long long n_cand, int m,
               const double* front_dev, int n_front, const double* ref_host, void* stream);

/* ------------------------------- opt-in: UCB + exact HVI as ONE fused per-candidate pass
 * BASELINE.json's north star: "UCB and 2-objective and 3-objective HVI become one fused per-candidate kernel
 * against a sorted Pareto front held in shared memory".  The front is PREPARED on the device (no host sort):
 *   bo_hvi_prepare_f64: raw points (n_points, ld) row-major (e.g. the standardised observed objectives) ->
 *     clipped to `ref`, dominated points removed (the warp-ballot dominance kernel of a10), sorted by objective 0
 *     descending, plus the per-front tables the evaluation needs (m = 2: staircase prefix areas, which make the
 *     2-objective HVI two binary searches + O(1) for any front size; m = 3: objective-2 levels and ranks, swept
 *     from shared memory up to 1024 points and from global memory beyond).  prepared_dev holds
 *     bo_hvi_front_doubles(n_points, m) doubles, n_front_dev one int.  Asynchronous, no host synchronisation.
 *   bo_acquisition_hvi_f64: stand-alone fused pass standardise + UCB + HVI over existing (m, ld) mu / var arrays
 *     (replaces standardize_objectives + update_ucb + update_hypervolume_improvement, numba_kernels.py:538-570,
 *     acquisition.py:55-108, with the exact HVI in place of the sum).  Outputs may be NULL.
 *   bo_score_hvi_f64: the whole scoring pass (same contract as bo_score_f64 / bo_score_i8) whose epilogue writes
 *     acq = HVI(ucb vector) -- the UCB array never makes an extra trip through HBM.  engine 0: factor_dev = wpack
 *     (FP64 DMMA engine), wscale_dev NULL; engine 1: factor_dev = the int8 digit planes wq, wscale_dev their
 *     scales.  workspace sized by bo_score_workspace_bytes / bo_score_i8_workspace_bytes.
 * m = 2 or 3.  `n_points` passed to the evaluation calls must be the value given to bo_hvi_prepare_f64.
 * No reference counterpart (its "HVI" is sum-UCB); specification: oracle/gp_oracle.py exact_hvi.           */
size_t bo_hvi_front_doubles(int n_points, int m);
size_t bo_hvi_workspace_bytes(int n_points, int m);
int bo_hvi_prepare_f64(double* prepared_dev, int* n_front_dev, const double* points_dev, long long ld, int n_points,
                       int m, const double* ref_host, void* workspace_dev, size_t workspace_bytes, void* stream);
int bo_acquisition_hvi_f64(double* std_mu_dev, double* std_var_dev, double* ucb_dev, double* hvi_dev,
                           const double* mu_dev, const double* var_dev, long long ld, long long n_cand, int m,
                           const double* prior_mean_host, const double* prior_variance_host,
                           const double* betas_host, const double* prepared_dev, const int* n_front_dev,
                           int n_points, const double* ref_host, void* stream);
int bo_score_hvi_f64(int engine, double* mu_dev, double* var_dev, double* std_mu_dev, double* std_var_dev,
                     double* ucb_dev, double* acq_dev, long long ld_out, const void* cand_dev, int cand_kind, int ldc,
                     long long n_cand, const double* x_dev, int ldx, int n, int d, int m, const void* factor_dev,
                     const double* wscale_dev, const double* alpha_dev, const double* prior_mean_host,
                     const double* prior_variance_host, const double* length_scales_host, const double* betas_host,
                     double min_variance, const double* prepared_dev, const int* n_front_dev, int n_points,
                     const double* ref_host, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------ candidate set on the device
 * Rows [row0, row0 + rows) of the integer Cartesian grid prod_k [lo_k, hi_k) in C order -- what the reference
 * builds on the host with np.meshgrid(*[np.arange(lo, hi)], indexing="ij") raveled
 * (bayesian_optimization.py:338-340).  out is (rows, ld) int64, d <= BO_MAX_DIMS.  A rank of a sharded run
 * generates its own slice instead of receiving it over PCIe.                                          */
int bo_grid_i64(long long* out_dev, long long ld, const long long* lo_host, const long long* hi_host, int d,
                long long row0, long long rows, void* stream);

/* ---------------------------------------------- function-level drop-ins on dense arrays
 * The reference's free functions exchange a materialised k_star (m, T, M).  These keep
 * that contract for callers that use the functions one by one.
 * bo_kstar_dense_f64   : update_k_star, numba_kernels.py:406-442
 * bo_mean_dense_f64    : update_mean,   numba_kernels.py:450-488
 * bo_variance_dense_f64: update_variance, numba_kernels.py:491-535                  */
int bo_kstar_dense_f64(double* kstar_dev, long long ld_row, long long ld_obj, const double* x_dev, int ldx,
                       const void* cand_dev, int cand_kind, int ldc, long long n_cand, int last_eval,
                       int current_eval, int d, int m, const double* prior_variance_host,
                       const double* length_scales_host, void* stream);
size_t bo_dense_workspace_bytes(int n, long long n_cand);
int bo_mean_dense_f64(double* mu_dev, long long ld_mu, const double* kstar_dev, long long ld_row, long long ld_obj,
                      const double* kinv_dev, int ld_kinv, long long ld_kinv_obj, const double* y_dev, int ldy,
                      const double* prior_mean_host, int n, long long n_cand, int m, void* workspace_dev,
                      size_t workspace_bytes, void* stream);
int bo_variance_dense_f64(double* var_dev, long long ld_var, const double* kstar_dev, long long ld_row,
                          long long ld_obj, const double* kinv_dev, int ld_kinv, long long ld_kinv_obj,
                          const double* prior_variance_host, double min_variance, int n, long long n_cand, int m,
                          void* workspace_dev, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ measurement aid
 * Plain FP64 GEMM C = A * B^T (row-major, all n x n) on the library's own DMMA kernel;
 * used by bench.py / tests to report the kernel's throughput next to the roofline.   */
int bo_dgemm_nt_f64(double* C_dev, const double* A_dev, const double* B_dev, int n, void* stream);

/* Kernel launches issued by this library since the last reset (process-wide counter; bench.py's
 * "gpu_launches").  reset != 0 zeroes it after reading.                                           */
long long bo_launch_count(int reset);

/* Live timing of the dominant kernel (trmm_sumsq_kernel) with CUDA events recorded on the launching
 * stream inside bo_score_f64.  bo_profile_enable(1) starts collecting, bo_profile_read() synchronises
 * the recorded events and returns the summed duration (ms), launch count and algorithmic FLOPs
 * (m * N^2 per candidate) since the last read, then clears them.                      */
int bo_profile_enable(int on);
int bo_profile_read(double* total_ms, long long* launches, double* flops);
/* The same hook per kernel class: every launch of the scoring pass is bracketed by events while profiling is on.
 * `work` is the algorithmic work the launches stood for -- FLOPs for BO_PROF_CONTRACTION (m * N^2 per candidate),
 * bytes for the others (K* staging written, partial sums read + API arrays written, 8 B per scanned score).
 * kernel = -1 sums (and clears) every class.                                                                */
enum {
  BO_PROF_CONTRACTION = 0, /* trmm_sumsq_kernel / oz_sumsq_kernel                               */
  BO_PROF_KSTAR = 1,       /* kstar_pack_kernel / oz_kstar_digits_kernel                        */
  BO_PROF_FINALIZE = 2,    /* finalize_kernel (variance, standardise, UCB, acquisition)         */
  BO_PROF_TOPK = 3,        /* the launches of one bo_topk_f64 call                              */
  BO_PROF_FIT = 4,         /* the launches of one bo_gp_fit_f64 / bo_gp_append_f64 call         */
  BO_PROF_KINDS = 5
};
int bo_profile_read_kernel(int kernel, double* total_ms, long long* launches, double* work);

#ifdef __cplusplus
}
#endif
#endif /* BO_B200_H */
