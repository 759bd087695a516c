// ozaki.cuh -- host interface of the INT8 tensor-core variance engine (ozaki.cu).
//
// The triangular product V = W K* of the scoring pass, computed by error-free splitting ("Ozaki scheme"):
// every row of W and every K* entry is written as 6 balanced base-256 digits (int8 planes), the 21 digit-pair
// products with s + t <= 5 run on the 5th-generation tensor cores (tcgen05.mma.kind::i8, exact int32
// accumulation in TMEM, one accumulator per digit weight 256^-(s+t)), and the planes are recombined in
// integer / FP64 arithmetic in the epilogue.  See DESIGN.md section 9 for the error analysis.
#pragma once
#include "common.cuh"
#include "score.cuh"

namespace bo {

constexpr int OZ_PLANES = 6;   // digits per operand element
constexpr int OZ_TM = 128;     // W rows per row block  (UMMA M, TMEM lanes)
constexpr int OZ_TN = 64;      // candidates per tile   (UMMA N, TMEM columns per accumulator)
constexpr int OZ_KS = 32;      // k depth of one tcgen05.mma.kind::i8
constexpr int OZ_A_PLANE = OZ_TM * OZ_KS;              // 4096 B
constexpr int OZ_B_PLANE = OZ_TN * OZ_KS;              // 2048 B
constexpr int OZ_A_STAGE = OZ_PLANES * OZ_A_PLANE;     // 24576 B: one k-step of a W row block, all planes
constexpr int OZ_B_STAGE = OZ_PLANES * OZ_B_PLANE;     // 12288 B: one k-step of a K* tile, all planes
constexpr int OZ_MAX_N = 16384;  // int32 accumulators: 6 pairs * 2^14 * npad < 2^31

struct OzPlan {
  int npad = 0, nb = 0, nk_tot = 0;
  int nsplit = 1;           // row blocks of one candidate tile are dealt to nsplit CTAs (keeps K* tiles in L2)
  int chunk_tiles = 0;      // candidate tiles (of 64) per chunk
  long long ld_chunk = 0;   // chunk_tiles * 64
  size_t kq_bytes = 0, part_doubles = 0, mean_doubles = 0;
};
OzPlan make_oz_plan(int n, int m, long long n_cand);
size_t oz_workspace_bytes(const OzPlan& p);

// bytes of the digit planes of W for ONE objective; wscale holds 2 * m * npad doubles (row scales, then the quantisation scales)
size_t oz_wq_bytes(int n);
// wpack (DMMA tile order, from bo_gp_fit_f64) -> digit planes + per-row scale 2^(e_i - 29)
int oz_quantize_w(unsigned char* wq, double* wscale, const double* wpack, int n, int m, cudaStream_t st);

// the scoring pass with the INT8 engine (same contract as score_candidates)
int oz_score_candidates(const ScoreOutputs& out, const void* cand, int cand_kind, int ldc, long long n_cand,
                        const double* x, int ldx, int n, int d, int m, const unsigned char* wq,
                        const double* wscale, const double* alpha, const ObjParams& hp, double min_variance,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream);

// Sampled guard of the INT8 engine: one candidate per window of `stride` (hashed offset) is scored with BOTH engines
// from the same factor; BO_ERR_GUARD if max |var_int8 - var_fp64| / prior_variance exceeds
// max(tol, 10 eps |K + jitter I|_inf |W|_F^2)  (the parity tolerance with a rigorous upper bound of cond).  Synchronising.
size_t oz_guard_workspace_bytes(int n, int m, int d, long long n_cand, long long stride);
int oz_guard(double* worst_host, double* tau_host, const void* cand, int cand_kind, int ldc, long long n_cand,
             long long stride, const double* x, int ldx, int n, int d, int m, const unsigned char* wq,
             const double* wscale, const double* wpack, const double* alpha, const ObjParams& hp, double jitter,
             double min_variance, double tol, void* workspace, size_t workspace_bytes, cudaStream_t stream);

// measured tcgen05.mma.kind::i8 rate of the kernel's own MMA batch on resident operands (TOP/s, mul + add)
int oz_peak_tops(double* tops, double seconds, cudaStream_t st);

// test hooks: the two halves of the pass on caller-provided buffers
int oz_kstar_digits(unsigned char* kq, double* meandot, const void* cand, int cand_kind, int ldc, long long cand0,
                    long long n_cand, int tiles, int chunk_tiles, const double* x, int ldx, int n, int d, int m,
                    const double* alpha, const ObjParams& hp, cudaStream_t st);
int oz_sumsq(double* part, long long ld_chunk, const unsigned char* wq, const double* wscale,
             const unsigned char* kq, int n, int m, int tiles, int chunk_tiles, int nsplit, const ObjParams& hp,
             cudaStream_t st);

}  // namespace bo
