// score.cuh -- host interface of the per-candidate scoring pass (score.cu).
#pragma once
#include "common.cuh"
#include "hvi.cuh"

namespace bo {

struct ScorePlan {
  int npad = 0, nb = 0;
  int chunk_tiles = 0;      // candidate tiles (of 128) per chunk
  long long ld_chunk = 0;   // chunk_tiles * 128
  int nbuf = 1;             // K* staging buffers (2 = generate chunk i+1 while TRMM(i) runs)
  size_t kp_doubles = 0, part_doubles = 0, mean_doubles = 0;
};
ScorePlan make_score_plan(int n, int m, long long n_cand);
size_t score_workspace_bytes(const ScorePlan& p);

struct ScoreOutputs {
  double *mu = nullptr, *var = nullptr, *std_mu = nullptr, *std_var = nullptr, *ucb = nullptr, *acq = nullptr;
  long long ld = 0;
  HviSpec hvi;  // hvi.prepared != nullptr: acq = exact hypervolume improvement of the UCB vector (opt-in mode)
};

int score_candidates(const ScoreOutputs& out, const void* cand, int cand_kind, int ldc, long long n_cand,
                     const double* x, int ldx, int n, int d, int m, const double* wpack, const double* alpha,
                     const ObjParams& hp, double min_variance, void* workspace, size_t workspace_bytes,
                     cudaStream_t stream);

// var = max(var0 - sum_b part[o][b][c], min_var), mu, standardise, UCB, acq for one chunk (shared by both engines)
int finalize_chunk(const ScoreOutputs& out, long long cand0, long long n_cand, const double* part,
                   const double* meandot, long long ld_chunk, int chunk_cands, int nb, int m, const ObjParams& hp,
                   double min_variance, cudaStream_t stream);

int acquisition_only(double* smu, double* svar, double* ucb, double* acq, const double* mu, const double* var,
                     long long ld, long long n_cand, int m, const ObjParams& hp, cudaStream_t stream);

}  // namespace bo
