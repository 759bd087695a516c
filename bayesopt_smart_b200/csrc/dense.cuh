// dense.cuh -- drop-ins on the reference's materialised arrays (dense.cu).
#pragma once
#include "common.cuh"

namespace bo {
int kstar_dense(double* ks, long long ld_row, long long ld_obj, const double* x, int ldx, const void* cand,
                int cand_kind, int ldc, long long n_cand, int last_eval, int current_eval, int d, int m,
                const ObjParams& hp, cudaStream_t stream);
size_t dense_workspace_bytes(int n, long long n_cand);
int mean_dense(double* mu, long long ld_mu, const double* ks, long long ld_row, long long ld_obj, const double* kinv,
               int ld_kinv, long long ld_kinv_obj, const double* y, int ldy, const ObjParams& hp, int n,
               long long n_cand, int m, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int variance_dense(double* var, long long ld_var, const double* ks, long long ld_row, long long ld_obj,
                   const double* kinv, int ld_kinv, long long ld_kinv_obj, const ObjParams& hp, double min_variance,
                   int n, long long n_cand, int m, void* workspace, size_t workspace_bytes, cudaStream_t stream);
}  // namespace bo
