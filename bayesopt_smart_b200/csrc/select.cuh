// select.cuh -- host interface of select.cu (top-k, exclusion test, Pareto mask, exact HVI).
#pragma once
#include "common.cuh"

namespace bo {

size_t topk_workspace_bytes(long long n_cand, int k);
// idx == nullptr: implicit indices index_base + position; entries with idx < 0 are ignored.
int topk_levels(double* out_val, long long* out_idx, const double* val, const long long* idx, long long n_in, int k,
                long long index_base, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int match_rows(uint8_t* flag, const long long* idx, int n_idx, long long index_base, const void* cand, int cand_kind,
               int ldc, const double* x, int ldx, int n, int d, cudaStream_t stream);
// out[i] = NaN where candidate i equals an evaluated row, acq[i] elsewhere (exhaustive exclusion test)
int mask_evaluated(double* out, const double* acq, const void* cand, int cand_kind, int ldc, long long n_cand,
                   const double* x, int ldx, int n, int d, cudaStream_t stream);
int pareto_mask(uint8_t* mask, const double* y, long long ldy, long long n, const double* z, long long ldz,
                long long nz, int m, cudaStream_t stream);
// the same mask for large n without host round trips: four rounds of thinning against the exact front of a strided
// sample, stream compaction on the device, then the plain test among the survivors (exact; see select.cu)
size_t pareto_filtered_workspace_bytes(long long n, int m);
int pareto_mask_filtered(uint8_t* mask, const double* y, long long ldy, long long n, int m, void* workspace,
                         size_t workspace_bytes, cudaStream_t stream);
int hvi(double* out, const double* ucb, long long ld, long long n_cand, int m, const double* front, int n_front,
        const double* ref, cudaStream_t stream);

}  // namespace bo
