// mll.cuh -- batched log marginal likelihood (mll.cu).
#pragma once
#include "common.cuh"

namespace bo {
size_t mll_workspace_bytes(int n, int m, int n_settings);
int mll_batched(double* out, const double* x, int ldx, const double* y, int ldy, int n, int d, int m,
                const double* prior_mean, const double* length_scales, const double* jitter, int n_settings,
                void* workspace, size_t workspace_bytes, cudaStream_t stream);
}  // namespace bo
