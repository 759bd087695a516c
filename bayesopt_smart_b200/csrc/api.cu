// api.cu -- the exported C ABI (include/bo_b200.h).  Argument checking, workspace carving, error state.
#include <stdarg.h>

#include <atomic>
#include <map>
#include <mutex>
#include <vector>

#include "dense.cuh"
#include "factor.cuh"
#include "gemm.cuh"
#include "hvi.cuh"
#include "mll.cuh"
#include "ozaki.cuh"
#include "score.cuh"
#include "select.cuh"

namespace bo {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return BO_ERR_CUDA;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

namespace {
struct ProfRecord {
  int kernel;
  cudaEvent_t start, stop;
  double work;
};
struct ProfState {
  std::mutex mu;
  bool on = false;
  std::vector<cudaEvent_t> pool;  // reusable events
  std::vector<ProfRecord> live;   // recorded (kernel id, start, stop, work) tuples
  cudaEvent_t pending = nullptr;
  int pending_kernel = 0;
  cudaEvent_t get() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
};
ProfState g_prof;
}  // namespace

bool profile_enabled() { return g_prof.on; }
void profile_begin(cudaStream_t st, int kernel) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  if (!g_prof.on) return;
  g_prof.pending = g_prof.get();
  g_prof.pending_kernel = kernel;
  cudaEventRecord(g_prof.pending, st);
}
void profile_end(cudaStream_t st, double work) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  if (!g_prof.on || !g_prof.pending) return;
  cudaEvent_t stop = g_prof.get();
  cudaEventRecord(stop, st);
  g_prof.live.push_back(ProfRecord{g_prof.pending_kernel, g_prof.pending, stop, work});
  g_prof.pending = nullptr;
}

constexpr int kMaxDevices = 64;

int device_sm_count() {
  static std::atomic<int> sms[kMaxDevices];  // zero-initialised; one slot per device ordinal
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  int v = sms[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

int ensure_dynamic_smem_impl(const void* fn, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> done;  // (function, device) -> largest size set so far
  int dev = 0;
  BO_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  size_t& have = done[std::make_pair(fn, dev)];
  if (bytes > have) {
    BO_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
  }
  return BO_OK;
}

static int make_params(ObjParams* hp, int m, const double* mean, const double* var, const double* ls,
                       const double* beta) {
  memset(hp, 0, sizeof(*hp));
  if (m < 1 || m > BO_MAX_OBJECTIVES) {
    set_error("invalid argument: 1 <= m <= %d required (got %d)", BO_MAX_OBJECTIVES, m);
    return BO_ERR_INVALID;
  }
  for (int o = 0; o < m; ++o) {
    hp->prior_mean[o] = mean ? mean[o] : 0.0;
    hp->prior_var[o] = var ? var[o] : 1.0;
    hp->neg_half_inv_ls2[o] = ls ? -0.5 / (ls[o] * ls[o]) : 0.0;
    hp->beta[o] = beta ? beta[o] : 0.0;
  }
  return BO_OK;
}

namespace {

// A[o] = K[o][:n,:n] + jitter I, identity padding up to npad
__global__ void pad_copy_kernel(double* __restrict__ A, int npad, const double* __restrict__ K, int ldk, int n,
                                double jitter) {
  const int o = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= npad) return;
  double v;
  if (i < n && j < n) {
    v = K[((long long)o * ldk + i) * ldk + j];
    if (i == j) v += jitter;
  } else {
    v = (i == j) ? 1.0 : 0.0;
  }
  A[((long long)o * npad + i) * npad + j] = v;
}

struct GridSpec {
  long long lo[BO_MAX_DIMS], extent[BO_MAX_DIMS];
};
// row i of the C-ordered grid: mixed-radix digits of i, last dimension fastest
__global__ void grid_kernel(long long* __restrict__ out, long long ld, GridSpec g, int d, long long row0,
                            long long rows) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  long long i = row0 + r;
  for (int k = d - 1; k >= 0; --k) {
    const long long q = i / g.extent[k];
    out[r * ld + k] = g.lo[k] + (i - q * g.extent[k]);
    i = q;
  }
}

struct FitBuffers {
  double *A, *W, *T, *D, *scratch, *pol, *append;
  int* info;
};

thread_local int g_last_clamped = 0;

size_t carve_fit(FitBuffers* fb, void* ws, int npad, int m) {
  unsigned char* p = static_cast<unsigned char*>(ws);
  size_t off = 0;
  const size_t mat = align256((size_t)m * npad * npad * sizeof(double));
  if (fb) fb->A = reinterpret_cast<double*>(p + off);
  off += mat;
  if (fb) fb->W = reinterpret_cast<double*>(p + off);
  off += mat;
  if (fb) fb->T = reinterpret_cast<double*>(p + off);
  off += align256((size_t)m * npad * npad / 2 * sizeof(double));
  if (fb) fb->D = reinterpret_cast<double*>(p + off);
  off += align256((size_t)m * npad * 64 * sizeof(double));
  if (fb) fb->scratch = reinterpret_cast<double*>(p + off);
  off += align256(alpha_scratch_doubles(npad, m) * sizeof(double));
  if (fb) fb->info = reinterpret_cast<int*>(p + off);
  off += align256((size_t)2 * m * sizeof(int));
  if (fb) fb->pol = reinterpret_cast<double*>(p + off);
  off += align256((size_t)2 * m * sizeof(double));
  if (fb) fb->append = reinterpret_cast<double*>(p + off);
  off += align256(append_scratch_doubles(npad, m) * sizeof(double));
  return off;
}

// factor the m padded matrices in fb.A, leave W = L^-1 in fb.W; synchronises to read the pivot status
int factor_and_invert(const FitBuffers& fb, int npad, int m, double jitter, cudaStream_t st) {
  const long long strideA = (long long)npad * npad, strideD = (long long)npad * 64;
  BO_CUDA(cudaMemsetAsync(fb.info, 0, sizeof(int) * 2 * m, st));
  int rc = cholesky_blocked(fb.A, npad, strideA, npad, m, fb.D, strideD, fb.info, fb.pol, nullptr, jitter, 1, st);
  if (rc) return rc;
  rc = tri_inverse(fb.W, npad, strideA, fb.A, npad, strideA, fb.D, strideD, fb.T, strideA / 2, npad, m, st);
  if (rc) return rc;
  int info_h[2 * BO_MAX_OBJECTIVES] = {0, 0, 0, 0, 0, 0, 0, 0};
  BO_CUDA(cudaMemcpyAsync(info_h, fb.info, sizeof(int) * 2 * m, cudaMemcpyDeviceToHost, st));
  BO_CUDA(cudaStreamSynchronize(st));
  g_last_clamped = 0;
  for (int o = 0; o < m; ++o) g_last_clamped += info_h[m + o];
  for (int o = 0; o < m; ++o) {
    if (info_h[o] != 0) {
      set_error("Matrix is not positive definite (objective %d, pivot %d)", o, info_h[o]);
      return BO_ERR_NOT_PD;
    }
  }
  return BO_OK;
}

}  // namespace
}  // namespace bo

using namespace bo;

extern "C" {

int bo_abi_version(void) { return BO_ABI_VERSION; }
const char* bo_last_error(void) { return g_err; }

int bo_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* l2_bytes, size_t* hbm_bytes) {
  int dev = 0;
  BO_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  BO_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (l2_bytes) *l2_bytes = (size_t)prop.l2CacheSize;
  if (hbm_bytes) *hbm_bytes = prop.totalGlobalMem;
  return BO_OK;
}

long long bo_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

int bo_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  g_prof.on = on != 0;
  return BO_OK;
}

// sums and clears the records of one kernel class (kernel < 0: every class)
static int profile_collect(int kernel, double* total_ms, long long* launches, double* work) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  double ms = 0.0, wk = 0.0;
  long long cnt = 0;
  std::vector<ProfRecord> keep;
  for (auto& r : g_prof.live) {
    if (kernel >= 0 && r.kernel != kernel) {
      keep.push_back(r);
      continue;
    }
    BO_CUDA(cudaEventSynchronize(r.stop));
    float t = 0.f;
    BO_CUDA(cudaEventElapsedTime(&t, r.start, r.stop));
    ms += t;
    wk += r.work;
    ++cnt;
    g_prof.pool.push_back(r.start);
    g_prof.pool.push_back(r.stop);
  }
  g_prof.live.swap(keep);
  if (total_ms) *total_ms = ms;
  if (launches) *launches = cnt;
  if (work) *work = wk;
  return BO_OK;
}

int bo_profile_read(double* total_ms, long long* launches, double* flops) {
  return profile_collect(BO_PROF_CONTRACTION, total_ms, launches, flops);
}

int bo_profile_read_kernel(int kernel, double* total_ms, long long* launches, double* work) {
  BO_REQUIRE(kernel >= -1 && kernel < BO_PROF_KINDS, "unknown kernel class");
  return profile_collect(kernel, total_ms, launches, work);
}

int bo_last_clamped_pivots(void) { return g_last_clamped; }

int bo_npad(int n) { return round_up(n, TM); }
size_t bo_wpack_doubles(int n) { return (size_t)wpack_tile_offset(round_up(n, TM) / TM) * TILE_DOUBLES; }

int bo_gram_f64(double* K_dev, int ldk, const double* x_dev, int ldx, int last_eval, int current_eval, int d, int m,
                const double* prior_variance_host, const double* length_scales_host, void* stream) {
  BO_REQUIRE(K_dev && x_dev && prior_variance_host && length_scales_host, "null pointer");
  BO_REQUIRE(d >= 1 && d <= BO_MAX_DIMS, "1 <= d <= 16");
  BO_REQUIRE(last_eval >= 0 && current_eval >= last_eval && current_eval <= ldk, "bad evaluation range");
  ObjParams hp;
  int rc = make_params(&hp, m, nullptr, prior_variance_host, length_scales_host, nullptr);
  if (rc) return rc;
  return gram(K_dev, ldk, (long long)ldk * ldk, x_dev, ldx, last_eval, current_eval, current_eval, d, m, hp, 0.0,
              (cudaStream_t)stream);
}

size_t bo_inverse_workspace_bytes(int n, int m) { return carve_fit(nullptr, nullptr, round_up(n, TM), m); }

int bo_inverse_f64(double* Kinv_dev, const double* K_dev, int ldk, int n, int m, double jitter, void* workspace_dev,
                   size_t workspace_bytes, void* stream) {
  BO_REQUIRE(Kinv_dev && K_dev && workspace_dev, "null pointer");
  BO_REQUIRE(n >= 1 && n <= ldk && m >= 1 && m <= BO_MAX_OBJECTIVES, "bad sizes");
  const int npad = round_up(n, TM);
  if (workspace_bytes < bo_inverse_workspace_bytes(n, m)) {
    set_error("inverse workspace too small");
    return BO_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  FitBuffers fb;
  carve_fit(&fb, workspace_dev, npad, m);
  pad_copy_kernel<<<dim3((npad + 127) / 128, npad, m), 128, 0, st>>>(fb.A, npad, K_dev, ldk, n, jitter);
  BO_LAUNCH_CHECK("pad_copy_kernel");
  int rc = factor_and_invert(fb, npad, m, jitter, st);
  if (rc) return rc;
  GemmArgs g;  // Kinv = W^T W   (C[i][j] = sum_k W[k][i] W[k][j])
  g.M = n; g.N = n; g.K = npad;
  g.A = fb.W; g.lda = npad; g.strideA = (long long)npad * npad;
  g.B = fb.W; g.ldb = npad; g.strideB = (long long)npad * npad;
  g.C = Kinv_dev; g.ldc = n; g.strideC = (long long)n * n;
  g.batch = m;
  rc = gemm(g, 1, 1, st);
  if (rc) return rc;
  BO_CUDA(cudaStreamSynchronize(st));
  return BO_OK;
}

size_t bo_fit_workspace_bytes(int n, int m) { return carve_fit(nullptr, nullptr, round_up(n, TM), m); }

int bo_gp_fit_f64(double* wpack_dev, double* alpha_dev, const double* x_dev, int ldx, const double* y_dev, int ldy,
                  int n, int d, int m, const double* prior_mean_host, const double* prior_variance_host,
                  const double* length_scales_host, double jitter, void* workspace_dev, size_t workspace_bytes,
                  void* stream) {
  BO_REQUIRE(wpack_dev && alpha_dev && x_dev && y_dev && workspace_dev, "null pointer");
  BO_REQUIRE(prior_mean_host && prior_variance_host && length_scales_host, "null hyper-parameter pointer");
  BO_REQUIRE(n >= 1 && d >= 1 && d <= BO_MAX_DIMS, "bad sizes");
  ObjParams hp;
  int rc = make_params(&hp, m, prior_mean_host, prior_variance_host, length_scales_host, nullptr);
  if (rc) return rc;
  const int npad = round_up(n, TM);
  if (workspace_bytes < bo_fit_workspace_bytes(n, m)) {
    set_error("fit workspace too small: %zu < %zu", workspace_bytes, bo_fit_workspace_bytes(n, m));
    return BO_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // algorithmic work: Cholesky n^3/3 + triangular inverse n^3/3 + (the W^T W u products are O(n^2)) per objective
  ProfileScope prof_scope(st, BO_PROF_FIT, (2.0 / 3.0) * m * (double)n * (double)n * (double)n);
  FitBuffers fb;
  carve_fit(&fb, workspace_dev, npad, m);
  const long long strideA = (long long)npad * npad;
  rc = gram(fb.A, npad, strideA, x_dev, ldx, 0, n, npad, d, m, hp, jitter, st, /*lower_only=*/true);
  if (rc) return rc;
  rc = factor_and_invert(fb, npad, m, jitter, st);
  if (rc) return rc;
  rc = compute_alpha(alpha_dev, fb.W, npad, strideA, y_dev, ldy, n, npad, m, hp, fb.scratch, st);
  if (rc) return rc;
  rc = pack_w(wpack_dev, (long long)bo_wpack_doubles(n), fb.W, npad, strideA, npad, n, m, st);
  if (rc) return rc;
  BO_CUDA(cudaStreamSynchronize(st));
  return BO_OK;
}

int bo_gp_append_f64(double* wpack_dev, double* alpha_dev, const double* x_dev, int ldx, const double* y_dev, int ldy,
                     int n_old, int n_new, int d, int m, const double* prior_mean_host,
                     const double* prior_variance_host, const double* length_scales_host, double jitter,
                     void* workspace_dev, size_t workspace_bytes, void* stream) {
  BO_REQUIRE(wpack_dev && alpha_dev && x_dev && y_dev && workspace_dev, "null pointer");
  BO_REQUIRE(prior_mean_host && prior_variance_host && length_scales_host, "null hyper-parameter pointer");
  BO_REQUIRE(n_old >= 1 && n_new > n_old && n_new - n_old <= BO_MAX_APPEND && d >= 1 && d <= BO_MAX_DIMS,
             "bad sizes (1 <= n_new - n_old <= BO_MAX_APPEND)");
  BO_REQUIRE(round_up(n_new, TM) == round_up(n_old, TM),
             "n_new leaves the 128-row padding of the existing factor: refit with bo_gp_fit_f64");
  ObjParams hp;
  int rc = make_params(&hp, m, prior_mean_host, prior_variance_host, length_scales_host, nullptr);
  if (rc) return rc;
  const int npad = round_up(n_new, TM);
  if (workspace_bytes < bo_fit_workspace_bytes(n_new, m)) {
    set_error("fit workspace too small: %zu < %zu", workspace_bytes, bo_fit_workspace_bytes(n_new, m));
    return BO_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // algorithmic work: two triangular products with the b new columns per objective
  ProfileScope prof_scope(st, BO_PROF_FIT, 2.0 * m * (double)(n_new - n_old) * (double)n_old * (double)n_old);
  FitBuffers fb;
  carve_fit(&fb, workspace_dev, npad, m);
  const long long strideA = (long long)npad * npad;
  BO_CUDA(cudaMemsetAsync(fb.info, 0, sizeof(int) * 2 * m, st));
  rc = append_rows(fb.A, fb.W, npad, strideA, x_dev, ldx, n_old, n_new, npad, d, m, hp, jitter, fb.append, fb.info, st);
  if (rc) return rc;
  rc = compute_alpha(alpha_dev, fb.W, npad, strideA, y_dev, ldy, n_new, npad, m, hp, fb.scratch, st);
  if (rc) return rc;
  rc = pack_w(wpack_dev, (long long)bo_wpack_doubles(n_new), fb.W, npad, strideA, npad, n_new, m, st);
  if (rc) return rc;
  int info_h[2 * BO_MAX_OBJECTIVES] = {0, 0, 0, 0, 0, 0, 0, 0};
  BO_CUDA(cudaMemcpyAsync(info_h, fb.info, sizeof(int) * 2 * m, cudaMemcpyDeviceToHost, st));
  BO_CUDA(cudaStreamSynchronize(st));
  g_last_clamped = 0;
  for (int o = 0; o < m; ++o) g_last_clamped += info_h[m + o];
  for (int o = 0; o < m; ++o) {
    if (info_h[o] != 0) {
      set_error("Matrix is not positive definite (objective %d, pivot %d of the appended rows)", o, info_h[o]);
      return BO_ERR_NOT_PD;
    }
  }
  return BO_OK;
}

size_t bo_score_workspace_bytes(int n, int m, long long n_cand) {
  return score_workspace_bytes(make_score_plan(n, m, n_cand));
}

int bo_score_f64(double* mu_dev, double* var_dev, double* std_mu_dev, double* std_var_dev, double* ucb_dev,
                 double* acq_dev, long long ld_out, const void* cand_dev, int cand_kind, int ldc, long long n_cand,
                 const double* x_dev, int ldx, int n, int d, int m, const double* wpack_dev, const double* alpha_dev,
                 const double* prior_mean_host, const double* prior_variance_host, const double* length_scales_host,
                 const double* betas_host, double min_variance, void* workspace_dev, size_t workspace_bytes,
                 void* stream) {
  BO_REQUIRE(cand_dev && x_dev && wpack_dev && alpha_dev && workspace_dev, "null pointer");
  BO_REQUIRE(prior_mean_host && prior_variance_host && length_scales_host && betas_host,
             "null hyper-parameter pointer");
  BO_REQUIRE(cand_kind == BO_CAND_F64 || cand_kind == BO_CAND_I64, "cand_kind");
  BO_REQUIRE(n >= 1 && d >= 1 && d <= BO_MAX_DIMS && ldc >= d && ld_out >= n_cand, "bad sizes");
  ObjParams hp;
  int rc = make_params(&hp, m, prior_mean_host, prior_variance_host, length_scales_host, betas_host);
  if (rc) return rc;
  ScoreOutputs out;
  out.mu = mu_dev; out.var = var_dev; out.std_mu = std_mu_dev; out.std_var = std_var_dev;
  out.ucb = ucb_dev; out.acq = acq_dev; out.ld = ld_out;
  return score_candidates(out, cand_dev, cand_kind, ldc, n_cand, x_dev, ldx, n, d, m, wpack_dev, alpha_dev, hp,
                          min_variance, workspace_dev, workspace_bytes, (cudaStream_t)stream);
}

size_t bo_i8_wq_bytes(int n) { return oz_wq_bytes(n); }
size_t bo_i8_wscale_doubles(int n, int m) { return (size_t)2 * m * round_up(n, OZ_TM); }

int bo_i8_quantize_w(uint8_t* wq_dev, double* wscale_dev, const double* wpack_dev, int n, int m, void* stream) {
  BO_REQUIRE(wq_dev && wscale_dev && wpack_dev, "null pointer");
  BO_REQUIRE(n >= 1 && n <= OZ_MAX_N && m >= 1 && m <= BO_MAX_OBJECTIVES, "bad sizes (int8 engine: n <= 16384)");
  return oz_quantize_w(wq_dev, wscale_dev, wpack_dev, n, m, (cudaStream_t)stream);
}

size_t bo_score_i8_workspace_bytes(int n, int m, long long n_cand) {
  return oz_workspace_bytes(make_oz_plan(n, m, n_cand));
}

int bo_score_i8(double* mu_dev, double* var_dev, double* std_mu_dev, double* std_var_dev, double* ucb_dev,
                double* acq_dev, long long ld_out, const void* cand_dev, int cand_kind, int ldc, long long n_cand,
                const double* x_dev, int ldx, int n, int d, int m, const uint8_t* wq_dev, const double* wscale_dev,
                const double* alpha_dev, const double* prior_mean_host, const double* prior_variance_host,
                const double* length_scales_host, const double* betas_host, double min_variance,
                void* workspace_dev, size_t workspace_bytes, void* stream) {
  BO_REQUIRE(cand_dev && x_dev && wq_dev && wscale_dev && alpha_dev && workspace_dev, "null pointer");
  BO_REQUIRE(prior_mean_host && prior_variance_host && length_scales_host && betas_host,
             "null hyper-parameter pointer");
  BO_REQUIRE(cand_kind == BO_CAND_F64 || cand_kind == BO_CAND_I64, "cand_kind");
  BO_REQUIRE(n >= 1 && d >= 1 && d <= BO_MAX_DIMS && ldc >= d && ld_out >= n_cand, "bad sizes");
  ObjParams hp;
  int rc = make_params(&hp, m, prior_mean_host, prior_variance_host, length_scales_host, betas_host);
  if (rc) return rc;
  ScoreOutputs out;
  out.mu = mu_dev; out.var = var_dev; out.std_mu = std_mu_dev; out.std_var = std_var_dev;
  out.ucb = ucb_dev; out.acq = acq_dev; out.ld = ld_out;
  return oz_score_candidates(out, cand_dev, cand_kind, ldc, n_cand, x_dev, ldx, n, d, m, wq_dev, wscale_dev,
                             alpha_dev, hp, min_variance, workspace_dev, workspace_bytes, (cudaStream_t)stream);
}

size_t bo_i8_guard_workspace_bytes(int n, int m, int d, long long n_cand, long long stride) {
  if (stride < 1) stride = 1;
  return oz_guard_workspace_bytes(n, m, d, n_cand, stride);
}

int bo_i8_guard_f64(double* worst_host, double* tau_host, const void* cand_dev, int cand_kind, int ldc,
                    long long n_cand, long long stride, const double* x_dev, int ldx, int n, int d, int m,
                    const uint8_t* wq_dev, const double* wscale_dev, const double* wpack_dev,
                    const double* alpha_dev, const double* prior_mean_host, const double* prior_variance_host,
                    const double* length_scales_host, double jitter, double min_variance, double tol,
                    void* workspace_dev, size_t workspace_bytes, void* stream) {
  BO_REQUIRE(cand_dev && x_dev && wq_dev && wscale_dev && wpack_dev && alpha_dev && workspace_dev, "null pointer");
  BO_REQUIRE(prior_mean_host && prior_variance_host && length_scales_host, "null hyper-parameter pointer");
  BO_REQUIRE(cand_kind == BO_CAND_F64 || cand_kind == BO_CAND_I64, "cand_kind");
  BO_REQUIRE(n >= 1 && n <= OZ_MAX_N && d >= 1 && d <= BO_MAX_DIMS && ldc >= d && n_cand >= 1 && stride >= 1,
             "bad sizes");
  ObjParams hp;
  int rc = make_params(&hp, m, prior_mean_host, prior_variance_host, length_scales_host, nullptr);
  if (rc) return rc;
  return oz_guard(worst_host, tau_host, cand_dev, cand_kind, ldc, n_cand, stride, x_dev, ldx, n, d, m, wq_dev,
                  wscale_dev, wpack_dev, alpha_dev, hp, jitter, min_variance, tol, workspace_dev, workspace_bytes,
                  (cudaStream_t)stream);
}

int bo_i8_peak_tops(double* tops_host, double seconds, void* stream) {
  BO_REQUIRE(tops_host && seconds > 0.0 && seconds < 5.0, "bad arguments");
  return oz_peak_tops(tops_host, seconds, (cudaStream_t)stream);
}

int bo_i8_kstar_digits(uint8_t* kq_dev, double* meandot_dev, const void* cand_dev, int cand_kind, int ldc,
                       long long n_cand, const double* x_dev, int ldx, int n, int d, int m,
                       const double* alpha_dev, const double* prior_variance_host,
                       const double* length_scales_host, void* stream) {
  BO_REQUIRE(kq_dev && meandot_dev && cand_dev && x_dev && alpha_dev, "null pointer");
  BO_REQUIRE(cand_kind == BO_CAND_F64 || cand_kind == BO_CAND_I64, "cand_kind");
  BO_REQUIRE(n >= 1 && n_cand >= 1 && d >= 1 && d <= BO_MAX_DIMS && ldc >= d, "bad sizes");
  ObjParams hp;
  int rc = make_params(&hp, m, nullptr, prior_variance_host, length_scales_host, nullptr);
  if (rc) return rc;
  const int tiles = (int)((n_cand + OZ_TN - 1) / OZ_TN), alloc_tiles = (tiles + 3) / 4 * 4;
  return oz_kstar_digits(kq_dev, meandot_dev, cand_dev, cand_kind, ldc, 0, n_cand, tiles, alloc_tiles, x_dev, ldx, n,
                         d, m, alpha_dev, hp, (cudaStream_t)stream);
}

int bo_i8_sumsq(double* q_dev, const uint8_t* wq_dev, const double* wscale_dev, const uint8_t* kq_dev, int n, int m,
                long long n_cand, int nsplit, const double* prior_variance_host, void* stream) {
  BO_REQUIRE(q_dev && wq_dev && wscale_dev && kq_dev, "null pointer");
  BO_REQUIRE(n >= 1 && n <= OZ_MAX_N && n_cand >= 1 && nsplit >= 1 && nsplit <= round_up(n, OZ_TM) / OZ_TM,
             "bad sizes");
  ObjParams hp;
  int rc = make_params(&hp, m, nullptr, prior_variance_host, nullptr, nullptr);
  if (rc) return rc;
  const int tiles = (int)((n_cand + OZ_TN - 1) / OZ_TN), alloc_tiles = (tiles + 3) / 4 * 4;
  return oz_sumsq(q_dev, (long long)alloc_tiles * OZ_TN, wq_dev, wscale_dev, kq_dev, n, m, tiles, alloc_tiles, nsplit,
                  hp, (cudaStream_t)stream);
}

int bo_acquisition_f64(double* std_mu_dev, double* std_var_dev, double* ucb_dev, double* acq_dev,
                       const double* mu_dev, const double* var_dev, long long ld, long long n_cand, int m,
                       const double* prior_mean_host, const double* prior_variance_host, const double* betas_host,
                       void* stream) {
  BO_REQUIRE(mu_dev && var_dev && prior_mean_host && prior_variance_host && betas_host, "null pointer");
  ObjParams hp;
  int rc = make_params(&hp, m, prior_mean_host, prior_variance_host, nullptr, betas_host);
  if (rc) return rc;
  return acquisition_only(std_mu_dev, std_var_dev, ucb_dev, acq_dev, mu_dev, var_dev, ld, n_cand, m, hp,
                          (cudaStream_t)stream);
}

size_t bo_topk_workspace_bytes(long long n_cand, int k) { return topk_workspace_bytes(n_cand, k); }

int bo_topk_f64(double* out_val_dev, long long* out_idx_dev, const double* acq_dev, long long n_cand, int k,
                long long index_base, void* workspace_dev, size_t workspace_bytes, void* stream) {
  BO_REQUIRE(out_val_dev && out_idx_dev && acq_dev && workspace_dev, "null pointer");
  BO_REQUIRE(k >= 1 && k <= BO_MAX_TOPK && n_cand >= 1, "1 <= k <= 1024, n_cand >= 1");
  ProfileScope prof_scope((cudaStream_t)stream, BO_PROF_TOPK, 8.0 * (double)n_cand);
  return topk_levels(out_val_dev, out_idx_dev, acq_dev, nullptr, n_cand, k, index_base, workspace_dev,
                     workspace_bytes, (cudaStream_t)stream);
}

int bo_topk_merge_f64(double* out_val_dev, long long* out_idx_dev, const double* val_dev, const long long* idx_dev,
                      int n_pairs, int k, void* workspace_dev, size_t workspace_bytes, void* stream) {
  BO_REQUIRE(out_val_dev && out_idx_dev && val_dev && idx_dev && workspace_dev, "null pointer");
  BO_REQUIRE(k >= 1 && k <= BO_MAX_TOPK && n_pairs >= 1, "1 <= k <= 1024, n_pairs >= 1");
  return topk_levels(out_val_dev, out_idx_dev, val_dev, idx_dev, n_pairs, k, 0, workspace_dev, workspace_bytes,
                     (cudaStream_t)stream);
}

int bo_match_rows_f64(uint8_t* out_flag_dev, const long long* idx_dev, int n_idx, long long index_base,
                      const void* cand_dev, int cand_kind, int ldc, const double* x_dev, int ldx, int n, int d,
                      void* stream) {
  BO_REQUIRE(out_flag_dev && idx_dev && cand_dev && (x_dev || n == 0), "null pointer");
  BO_REQUIRE(cand_kind == BO_CAND_F64 || cand_kind == BO_CAND_I64, "cand_kind");
  return match_rows(out_flag_dev, idx_dev, n_idx, index_base, cand_dev, cand_kind, ldc, x_dev, ldx, n, d,
                    (cudaStream_t)stream);
}

int bo_mask_evaluated_f64(double* out_dev, const double* acq_dev, const void* cand_dev, int cand_kind, int ldc,
                          long long n_cand, const double* x_dev, int ldx, int n, int d, void* stream) {
  BO_REQUIRE(out_dev && acq_dev && cand_dev && (x_dev || n == 0), "null pointer");
  BO_REQUIRE(cand_kind == BO_CAND_F64 || cand_kind == BO_CAND_I64, "cand_kind");
  BO_REQUIRE(d >= 1 && d <= BO_MAX_DIMS && ldc >= d && n >= 0, "bad sizes");
  return mask_evaluated(out_dev, acq_dev, cand_dev, cand_kind, ldc, n_cand, x_dev, ldx, n, d, (cudaStream_t)stream);
}

int bo_pareto_mask_f64(uint8_t* mask_dev, const double* y_dev, long long ldy, long long n, int m, void* stream) {
  BO_REQUIRE(mask_dev && y_dev, "null pointer");
  BO_REQUIRE(m >= 1 && m <= BO_MAX_OBJECTIVES && ldy >= m, "bad sizes");
  return pareto_mask(mask_dev, y_dev, ldy, n, y_dev, ldy, n, m, (cudaStream_t)stream);
}

int bo_pareto_mask_against_f64(uint8_t* mask_dev, const double* y_dev, long long ldy, long long n,
                               const double* z_dev, long long ldz, long long nz, int m, void* stream) {
  BO_REQUIRE(mask_dev && y_dev && (z_dev || nz == 0), "null pointer");
  BO_REQUIRE(m >= 1 && m <= BO_MAX_OBJECTIVES && ldy >= m && ldz >= m, "bad sizes");
  return pareto_mask(mask_dev, y_dev, ldy, n, z_dev, ldz, nz, m, (cudaStream_t)stream);
}

size_t bo_pareto_workspace_bytes(long long n, int m) { return pareto_filtered_workspace_bytes(n, m); }

int bo_pareto_mask_filtered_f64(uint8_t* mask_dev, const double* y_dev, long long ldy, long long n, int m,
                                void* workspace_dev, size_t workspace_bytes, void* stream) {
  BO_REQUIRE(mask_dev && y_dev && workspace_dev, "null pointer");
  BO_REQUIRE(m >= 1 && m <= BO_MAX_OBJECTIVES && ldy >= m, "bad sizes");
  return pareto_mask_filtered(mask_dev, y_dev, ldy, n, m, workspace_dev, workspace_bytes, (cudaStream_t)stream);
}

size_t bo_mll_workspace_bytes(int n, int m, int n_settings) { return mll_workspace_bytes(n, m, n_settings); }

int bo_mll_batched_f64(double* out_dev, const double* x_dev, int ldx, const double* y_dev, int ldy, int n, int d,
                       int m, const double* prior_mean_host, const double* length_scales_host,
                       const double* jitter_host, int n_settings, void* workspace_dev, size_t workspace_bytes,
                       void* stream) {
  BO_REQUIRE(out_dev && x_dev && y_dev && workspace_dev, "null pointer");
  BO_REQUIRE(prior_mean_host && length_scales_host && jitter_host, "null hyper-parameter pointer");
  BO_REQUIRE(n >= 1 && n <= 16384 && d >= 1 && d <= BO_MAX_DIMS && m >= 1 && m <= BO_MAX_OBJECTIVES &&
                 n_settings >= 1,
             "bad sizes");
  return mll_batched(out_dev, x_dev, ldx, y_dev, ldy, n, d, m, prior_mean_host, length_scales_host, jitter_host,
                     n_settings, workspace_dev, workspace_bytes, (cudaStream_t)stream);
}

int bo_hvi_f64(double* hvi_dev, const double* ucb_dev, long long ld, long long n_cand, int m,
               const double* front_dev, int n_front, const double* ref_host, void* stream) {
  BO_REQUIRE(hvi_dev && ucb_dev && ref_host && (front_dev || n_front == 0), "null pointer");
  BO_REQUIRE(m == 2 || m == 3, "exact HVI supports 2 or 3 objectives");
  return hvi(hvi_dev, ucb_dev, ld, n_cand, m, front_dev, n_front, ref_host, (cudaStream_t)stream);
}

size_t bo_hvi_front_doubles(int n_points, int m) { return hvi_front_doubles(n_points, m); }
size_t bo_hvi_workspace_bytes(int n_points, int m) { return hvi_workspace_bytes(n_points, m); }

int bo_hvi_prepare_f64(double* prepared_dev, int* n_front_dev, const double* points_dev, long long ld, int n_points,
                       int m, const double* ref_host, void* workspace_dev, size_t workspace_bytes, void* stream) {
  BO_REQUIRE(prepared_dev && n_front_dev && ref_host && workspace_dev && (points_dev || n_points == 0), "null pointer");
  BO_REQUIRE((m == 2 || m == 3) && ld >= m && n_points >= 0, "exact HVI supports 2 or 3 objectives");
  return hvi_prepare(prepared_dev, n_front_dev, points_dev, ld, n_points, m, ref_host, workspace_dev, workspace_bytes,
                     (cudaStream_t)stream);
}

static int make_hvi_spec(HviSpec* spec, const double* prepared_dev, const int* n_front_dev, int n_points, int m,
                         const double* ref_host) {
  BO_REQUIRE(prepared_dev && n_front_dev && ref_host, "null pointer (prepared front)");
  BO_REQUIRE(m == 2 || m == 3, "exact HVI supports 2 or 3 objectives");
  spec->prepared = prepared_dev;
  spec->n_front = n_front_dev;
  spec->cap = n_points > 0 ? n_points : 1;
  for (int o = 0; o < 3; ++o) spec->ref[o] = o < m ? ref_host[o] : 0.0;
  return BO_OK;
}

int bo_acquisition_hvi_f64(double* std_mu_dev, double* std_var_dev, double* ucb_dev, double* hvi_dev,
                           const double* mu_dev, const double* var_dev, long long ld, long long n_cand, int m,
                           const double* prior_mean_host, const double* prior_variance_host,
                           const double* betas_host, const double* prepared_dev, const int* n_front_dev,
                           int n_points, const double* ref_host, void* stream) {
  BO_REQUIRE(mu_dev && var_dev && prior_mean_host && prior_variance_host && betas_host, "null pointer");
  ObjParams hp;
  int rc = make_params(&hp, m, prior_mean_host, prior_variance_host, nullptr, betas_host);
  if (rc) return rc;
  HviSpec spec;
  rc = make_hvi_spec(&spec, prepared_dev, n_front_dev, n_points, m, ref_host);
  if (rc) return rc;
  return acquisition_hvi(std_mu_dev, std_var_dev, ucb_dev, hvi_dev, mu_dev, var_dev, ld, n_cand, m, hp, spec,
                         (cudaStream_t)stream);
}

int bo_score_hvi_f64(int engine, double* mu_dev, double* var_dev, double* std_mu_dev, double* std_var_dev,
                     double* ucb_dev, double* acq_dev, long long ld_out, const void* cand_dev, int cand_kind, int ldc,
                     long long n_cand, const double* x_dev, int ldx, int n, int d, int m, const void* factor_dev,
                     const double* wscale_dev, const double* alpha_dev, const double* prior_mean_host,
                     const double* prior_variance_host, const double* length_scales_host, const double* betas_host,
                     double min_variance, const double* prepared_dev, const int* n_front_dev, int n_points,
                     const double* ref_host, void* workspace_dev, size_t workspace_bytes, void* stream) {
  BO_REQUIRE(engine == 0 || engine == 1, "engine: 0 = FP64 DMMA, 1 = INT8 tensor cores");
  BO_REQUIRE(cand_dev && x_dev && factor_dev && alpha_dev && workspace_dev && (engine == 0 || wscale_dev),
             "null pointer");
  BO_REQUIRE(prior_mean_host && prior_variance_host && length_scales_host && betas_host,
             "null hyper-parameter pointer");
  BO_REQUIRE(cand_kind == BO_CAND_F64 || cand_kind == BO_CAND_I64, "cand_kind");
  BO_REQUIRE(n >= 1 && d >= 1 && d <= BO_MAX_DIMS && ldc >= d && ld_out >= n_cand, "bad sizes");
  ObjParams hp;
  int rc = make_params(&hp, m, prior_mean_host, prior_variance_host, length_scales_host, betas_host);
  if (rc) return rc;
  ScoreOutputs out;
  out.mu = mu_dev; out.var = var_dev; out.std_mu = std_mu_dev; out.std_var = std_var_dev;
  out.ucb = ucb_dev; out.acq = acq_dev; out.ld = ld_out;
  rc = make_hvi_spec(&out.hvi, prepared_dev, n_front_dev, n_points, m, ref_host);
  if (rc) return rc;
  if (engine == 0)
    return score_candidates(out, cand_dev, cand_kind, ldc, n_cand, x_dev, ldx, n, d, m,
                            static_cast<const double*>(factor_dev), alpha_dev, hp, min_variance, workspace_dev,
                            workspace_bytes, (cudaStream_t)stream);
  return oz_score_candidates(out, cand_dev, cand_kind, ldc, n_cand, x_dev, ldx, n, d, m,
                             static_cast<const unsigned char*>(factor_dev), wscale_dev, alpha_dev, hp, min_variance,
                             workspace_dev, workspace_bytes, (cudaStream_t)stream);
}

int bo_grid_i64(long long* out_dev, long long ld, const long long* lo_host, const long long* hi_host, int d,
                long long row0, long long rows, void* stream) {
  BO_REQUIRE(out_dev && lo_host && hi_host, "null pointer");
  BO_REQUIRE(d >= 1 && d <= BO_MAX_DIMS && ld >= d && row0 >= 0 && rows >= 0, "bad sizes");
  GridSpec g;
  long long total = 1;
  for (int k = 0; k < d; ++k) {
    BO_REQUIRE(hi_host[k] > lo_host[k], "empty grid dimension");
    g.lo[k] = lo_host[k];
    g.extent[k] = hi_host[k] - lo_host[k];
    total *= g.extent[k];
  }
  BO_REQUIRE(row0 + rows <= total, "row range exceeds the grid");
  if (rows == 0) return BO_OK;
  grid_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out_dev, ld, g, d, row0, rows);
  BO_LAUNCH_CHECK("grid_kernel");
  return BO_OK;
}

int bo_kstar_dense_f64(double* kstar_dev, long long ld_row, long long ld_obj, const double* x_dev, int ldx,
                       const void* cand_dev, int cand_kind, int ldc, long long n_cand, int last_eval,
                       int current_eval, int d, int m, const double* prior_variance_host,
                       const double* length_scales_host, void* stream) {
  BO_REQUIRE(kstar_dev && x_dev && cand_dev && prior_variance_host && length_scales_host, "null pointer");
  BO_REQUIRE(d >= 1 && d <= BO_MAX_DIMS, "1 <= d <= 16");
  ObjParams hp;
  int rc = make_params(&hp, m, nullptr, prior_variance_host, length_scales_host, nullptr);
  if (rc) return rc;
  return kstar_dense(kstar_dev, ld_row, ld_obj, x_dev, ldx, cand_dev, cand_kind, ldc, n_cand, last_eval, current_eval,
                     d, m, hp, (cudaStream_t)stream);
}

size_t bo_dense_workspace_bytes(int n, long long n_cand) { return dense_workspace_bytes(n, n_cand); }

int bo_mean_dense_f64(double* mu_dev, long long ld_mu, const double* kstar_dev, long long ld_row, long long ld_obj,
                      const double* kinv_dev, int ld_kinv, long long ld_kinv_obj, const double* y_dev, int ldy,
                      const double* prior_mean_host, int n, long long n_cand, int m, void* workspace_dev,
                      size_t workspace_bytes, void* stream) {
  BO_REQUIRE(mu_dev && kstar_dev && kinv_dev && y_dev && prior_mean_host && workspace_dev, "null pointer");
  ObjParams hp;
  int rc = make_params(&hp, m, prior_mean_host, nullptr, nullptr, nullptr);
  if (rc) return rc;
  return mean_dense(mu_dev, ld_mu, kstar_dev, ld_row, ld_obj, kinv_dev, ld_kinv, ld_kinv_obj, y_dev, ldy, hp, n,
                    n_cand, m, workspace_dev, workspace_bytes, (cudaStream_t)stream);
}

int bo_variance_dense_f64(double* var_dev, long long ld_var, const double* kstar_dev, long long ld_row,
                          long long ld_obj, const double* kinv_dev, int ld_kinv, long long ld_kinv_obj,
                          const double* prior_variance_host, double min_variance, int n, long long n_cand, int m,
                          void* workspace_dev, size_t workspace_bytes, void* stream) {
  BO_REQUIRE(var_dev && kstar_dev && kinv_dev && prior_variance_host && workspace_dev, "null pointer");
  ObjParams hp;
  int rc = make_params(&hp, m, nullptr, prior_variance_host, nullptr, nullptr);
  if (rc) return rc;
  return variance_dense(var_dev, ld_var, kstar_dev, ld_row, ld_obj, kinv_dev, ld_kinv, ld_kinv_obj, hp, min_variance,
                        n, n_cand, m, workspace_dev, workspace_bytes, (cudaStream_t)stream);
}

int bo_dgemm_nt_f64(double* C_dev, const double* A_dev, const double* B_dev, int n, void* stream) {
  BO_REQUIRE(C_dev && A_dev && B_dev && n >= 1, "bad arguments");
  GemmArgs g;
  g.M = g.N = g.K = n;
  g.A = A_dev; g.lda = n;
  g.B = B_dev; g.ldb = n;
  g.C = C_dev; g.ldc = n;
  return gemm(g, 0, 0, (cudaStream_t)stream);
}

}  // extern "C"
