// hvi.cu -- preparation of the Pareto front on the device and the fused UCB + exact-HVI pass (see hvi.cuh).
#include "hvi.cuh"

#include "select.cuh"

namespace bo {

namespace {

constexpr int HVI_SMEM_FRONT = 1024;  // m = 3: fronts up to this size are swept from shared memory

// q[i][o] = max(p[i][o], ref[o])   (fmax drops NaN: a NaN coordinate is clipped to the reference point)
__global__ void hvi_clip_kernel(double* __restrict__ q, const double* __restrict__ p, long long ld, int n, int m,
                                double r0, double r1, double r2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  q[(long long)i * m + 0] = fmax(p[i * ld + 0], r0);
  q[(long long)i * m + 1] = fmax(p[i * ld + 1], r1);
  if (m == 3) q[(long long)i * m + 2] = fmax(p[i * ld + 2], r2);
}

// order of the prepared front: objective 0 desc, then objective 1 desc, (then objective 2 desc), then index asc
__device__ __forceinline__ bool hvi_before(const double* __restrict__ q, int m, int j, int i) {
  for (int o = 0; o < m; ++o) {
    const double a = q[(long long)j * m + o], b = q[(long long)i * m + o];
    if (a != b) return a > b;
  }
  return j < i;
}

// rank of every kept point among the kept points (counting sort by comparison: O(n^2), n is a few thousand at
// most -- the evaluated points of a BO run), scatter into the prepared layout, count the live points
__global__ void __launch_bounds__(256)
    hvi_rank_scatter_kernel(double* __restrict__ prepared, int* __restrict__ n_front, const double* __restrict__ q,
                            const uint8_t* __restrict__ mask, int n, int m, int cap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int rank = 0, live = 0;
  for (int j = 0; j < n; ++j) {
    if (mask[j]) {
      ++live;
      if (j != i && hvi_before(q, m, j, i)) ++rank;
    }
  }
  if (i == 0) *n_front = live;
  if (!mask[i]) return;
  for (int o = 0; o < m; ++o) prepared[(long long)o * cap + rank] = q[(long long)i * m + o];
}

// m = 2: S[i] = sum_{k<=i} (f0[k]-r0)(h[k]-h[k-1]); one warp, sequential carry (P is small, once per iteration)
__global__ void hvi_prefix2_kernel(double* __restrict__ prepared, const int* __restrict__ n_front, int cap, double r0,
                                   double r1) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int P = *n_front;
  const double* f0 = prepared;
  const double* h = prepared + cap;
  double* S = prepared + 2LL * cap;
  double acc = 0.0, prev = r1;
  for (int i = 0; i < P; ++i) {
    acc += (f0[i] - r0) * (h[i] - prev);
    prev = h[i];
    S[i] = acc;
  }
}

// m = 2: the bucket tables of hvi2_eval (hvi.cuh); one thread per bucket edge, full binary search per edge
__global__ void __launch_bounds__(HVI2_NB + 1)
    hvi_buckets2_kernel(double* __restrict__ prepared, const int* __restrict__ n_front, int cap) {
  const int P = *n_front;
  const double* f0 = prepared;
  const double* h = prepared + cap;
  double* tab = prepared + 3LL * cap;
  const int k = threadIdx.x;  // 0..NB
  const double f0_min = P > 0 ? f0[P - 1] : 0.0, f0_max = P > 0 ? f0[0] : 0.0;
  const double h_min = P > 0 ? h[0] : 0.0, h_max = P > 0 ? h[P - 1] : 0.0;
  const double w0 = (f0_max - f0_min) / HVI2_NB, w1 = (h_max - h_min) / HVI2_NB;
  const bool ok0 = w0 > 0.0 && w0 < 1e300, ok1 = w1 > 0.0 && w1 < 1e300;
  // A[k] = #{f0 >= e_k}: degenerate range -> A = [P, 0, 0, ...] (every lookup then searches the whole front)
  int a = 0;
  if (k == 0) a = P;
  else if (k < HVI2_NB && ok0) {
    const double e = f0_min + k * w0;
    int lo = 0, hi = P;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (f0[mid - 1] >= e) lo = mid;
      else hi = mid - 1;
    }
    a = lo;
  }
  tab[k] = (double)a;
  int b = P;
  if (k == 0) b = 0;
  else if (k < HVI2_NB && ok1) {
    const double e = h_min + k * w1;
    int lo = 0, hi = P;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (h[mid - 1] < e) lo = mid;
      else hi = mid - 1;
    }
    b = lo;
  }
  tab[HVI2_NB + 1 + k] = (double)b;
  if (k == 0) {
    double* sc = tab + 2 * HVI2_NB + 2;
    sc[0] = f0_min;
    sc[1] = ok0 ? 1.0 / w0 : 0.0;
    sc[2] = h_min;
    sc[3] = ok1 ? 1.0 / w1 : 0.0;
  }
}

// m = 3: rank2[p] = position of p in objective-2-descending order (ties by position), zlev[rank] = objective 2
__global__ void __launch_bounds__(256)
    hvi_levels3_kernel(double* __restrict__ prepared, const int* __restrict__ n_front, int cap, double r2) {
  const int P = *n_front;
  const double* f2 = prepared + 2LL * cap;
  double* zlev = prepared + 3LL * cap;
  double* rank2 = prepared + 4LL * cap + 1;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p == 0) zlev[P] = r2;
  if (p >= P) return;
  int r = 0;
  const double mine = f2[p];
  for (int q = 0; q < P; ++q) {
    const double v = f2[q];
    r += (v > mine || (v == mine && q < p)) ? 1 : 0;
  }
  rank2[p] = (double)r;
  zlev[r] = mine;
}

// m = 3: the 2-D staircase (objectives 0, 1) of the s points of largest objective 2, for every s = 1..P, with prefix
// areas; one thread per slab walks the points in prepared order (objective 0 descending).  Only for P <= PS.
__global__ void __launch_bounds__(256)
    hvi_slabs3_kernel(double* __restrict__ prepared, const int* __restrict__ n_front, int cap, double r0, double r1) {
  const int P = *n_front;
  const long long ps = cap < HVI3_SLAB_FRONT ? cap : HVI3_SLAB_FRONT;
  if (P > ps) return;  // the evaluation falls back to the sweep
  const double* f0 = prepared;
  const double* f1 = prepared + cap;
  const double* rank2 = prepared + 4LL * cap + 1;
  double* slabs = prepared + 5LL * cap + 1;
  const long long T = ps * (ps + 1) / 2;
  double* slab_len = slabs;
  double* slab_f0 = slabs + ps + 1;
  double* slab_h = slab_f0 + T;
  double* slab_S = slab_h + T;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;  // slab index; slab 0 has no active point
  if (s == 0) slab_len[0] = 0.0;
  if (s < 1 || s > P) return;
  const long long off = (long long)s * (s - 1) / 2;
  const double sd = (double)s;
  double best1 = r1, acc = 0.0;
  int len = 0;
  for (int p = 0; p < P; ++p) {
    if (rank2[p] < sd && f1[p] > best1) {
      acc += (f0[p] - r0) * (f1[p] - best1);
      best1 = f1[p];
      slab_f0[off + len] = f0[p];
      slab_h[off + len] = f1[p];
      slab_S[off + len] = acc;
      ++len;
    }
  }
  slab_len[s] = (double)len;
}

// standardise + UCB + exact HVI in one pass over (m, ld) mu / var (numba_kernels.py:538-570, acquisition.py:55-81,
// then the exact hypervolume improvement instead of acquisition.py:104-108's sum).  The front tables are staged in
// shared memory (m = 2: up to HVI2_SMEM_FRONT points, m = 3: up to HVI_SMEM_FRONT), larger fronts are read through L1.
constexpr int HVI2_SMEM_FRONT = 4096;  // m = 2: 3 x 4096 doubles = 96 KB of dynamic shared memory at most

template <int MOBJ, bool VEC>
__global__ void __launch_bounds__(256)
    acquisition_hvi_kernel(double* __restrict__ smu_out, double* __restrict__ svar_out, double* __restrict__ ucb_out,
                           double* __restrict__ hvi_out, const double* __restrict__ mu_in,
                           const double* __restrict__ var_in, long long ld, long long n_cand, ObjParams hp,
                           HviSpec spec, int smem_points) {
  // dynamic shared memory sized by the host for min(allocated front points, limit): small fronts keep occupancy
  extern __shared__ double hvi_dyn[];
  const int SM = smem_points;
  double* sa = hvi_dyn;
  double* sb = sa + SM;
  double* sc = sb + SM;            // m = 2: S;  m = 3: zlev (SM + 1 entries)
  double* sd_ = sc + SM + 1;       // m = 2: bucket tables (HVI2_TAB);  m = 3: rank2
  const int P = *spec.n_front;
  const int cap = spec.cap;
  const double* f0 = spec.prepared;
  const double* f1 = spec.prepared + cap;              // m = 2: h
  const double* t2 = spec.prepared + (MOBJ == 3 ? 3LL : 2LL) * cap;  // m = 2: S;  m = 3: zlev
  const double* t3 = MOBJ == 3 ? spec.prepared + 4LL * cap + 1 : spec.prepared + 3LL * cap;  // rank2 / bucket tables
  if (P <= SM) {
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
      sa[p] = f0[p];
      sb[p] = f1[p];
      sc[p] = t2[p];
      if (MOBJ == 3) sd_[p] = t3[p];
    }
    if (MOBJ == 3 && threadIdx.x == 0) sc[P] = t2[P];
    if (MOBJ == 2)
      for (int p = threadIdx.x; p < HVI2_TAB; p += blockDim.x) sd_[p] = t3[p];
    __syncthreads();
    f0 = sa; f1 = sb; t2 = sc; t3 = sd_;
  }
  double sd[MOBJ];
#pragma unroll
  for (int o = 0; o < MOBJ; ++o) sd[o] = sqrt(hp.prior_var[o]);
  // a6, a7 for one candidate (numba_kernels.py:563-570, acquisition.py:52)
  auto ucb_of = [&](int o, double mu, double var, double& smu, double& svar) {
    smu = (mu - hp.prior_mean[o]) / sd[o];
    svar = var / hp.prior_var[o];
    // beta == 0 (raw vectors passed through bo_hvi_f64): no 0 * inf = NaN from an unused variance slot
    return hp.beta[o] != 0.0 ? smu + hp.beta[o] * sqrt(fabs(svar)) : smu;
  };
  auto hvi_of = [&](const double* u) {
    if (MOBJ == 2) return hvi2_eval(u[0], u[1], f0, f1, t2, t3, P, spec.ref[0], spec.ref[1]);
    if (P <= HVI3_SLAB_FRONT)  // per-slab staircases (global, through L1); zlev from shared memory
      return hvi3_eval_slabs(u[0], u[1], u[MOBJ - 1], t2, spec.prepared + 5LL * cap + 1, cap, P, spec.ref[0],
                             spec.ref[1], spec.ref[2]);
    return hvi3_eval(u[0], u[1], u[MOBJ - 1], f0, f1, t2, t3, P, spec.ref[0], spec.ref[1], spec.ref[2]);
  };
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (VEC) {
    // two candidates per thread: 16-byte loads / stores, and two independent search chains in flight per thread
    // all loads of an iteration are issued before any arithmetic, and the loads of the NEXT iteration are in flight
    // while this one's two searches run (the pass is latency-bound otherwise: ncu long-scoreboard stalls)
    const long long npair = n_cand / 2;
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double2 mu[MOBJ], va[MOBJ], mu_n[MOBJ], va_n[MOBJ];
    if (p < npair) {
#pragma unroll
      for (int o = 0; o < MOBJ; ++o) {
        mu[o] = *reinterpret_cast<const double2*>(mu_in + o * ld + 2 * p);
        va[o] = *reinterpret_cast<const double2*>(var_in + o * ld + 2 * p);
      }
    }
    for (; p < npair; p += stride) {
      const long long i = 2 * p, pn = p + stride;
      if (pn < npair) {
#pragma unroll
        for (int o = 0; o < MOBJ; ++o) {
          mu_n[o] = *reinterpret_cast<const double2*>(mu_in + o * ld + 2 * pn);
          va_n[o] = *reinterpret_cast<const double2*>(var_in + o * ld + 2 * pn);
        }
      }
      double ua[MOBJ], ub[MOBJ];
#pragma unroll
      for (int o = 0; o < MOBJ; ++o) {
        double2 smu, svar;
        ua[o] = ucb_of(o, mu[o].x, va[o].x, smu.x, svar.x);
        ub[o] = ucb_of(o, mu[o].y, va[o].y, smu.y, svar.y);
        if (smu_out) *reinterpret_cast<double2*>(smu_out + o * ld + i) = smu;
        if (svar_out) *reinterpret_cast<double2*>(svar_out + o * ld + i) = svar;
        if (ucb_out) *reinterpret_cast<double2*>(ucb_out + o * ld + i) = make_double2(ua[o], ub[o]);
      }
      const double ha = hvi_of(ua), hb = hvi_of(ub);
      if (hvi_out) *reinterpret_cast<double2*>(hvi_out + i) = make_double2(ha, hb);
#pragma unroll
      for (int o = 0; o < MOBJ; ++o) {
        mu[o] = mu_n[o];
        va[o] = va_n[o];
      }
    }
    return;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_cand; i += stride) {
    double u[MOBJ];
#pragma unroll
    for (int o = 0; o < MOBJ; ++o) {
      double smu, svar;
      u[o] = ucb_of(o, mu_in[o * ld + i], var_in[o * ld + i], smu, svar);
      if (smu_out) smu_out[o * ld + i] = smu;
      if (svar_out) svar_out[o * ld + i] = svar;
      if (ucb_out) ucb_out[o * ld + i] = u[o];
    }
    const double v = hvi_of(u);
    if (hvi_out) hvi_out[i] = v;
  }
}

}  // namespace

size_t hvi_front_doubles(int n_points, int m) {
  const size_t cap = n_points > 0 ? (size_t)n_points : 1;
  return (m == 2 ? 3 * cap + HVI2_TAB : 5 * cap + 1 + (size_t)hvi3_slab_doubles((int)cap)) + 2;  // layout: hvi.cuh
}

size_t hvi_workspace_bytes(int n_points, int m) {
  const size_t n = n_points > 0 ? (size_t)n_points : 1;
  return align256(n * m * sizeof(double)) + align256(n);  // clipped points + dominance mask
}

int hvi_prepare(double* prepared, int* n_front, const double* points, long long ld, int n_points, int m,
                const double* ref, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (workspace_bytes < hvi_workspace_bytes(n_points, m)) {
    set_error("hvi workspace too small");
    return BO_ERR_WORKSPACE;
  }
  if (n_points <= 0) {
    BO_CUDA(cudaMemsetAsync(n_front, 0, sizeof(int), stream));
    if (m == 3) {  // zlev[0] = r2 closes the single slab of an empty front
      const double r2 = ref[2];
      BO_CUDA(cudaMemcpyAsync(prepared + 3, &r2, sizeof(double), cudaMemcpyHostToDevice, stream));
    } else {
      hvi_buckets2_kernel<<<1, HVI2_NB + 1, 0, stream>>>(prepared, n_front, 1);  // all-zero tables (P = 0)
      BO_LAUNCH_CHECK("hvi_buckets2_kernel");
    }
    return BO_OK;
  }
  const int cap = n_points;
  double* q = static_cast<double*>(workspace);
  uint8_t* mask = reinterpret_cast<uint8_t*>(static_cast<unsigned char*>(workspace) +
                                             align256((size_t)n_points * m * sizeof(double)));
  const unsigned blocks = (unsigned)((n_points + 255) / 256);
  hvi_clip_kernel<<<blocks, 256, 0, stream>>>(q, points, ld, n_points, m, ref[0], ref[1], m == 3 ? ref[2] : 0.0);
  BO_LAUNCH_CHECK("hvi_clip_kernel");
  int rc = pareto_mask(mask, q, m, n_points, q, m, n_points, m, stream);  // warp-ballot dominance kernel (select.cu)
  if (rc) return rc;
  hvi_rank_scatter_kernel<<<blocks, 256, 0, stream>>>(prepared, n_front, q, mask, n_points, m, cap);
  BO_LAUNCH_CHECK("hvi_rank_scatter_kernel");
  if (m == 2) {
    hvi_prefix2_kernel<<<1, 32, 0, stream>>>(prepared, n_front, cap, ref[0], ref[1]);
    BO_LAUNCH_CHECK("hvi_prefix2_kernel");
    hvi_buckets2_kernel<<<1, HVI2_NB + 1, 0, stream>>>(prepared, n_front, cap);
    BO_LAUNCH_CHECK("hvi_buckets2_kernel");
  } else {
    hvi_levels3_kernel<<<blocks, 256, 0, stream>>>(prepared, n_front, cap, ref[2]);
    BO_LAUNCH_CHECK("hvi_levels3_kernel");
    const int ps = cap < HVI3_SLAB_FRONT ? cap : HVI3_SLAB_FRONT;
    hvi_slabs3_kernel<<<(ps + 1 + 255) / 256, 256, 0, stream>>>(prepared, n_front, cap, ref[0], ref[1]);
    BO_LAUNCH_CHECK("hvi_slabs3_kernel");
  }
  return BO_OK;
}

int acquisition_hvi(double* smu, double* svar, double* ucb, double* hvi_out, const double* mu, const double* var,
                    long long ld, long long n_cand, int m, const ObjParams& hp, const HviSpec& spec,
                    cudaStream_t stream) {
  if (n_cand <= 0) return BO_OK;
  long long blocks = (n_cand + 255) / 256;
  const long long cap = 32LL * device_sm_count();
  if (blocks > cap) blocks = cap;
  const int limit = m == 2 ? HVI2_SMEM_FRONT : HVI_SMEM_FRONT;
  const int pts = spec.cap < limit ? spec.cap : limit;
  const size_t smem = (size_t)(m == 2 ? 3 * pts + 1 + HVI2_TAB : 4 * pts + 1) * sizeof(double);
  auto aligned = [](const void* p) { return p == nullptr || (((uintptr_t)p) & 15) == 0; };
  const bool vec = (ld % 2 == 0) && (n_cand % 2 == 0) && aligned(smu) && aligned(svar) && aligned(ucb) &&
                   aligned(hvi_out) && aligned(mu) && aligned(var);
  const long long items = vec ? n_cand / 2 : n_cand;
  blocks = (items + 255) / 256;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
#define BO_HVI_LAUNCH(MO, V)                                                                                    \
  do {                                                                                                          \
    int rc = ensure_dynamic_smem(acquisition_hvi_kernel<MO, V>, smem);                                          \
    if (rc) return rc;                                                                                          \
    acquisition_hvi_kernel<MO, V><<<(unsigned)blocks, 256, smem, stream>>>(smu, svar, ucb, hvi_out, mu, var, ld, \
                                                                           n_cand, hp, spec, pts);              \
  } while (0)
  if (m == 2) {
    if (vec) BO_HVI_LAUNCH(2, true);
    else BO_HVI_LAUNCH(2, false);
  } else {
    if (vec) BO_HVI_LAUNCH(3, true);
    else BO_HVI_LAUNCH(3, false);
  }
#undef BO_HVI_LAUNCH
  BO_LAUNCH_CHECK("acquisition_hvi_kernel");
  return BO_OK;
}

}  // namespace bo
