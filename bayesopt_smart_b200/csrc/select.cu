// select.cu -- batch selection, Pareto filtering and exact hypervolume improvement.
//   topk_slice_kernel   per-CTA top-k of an 8192-element slice by repeated block arg-max with the total order
//                       (value desc, index asc, NaN last); applied level by level until one CTA remains.
//                       Replaces the full argsort of select_next_batch (acquisition.py:134).
//   match_rows_kernel   "candidate == some evaluated row" test of acquisition.py:139 for the few listed rows.
//   pareto_kernel       tiled O(n * nz) dominance test with warp-ballot early-out (pareto.py:27-45 semantics:
//                       maximisation, duplicates kept, NaN rows neither dominate nor are dominated).
//   hvi_kernel          exact 2-/3-objective hypervolume improvement against a front held in shared memory.
#include "select.cuh"

#include <stdlib.h>

#include "hvi.cuh"

namespace bo {

namespace {

constexpr int SEL_THREADS = 256;
constexpr int SEL_PER_THREAD = 32;
constexpr int SEL_SLICE = SEL_THREADS * SEL_PER_THREAD;

struct Best {
  unsigned long long key;
  long long idx;
  int owner;  // (tid << 5) | p, -1 = nothing left
};

__device__ __forceinline__ bool better(const Best& a, const Best& b) {
  if (a.owner < 0) return false;
  if (b.owner < 0) return true;
  if (a.key != b.key) return a.key > b.key;
  return a.idx < b.idx;
}
__device__ __forceinline__ Best shfl_best(const Best& v, int off) {
  Best r;
  r.key = __shfl_xor_sync(0xffffffffu, v.key, off);
  r.idx = __shfl_xor_sync(0xffffffffu, v.idx, off);
  r.owner = __shfl_xor_sync(0xffffffffu, v.owner, off);
  return r;
}
__device__ __forceinline__ double key_to_double(unsigned long long key) {
  if (key == 0ull) return __longlong_as_double(0x7ff8000000000000ll);
  const unsigned long long b = (key & 0x8000000000000000ull) ? (key & 0x7fffffffffffffffull) : ~key;
  return __longlong_as_double((long long)b);
}

// position of sample element e in the full array: one element per stride-wide window, at a hashed offset
// (a plain stride would alias with periodic score patterns of Cartesian candidate grids)
__device__ __forceinline__ long long sample_pos(long long e, int stride) {
  const unsigned h = (unsigned)e * 2654435761u;
  return e * stride + (long long)((h >> 8) % (unsigned)stride);
}

// Device-side switch between the filtered scan and its exact fallback (no host round trip): a gated launch runs
// only if "the survivor count is usable" (k <= *count <= capacity) equals `want_usable`, otherwise every CTA exits
// at once and leaves the outputs alone.
struct TopkGate {
  const int* count;  // nullptr: not gated
  int capacity;
  int want_usable;
};

// sample_stride > 0: the input is the virtual array in_val[sample_pos(e)], e < n_in, positions >= n_total
// are skipped.  n_in_dev != nullptr: the element count is read from device memory (clamped to n_in).
__global__ void __launch_bounds__(SEL_THREADS)
    topk_slice_kernel(double* __restrict__ out_val, long long* __restrict__ out_idx, const double* __restrict__ in_val,
                      const long long* __restrict__ in_idx, long long n_in, int k, long long index_base,
                      int sample_stride, long long n_total, const int* __restrict__ n_in_dev, TopkGate gate) {
  if (gate.count) {
    const int cnt = *gate.count;
    const int usable = (cnt >= k && cnt <= gate.capacity) ? 1 : 0;
    if (usable != gate.want_usable) return;
  }
  if (n_in_dev) {
    const long long cnt = *n_in_dev;
    n_in = cnt < n_in ? cnt : n_in;
  }
  __shared__ Best warp_best[SEL_THREADS / 32];
  __shared__ Best block_best;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long slice0 = (long long)blockIdx.x * SEL_SLICE;

  unsigned long long keys[SEL_PER_THREAD];
  long long idxs[SEL_PER_THREAD];
  unsigned valid = 0u;
#pragma unroll
  for (int p = 0; p < SEL_PER_THREAD; ++p) {
    const long long e = slice0 + (long long)p * SEL_THREADS + tid;
    keys[p] = 0ull;
    idxs[p] = -1;
    if (e < n_in) {
      const long long pos = sample_stride > 0 ? sample_pos(e, sample_stride) : e;
      if (sample_stride == 0 || pos < n_total) {
        const long long id = in_idx ? in_idx[pos] : index_base + pos;
        if (id >= 0) {
          keys[p] = order_key(in_val[pos]);
          idxs[p] = id;
          valid |= 1u << p;
        }
      }
    }
  }
  auto local_best = [&]() {
    Best b;
    b.key = 0ull; b.idx = 0; b.owner = -1;
#pragma unroll
    for (int p = 0; p < SEL_PER_THREAD; ++p) {
      if (valid & (1u << p)) {
        Best c;
        c.key = keys[p]; c.idx = idxs[p]; c.owner = (tid << 5) | p;
        if (better(c, b)) b = c;
      }
    }
    return b;
  };
  Best mine = local_best();
  for (int r = 0; r < k; ++r) {
    Best b = mine;
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      const Best o = shfl_best(b, off);
      if (better(o, b)) b = o;
    }
    if (lane == 0) warp_best[warp] = b;
    __syncthreads();
    if (warp == 0) {
      Best c;
      c.key = 0ull; c.idx = 0; c.owner = -1;
      if (lane < SEL_THREADS / 32) c = warp_best[lane];
#pragma unroll
      for (int off = 4; off; off >>= 1) {
        const Best o = shfl_best(c, off);
        if (better(o, c)) c = o;
      }
      if (lane == 0) block_best = c;
    }
    __syncthreads();
    const Best win = block_best;
    if (tid == 0) {
      const long long slot = (long long)blockIdx.x * k + r;
      if (win.owner >= 0) {
        out_val[slot] = key_to_double(win.key);
        out_idx[slot] = win.idx;
      } else {
        out_val[slot] = __longlong_as_double(0x7ff8000000000000ll);
        out_idx[slot] = -1;  // fewer than k elements in this slice
      }
    }
    if (win.owner >= 0 && (win.owner >> 5) == tid) {
      valid &= ~(1u << (win.owner & 31));
      mine = local_best();
    }
    // block_best is rewritten only after the next __syncthreads pair, so no extra barrier is needed here
  }
}

template <typename CT>
__global__ void match_rows_kernel(uint8_t* __restrict__ flag, const long long* __restrict__ idx, long long index_base,
                                  const CT* __restrict__ cand, int ldc, const double* __restrict__ x, int ldx, int n,
                                  int d) {
  __shared__ int hit;
  if (threadIdx.x == 0) hit = 0;
  __syncthreads();
  const long long id = idx[blockIdx.x];
  if (id >= 0) {
    const CT* row = cand + (id - index_base) * ldc;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      bool same = true;
      for (int k = 0; k < d; ++k) same = same && ((double)row[k] == x[(long long)e * ldx + k]);
      if (same) hit = 1;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) flag[blockIdx.x] = (uint8_t)hit;
}

// Exhaustive form of the exclusion test (acquisition.py:139) for the rare case in which the listed top-k rows were
// ALL evaluated points (more than BO_MAX_TOPK evaluated rows rank above the batch): out[i] = NaN (ranked last) when
// candidate i equals some evaluated row, acq[i] otherwise.  One candidate per thread, evaluated rows staged through
// shared memory 256 at a time and read as broadcasts; the first coordinate decides almost every comparison.
constexpr int MASK_ROWS = 256;
template <typename CT>
__global__ void __launch_bounds__(256)
    mask_evaluated_kernel(double* __restrict__ out, const double* __restrict__ acq, const CT* __restrict__ cand, int ldc,
                          long long n_cand, const double* __restrict__ x, int ldx, int n, int d) {
  __shared__ double xs[MASK_ROWS][BO_MAX_DIMS];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double c[BO_MAX_DIMS];
#pragma unroll
  for (int k = 0; k < BO_MAX_DIMS; ++k) c[k] = (k < d && i < n_cand) ? (double)cand[i * ldc + k] : 0.0;
  bool hit = false;
  for (int e0 = 0; e0 < n; e0 += MASK_ROWS) {
    const int cnt = min(MASK_ROWS, n - e0);
    __syncthreads();
    for (int t = threadIdx.x; t < cnt * d; t += blockDim.x) xs[t / d][t % d] = x[(long long)(e0 + t / d) * ldx + t % d];
    __syncthreads();
    for (int e = 0; e < cnt; ++e) {
      if (xs[e][0] == c[0]) {
        bool same = true;
#pragma unroll
        for (int k = 1; k < BO_MAX_DIMS; ++k) same = same && (k >= d || xs[e][k] == c[k]);
        hit = hit || same;
      }
    }
  }
  if (i < n_cand) out[i] = hit ? __longlong_as_double(0x7ff8000000000000ll) : acq[i];
}

// Streaming filter of the top-k scan: keep every element that is not worse than the threshold element
// (value desc, index asc order).  Survivors are appended with warp-aggregated atomics; their order is
// irrelevant because the exact kernel that follows ranks by (value, index).
__global__ void __launch_bounds__(256)
    topk_filter_kernel(double* __restrict__ out_val, long long* __restrict__ out_idx, int* __restrict__ count,
                       int capacity, const double* __restrict__ in_val, long long n, long long index_base,
                       const double* __restrict__ thr_val, const long long* __restrict__ thr_idx) {
  const unsigned long long key_t = order_key(*thr_val);
  const long long idx_t = *thr_idx;
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long rounds = (n + stride - 1) / stride;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long r = 0; r < rounds; ++r, i += stride) {
    bool keep = false;
    double v = 0.0;
    if (i < n) {
      v = in_val[i];
      const unsigned long long key = order_key(v);
      keep = key > key_t || (key == key_t && index_base + i <= idx_t);
    }
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (mask) {
      int base = 0;
      if (lane == __ffs(mask) - 1) base = atomicAdd(count, __popc(mask));
      base = __shfl_sync(0xffffffffu, base, __ffs(mask) - 1);
      if (keep) {
        const int slot = base + __popc(mask & ((1u << lane) - 1));
        if (slot < capacity) {
          out_val[slot] = v;
          out_idx[slot] = index_base + i;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------- Pareto
constexpr int PAR_TILE = 1024;

// n_dev / nz_dev (nullable): the row counts live on the device (stages of the filtered pass below); the grid is
// then sized for the largest possible n and surplus CTAs leave at once.
template <int MOBJ>
__global__ void __launch_bounds__(256)
    pareto_kernel(uint8_t* __restrict__ mask, const double* __restrict__ y, long long ldy, long long n,
                  const double* __restrict__ z, long long ldz, long long nz, const int* __restrict__ n_dev,
                  const int* __restrict__ nz_dev) {
  __shared__ double zs[MOBJ][PAR_TILE];
  if (n_dev) n = *n_dev;
  if (nz_dev) nz = *nz_dev;
  if ((long long)blockIdx.x * blockDim.x >= n) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double yi[MOBJ];
#pragma unroll
  for (int o = 0; o < MOBJ; ++o) yi[o] = (i < n) ? y[i * ldy + o] : 0.0;
  bool dominated = (i >= n);  // out-of-range lanes count as done for the early-out votes
  for (long long j0 = 0; j0 < nz; j0 += PAR_TILE) {
    const int cnt = (int)((nz - j0 < PAR_TILE) ? (nz - j0) : PAR_TILE);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
#pragma unroll
      for (int o = 0; o < MOBJ; ++o) zs[o][e] = z[(j0 + e) * ldz + o];
    }
    __syncthreads();
    // warp-ballot early-out, re-evaluated every 16 rows of z: a warp whose 32 points are all already
    // dominated stops comparing (callers put the strongest rows of z first)
    for (int jb = 0; jb < cnt; jb += 16) {
      if (__ballot_sync(0xffffffffu, !dominated) == 0u) break;
      const int je = min(cnt, jb + 16);
      for (int j = jb; j < je; ++j) {
        bool ge = true, gt = false;
#pragma unroll
        for (int o = 0; o < MOBJ; ++o) {
          const double zj = zs[o][j];  // broadcast read
          ge = ge && (zj >= yi[o]);
          gt = gt || (zj > yi[o]);
        }
        dominated = dominated || (ge && gt);
      }
    }
    if (__syncthreads_and(dominated ? 1 : 0)) break;
  }
  if (i < n) mask[i] = dominated ? 0 : 1;
}

// ------------------------------------------------------------------------------------------- exact HVI
constexpr int HVI_MAX_FRONT = 1024;

// front: (P, MOBJ) rows sorted by objective 0 DESCENDING.  For MOBJ == 3 slabs are cut at the front's
// objective-2 levels; inside a slab the covered area of the box [ref, u] is the f0-descending sweep over the
// points whose objective 2 reaches the slab.
template <int MOBJ>
__global__ void __launch_bounds__(256)
    hvi_kernel(double* __restrict__ hvi, const double* __restrict__ ucb, long long ld, long long n_cand,
               const double* __restrict__ front, int P, double r0, double r1, double r2) {
  __shared__ double f0[HVI_MAX_FRONT], f1[HVI_MAX_FRONT], f2[HVI_MAX_FRONT], zlev[HVI_MAX_FRONT + 1];
  __shared__ int rank2[HVI_MAX_FRONT];
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    f0[p] = fmax(front[(long long)p * MOBJ + 0], r0);
    f1[p] = fmax(front[(long long)p * MOBJ + 1], r1);
    f2[p] = (MOBJ == 3) ? fmax(front[(long long)p * MOBJ + 2], r2) : 0.0;
  }
  __syncthreads();
  if (MOBJ == 3) {
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
      int r = 0;
      for (int q = 0; q < P; ++q) r += (f2[q] > f2[p] || (f2[q] == f2[p] && q < p)) ? 1 : 0;
      rank2[p] = r;
      zlev[r] = f2[p];  // zlev[s] = s-th largest objective-2 value
    }
    if (threadIdx.x == 0) zlev[P] = r2;
    __syncthreads();
  }
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cand) return;
  const double u0 = ucb[i], u1 = ucb[ld + i];
  const double u2 = (MOBJ == 3) ? ucb[2 * ld + i] : 0.0;
  const double w0 = u0 - r0, w1 = u1 - r1;
  double total = 0.0;
  if (w0 > 0.0 && w1 > 0.0 && (MOBJ == 2 || u2 > r2)) {
    const double box = w0 * w1;
    if (MOBJ == 2) {
      double covered = 0.0, best1 = r1;
      for (int p = 0; p < P; ++p) {
        const double a0 = fmin(f0[p], u0), a1 = fmin(f1[p], u1);
        if (a1 > best1) {
          covered += (a0 - r0) * (a1 - best1);
          best1 = a1;
        }
      }
      total = box - covered;
    } else {
      // slab s: z in (zlev[s], z_hi], z_hi = +inf for s = 0 else zlev[s-1]; active points: rank2 < s
      for (int s = 0; s <= P; ++s) {
        const double z_hi = (s == 0) ? u2 : fmin(zlev[s - 1], u2);
        const double z_lo = zlev[s];
        const double thick = z_hi - z_lo;
        if (!(thick > 0.0)) continue;
        double covered = 0.0, best1 = r1;
        for (int p = 0; p < P; ++p) {
          if (rank2[p] < s) {
            const double a0 = fmin(f0[p], u0), a1 = fmin(f1[p], u1);
            if (a1 > best1) {
              covered += (a0 - r0) * (a1 - best1);
              best1 = a1;
            }
          }
        }
        total += thick * (box - covered);
      }
    }
  }
  hvi[i] = total;
}

}  // namespace

// =========================================================================================== host drivers
constexpr int FILTER_CAPACITY = 1 << 16;      // survivor pairs of the filtered scan
constexpr long long FILTER_MIN_N = 1 << 18;   // below this the plain multi-level scan is used
constexpr int FILTER_MAX_K = 256;

static size_t levels_bytes(long long n_in, int k) {
  const long long blocks0 = (n_in + SEL_SLICE - 1) / SEL_SLICE;
  const size_t pairs = (size_t)(blocks0 > 0 ? blocks0 : 1) * (size_t)k;
  return 2 * (align256(pairs * sizeof(double)) + align256(pairs * sizeof(long long)));
}

size_t topk_workspace_bytes(long long n_cand, int k) {
  // multi-level ping-pong buffers + survivor list + threshold pair + counter
  return levels_bytes(n_cand, k) + align256(FILTER_CAPACITY * sizeof(double)) +
         align256(FILTER_CAPACITY * sizeof(long long)) + align256((size_t)k * sizeof(double)) +
         align256((size_t)k * sizeof(long long)) + 256;
}

// stride > 0: first level reads the hashed sample of `val` (n_in = number of sample elements)
static int topk_levels_impl(double* out_val, long long* out_idx, const double* val, const long long* idx,
                            long long n_in, int k, long long index_base, int stride, long long n_total,
                            const int* n_in_dev, void* workspace, cudaStream_t stream,
                            TopkGate gate = TopkGate{nullptr, 0, 0}) {
  const long long blocks0 = (n_in + SEL_SLICE - 1) / SEL_SLICE;
  const size_t pairs = (size_t)(blocks0 > 0 ? blocks0 : 1) * (size_t)k;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  double* v[2];
  long long* ix[2];
  size_t off = 0;
  for (int b = 0; b < 2; ++b) {
    v[b] = reinterpret_cast<double*>(ws + off);
    off += align256(pairs * sizeof(double));
    ix[b] = reinterpret_cast<long long*>(ws + off);
    off += align256(pairs * sizeof(long long));
  }
  const double* cur_v = val;
  const long long* cur_i = idx;
  long long cur_n = n_in;
  int buf = 0;
  bool first = true;
  while (true) {
    long long blocks = (cur_n + SEL_SLICE - 1) / SEL_SLICE;
    if (blocks < 1) blocks = 1;
    const bool last = blocks == 1;
    double* ov = last ? out_val : v[buf];
    long long* oi = last ? out_idx : ix[buf];
    topk_slice_kernel<<<(unsigned)blocks, SEL_THREADS, 0, stream>>>(ov, oi, cur_v, cur_i, cur_n, k, index_base,
                                                                    first ? stride : 0, n_total,
                                                                    first ? n_in_dev : nullptr, gate);
    BO_LAUNCH_CHECK("topk_slice_kernel");
    if (last) break;
    cur_v = ov;
    cur_i = oi;
    cur_n = blocks * k;
    buf ^= 1;
    first = false;
  }
  return BO_OK;
}

int topk_levels(double* out_val, long long* out_idx, const double* val, const long long* idx, long long n_in, int k,
                long long index_base, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (workspace_bytes < topk_workspace_bytes(n_in, k)) {
    set_error("top-k workspace too small");
    return BO_ERR_WORKSPACE;
  }
  static const bool no_filter = [] {
    const char* e = getenv("BO_TOPK_NO_FILTER");
    return e && e[0] == '1';
  }();
  if (idx == nullptr && n_in >= FILTER_MIN_N && k <= FILTER_MAX_K && !no_filter) {
    // --- filtered scan: threshold from a hashed sample, one streaming pass, exact top-k of the survivors ---
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    size_t off = levels_bytes(n_in, k);
    double* sv = reinterpret_cast<double*>(ws + off);       off += align256(FILTER_CAPACITY * sizeof(double));
    long long* si = reinterpret_cast<long long*>(ws + off); off += align256(FILTER_CAPACITY * sizeof(long long));
    double* tv = reinterpret_cast<double*>(ws + off);       off += align256((size_t)k * sizeof(double));
    long long* ti = reinterpret_cast<long long*>(ws + off); off += align256((size_t)k * sizeof(long long));
    int* count = reinterpret_cast<int*>(ws + off);
    int stride = 16384 / k;  // expected survivors ~ k * stride = 16384 << capacity
    if (stride > 256) stride = 256;
    if (stride < 2) stride = 2;
    const long long n_sample = (n_in + stride - 1) / stride;
    int rc = topk_levels_impl(tv, ti, val, nullptr, n_sample, k, index_base, stride, n_in, nullptr, workspace,
                              stream);
    if (rc) return rc;
    BO_CUDA(cudaMemsetAsync(count, 0, sizeof(int), stream));
    long long blocks = (n_in + 256 * 8 - 1) / (256 * 8);
    const long long cap = 16LL * device_sm_count();
    if (blocks > cap) blocks = cap;
    topk_filter_kernel<<<(unsigned)blocks, 256, 0, stream>>>(sv, si, count, FILTER_CAPACITY, val, n_in, index_base,
                                                             tv + (k - 1), ti + (k - 1));
    BO_LAUNCH_CHECK("topk_filter_kernel");
    // The survivor count stays on the device.  Both continuations are enqueued and gated by it: the exact top-k of
    // the survivors runs when the count is usable (k <= count <= capacity); otherwise (fewer than k valid sample
    // elements, NaN threshold, massive ties) the plain multi-level scan of the whole array runs instead.  The
    // launches of the branch not taken exit in their first instruction, so there is no host synchronisation.
    rc = topk_levels_impl(out_val, out_idx, sv, si, FILTER_CAPACITY, k, 0, 0, 0, count, workspace, stream,
                          TopkGate{count, FILTER_CAPACITY, 1});
    if (rc) return rc;
    return topk_levels_impl(out_val, out_idx, val, idx, n_in, k, index_base, 0, 0, nullptr, workspace, stream,
                            TopkGate{count, FILTER_CAPACITY, 0});
  }
  return topk_levels_impl(out_val, out_idx, val, idx, n_in, k, index_base, 0, 0, nullptr, workspace, stream);
}

int match_rows(uint8_t* flag, const long long* idx, int n_idx, long long index_base, const void* cand, int cand_kind,
               int ldc, const double* x, int ldx, int n, int d, cudaStream_t stream) {
  if (n_idx <= 0) return BO_OK;
  if (cand_kind == BO_CAND_I64)
    match_rows_kernel<long long><<<n_idx, 128, 0, stream>>>(flag, idx, index_base,
                                                            static_cast<const long long*>(cand), ldc, x, ldx, n, d);
  else
    match_rows_kernel<double><<<n_idx, 128, 0, stream>>>(flag, idx, index_base, static_cast<const double*>(cand),
                                                         ldc, x, ldx, n, d);
  BO_LAUNCH_CHECK("match_rows_kernel");
  return BO_OK;
}

int mask_evaluated(double* out, const double* acq, const void* cand, int cand_kind, int ldc, long long n_cand,
                   const double* x, int ldx, int n, int d, cudaStream_t stream) {
  if (n_cand <= 0) return BO_OK;
  const unsigned grid = (unsigned)((n_cand + 255) / 256);
  if (cand_kind == BO_CAND_I64)
    mask_evaluated_kernel<long long><<<grid, 256, 0, stream>>>(out, acq, static_cast<const long long*>(cand), ldc,
                                                              n_cand, x, ldx, n, d);
  else
    mask_evaluated_kernel<double><<<grid, 256, 0, stream>>>(out, acq, static_cast<const double*>(cand), ldc, n_cand,
                                                           x, ldx, n, d);
  BO_LAUNCH_CHECK("mask_evaluated_kernel");
  return BO_OK;
}

static int pareto_mask_dev(uint8_t* mask, const double* y, long long ldy, long long n, const double* z, long long ldz,
                           long long nz, int m, const int* n_dev, const int* nz_dev, cudaStream_t stream) {
  if (n <= 0) return BO_OK;
  const unsigned grid = (unsigned)((n + 255) / 256);
  switch (m) {
    case 1: pareto_kernel<1><<<grid, 256, 0, stream>>>(mask, y, ldy, n, z, ldz, nz, n_dev, nz_dev); break;
    case 2: pareto_kernel<2><<<grid, 256, 0, stream>>>(mask, y, ldy, n, z, ldz, nz, n_dev, nz_dev); break;
    case 3: pareto_kernel<3><<<grid, 256, 0, stream>>>(mask, y, ldy, n, z, ldz, nz, n_dev, nz_dev); break;
    default: pareto_kernel<4><<<grid, 256, 0, stream>>>(mask, y, ldy, n, z, ldz, nz, n_dev, nz_dev); break;
  }
  BO_LAUNCH_CHECK("pareto_kernel");
  return BO_OK;
}

int pareto_mask(uint8_t* mask, const double* y, long long ldy, long long n, const double* z, long long ldz,
                long long nz, int m, cudaStream_t stream) {
  return pareto_mask_dev(mask, y, ldy, n, z, ldz, nz, m, nullptr, nullptr, stream);
}

// ------------------------------------------------------------------------------------------- filtered Pareto pass
// Large sets (n > 65 536: the 8 M UCB vectors of BASELINE config 4) are thinned before the n x n test: a point
// dominated by a member of the exact front of a strided SAMPLE is dominated, and efficient points always survive,
// so   [sample -> its front, strongest points first -> drop everything it dominates -> compact]  (four times), followed
// by the plain test among the survivors, gives exactly the mask of the direct test.  Every count stays on the device
// (no host round trip); grids are sized for the worst case and surplus CTAs exit.
constexpr int PAR_SAMPLE = 1 << 13;

// sample[j] = cur[j * step], step = ceil(n_cur / PAR_SAMPLE) (so at most PAR_SAMPLE rows); *ns = number of sample rows
__global__ void pareto_sample_kernel(double* __restrict__ sample, int* __restrict__ ns, const double* __restrict__ cur,
                                     long long ld, long long n_cur, const int* __restrict__ n_cur_dev, int m) {
  if (n_cur_dev) n_cur = *n_cur_dev;
  const long long step = n_cur > PAR_SAMPLE ? (n_cur + PAR_SAMPLE - 1) / PAR_SAMPLE : 1;
  const long long cnt = (n_cur + step - 1) / step;
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0) *ns = (int)cnt;
  if (j >= cnt) return;
  for (int o = 0; o < m; ++o) sample[j * m + o] = cur[j * step * ld + o];
}

// front = kept sample rows ordered by the sum of their objectives, descending (NaN sums last): most rows are then
// dominated within the first few comparisons and their warps leave the loop early.  Ordering only.  *nf = kept rows.
__global__ void __launch_bounds__(256)
    pareto_order_front_kernel(double* __restrict__ front, int* __restrict__ nf, const double* __restrict__ sample,
                              const uint8_t* __restrict__ smask, const int* __restrict__ ns_dev, int m) {
  const int ns = *ns_dev;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  auto key = [&](int r) {
    double s = 0.0;
    for (int o = 0; o < m; ++o) s += sample[(long long)r * m + o];
    return (s != s) ? -__longlong_as_double(0x7ff0000000000000ll) : s;
  };
  const double mine = key(i);
  int rank = 0, kept = 0;
  for (int j = 0; j < ns; ++j) {
    if (!smask[j]) continue;
    ++kept;
    const double kj = key(j);
    if (kj > mine || (kj == mine && j < i)) ++rank;
  }
  if (i == 0) *nf = kept;
  if (!smask[i]) return;
  for (int o = 0; o < m; ++o) front[(long long)rank * m + o] = sample[(long long)i * m + o];
}

// survivors (keep[i] != 0) appended to (rows_out, idx_out) with warp-aggregated atomics; order is irrelevant
__global__ void __launch_bounds__(256)
    pareto_compact_kernel(double* __restrict__ rows_out, long long* __restrict__ idx_out, int* __restrict__ count,
                          const double* __restrict__ cur, long long ld, const long long* __restrict__ idx_in,
                          const uint8_t* __restrict__ keep, long long n_cur, const int* __restrict__ n_cur_dev, int m) {
  if (n_cur_dev) n_cur = *n_cur_dev;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool k = i < n_cur && keep[i];
  const unsigned ballot = __ballot_sync(0xffffffffu, k);
  if (!ballot) return;
  int base = 0;
  if (lane == __ffs(ballot) - 1) base = atomicAdd(count, __popc(ballot));
  base = __shfl_sync(0xffffffffu, base, __ffs(ballot) - 1);
  if (k) {
    const long long slot = base + __popc(ballot & ((1u << lane) - 1));
    for (int o = 0; o < m; ++o) rows_out[slot * m + o] = cur[i * ld + o];
    idx_out[slot] = idx_in ? idx_in[i] : i;
  }
}

__global__ void pareto_scatter_kernel(uint8_t* __restrict__ mask, const uint8_t* __restrict__ fin,
                                      const long long* __restrict__ idx, const int* __restrict__ n_dev) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < *n_dev && fin[i]) mask[idx[i]] = 1;
}

size_t pareto_filtered_workspace_bytes(long long n, int m) {
  const size_t rows = align256((size_t)n * m * sizeof(double)), idx = align256((size_t)n * sizeof(long long));
  return 2 * (rows + idx) + 2 * align256((size_t)n) + 2 * align256((size_t)(PAR_SAMPLE + 1) * m * sizeof(double)) +
         align256(PAR_SAMPLE + 1) + 256;
}

int pareto_mask_filtered(uint8_t* mask, const double* y, long long ldy, long long n, int m, void* workspace,
                         size_t workspace_bytes, cudaStream_t stream) {
  if (n <= 0) return BO_OK;
  if (n > 0x7fffffffLL) {
    set_error("pareto: more than 2^31 - 1 rows");
    return BO_ERR_INVALID;
  }
  if (workspace_bytes < pareto_filtered_workspace_bytes(n, m)) {
    set_error("pareto workspace too small");
    return BO_ERR_WORKSPACE;
  }
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  size_t off = 0;
  const size_t rows_b = align256((size_t)n * m * sizeof(double)), idx_b = align256((size_t)n * sizeof(long long));
  double* rows[2];
  long long* idx[2];
  for (int b = 0; b < 2; ++b) {
    rows[b] = reinterpret_cast<double*>(ws + off);  off += rows_b;
    idx[b] = reinterpret_cast<long long*>(ws + off); off += idx_b;
  }
  uint8_t* keep = ws + off;    off += align256((size_t)n);
  uint8_t* fin = ws + off;     off += align256((size_t)n);
  double* sample = reinterpret_cast<double*>(ws + off);  off += align256((size_t)(PAR_SAMPLE + 1) * m * sizeof(double));
  double* front = reinterpret_cast<double*>(ws + off);   off += align256((size_t)(PAR_SAMPLE + 1) * m * sizeof(double));
  uint8_t* smask = ws + off;   off += align256(PAR_SAMPLE + 1);
  int* counts = reinterpret_cast<int*>(ws + off);  // [0] ns, [1] nf, [2 + r] survivors of round r
  BO_CUDA(cudaMemsetAsync(counts, 0, 8 * sizeof(int), stream));
  BO_CUDA(cudaMemsetAsync(mask, 0, (size_t)n, stream));
  const unsigned blocks_n = (unsigned)((n + 255) / 256);
  const unsigned blocks_s = (PAR_SAMPLE + 1 + 255) / 256;
  const double* cur = y;
  long long cur_ld = ldy;
  const long long* cur_idx = nullptr;
  const int* cur_n = nullptr;  // device count of the current set (nullptr: n, by value)
  // four rounds: each samples the SURVIVORS of the previous one, whose sample front hugs the true front more closely
  // (8 M three-objective UCB vectors: ~10 % / 2 % / 0.5 % / 0.2 % survive); the rounds after the first work on few
  // rows, so they cost little more than their launches, and they keep the final n_s x n_s test small
  constexpr int PAR_ROUNDS = 4;
  for (int round = 0; round < PAR_ROUNDS; ++round) {
    pareto_sample_kernel<<<blocks_s, 256, 0, stream>>>(sample, counts + 0, cur, cur_ld, n, cur_n, m);
    BO_LAUNCH_CHECK("pareto_sample_kernel");
    int rc = pareto_mask_dev(smask, sample, m, PAR_SAMPLE + 1, sample, m, PAR_SAMPLE + 1, m, counts + 0, counts + 0,
                             stream);
    if (rc) return rc;
    pareto_order_front_kernel<<<blocks_s, 256, 0, stream>>>(front, counts + 1, sample, smask, counts + 0, m);
    BO_LAUNCH_CHECK("pareto_order_front_kernel");
    rc = pareto_mask_dev(keep, cur, cur_ld, n, front, m, PAR_SAMPLE + 1, m, cur_n, counts + 1, stream);
    if (rc) return rc;
    pareto_compact_kernel<<<blocks_n, 256, 0, stream>>>(rows[round & 1], idx[round & 1], counts + 2 + round, cur, cur_ld,
                                                        cur_idx, keep, n, cur_n, m);
    BO_LAUNCH_CHECK("pareto_compact_kernel");
    cur = rows[round & 1];
    cur_ld = m;
    cur_idx = idx[round & 1];
    cur_n = counts + 2 + round;
  }
  int rc = pareto_mask_dev(fin, cur, m, n, cur, m, n, m, cur_n, cur_n, stream);
  if (rc) return rc;
  pareto_scatter_kernel<<<blocks_n, 256, 0, stream>>>(mask, fin, cur_idx, cur_n);
  BO_LAUNCH_CHECK("pareto_scatter_kernel");
  return BO_OK;
}

int hvi(double* out, const double* ucb, long long ld, long long n_cand, int m, const double* front, int n_front,
        const double* ref, cudaStream_t stream) {
  if (n_cand <= 0) return BO_OK;
  if (n_front > HVI_MAX_FRONT) {
    // larger fronts go through the prepared-front path (hvi.cu: no cap).  This entry point has no workspace
    // argument, so the few KB of front tables are stream-ordered temporaries.
    double* prepared = nullptr;
    int* count = nullptr;
    void* ws = nullptr;
    const size_t ws_bytes = hvi_workspace_bytes(n_front, m);
    BO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&prepared), hvi_front_doubles(n_front, m) * sizeof(double), stream));
    BO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&count), sizeof(int), stream));
    BO_CUDA(cudaMallocAsync(&ws, ws_bytes, stream));
    int rc = hvi_prepare(prepared, count, front, m, n_front, m, ref, ws, ws_bytes, stream);
    if (rc == BO_OK) {
      ObjParams unit;  // (u - 0) / sqrt(1) + 0 * sqrt(|.|) = u: the "UCB" of the fused kernel is the given vector
      memset(&unit, 0, sizeof(unit));
      for (int o = 0; o < BO_MAX_OBJECTIVES; ++o) unit.prior_var[o] = 1.0;
      HviSpec spec;
      spec.prepared = prepared;
      spec.n_front = count;
      spec.cap = n_front;
      for (int o = 0; o < 3; ++o) spec.ref[o] = o < m ? ref[o] : 0.0;
      rc = acquisition_hvi(nullptr, nullptr, nullptr, out, ucb, ucb, ld, n_cand, m, unit, spec, stream);
    }
    cudaFreeAsync(ws, stream);
    cudaFreeAsync(count, stream);
    cudaFreeAsync(prepared, stream);
    return rc;
  }
  const unsigned grid = (unsigned)((n_cand + 255) / 256);
  if (m == 2)
    hvi_kernel<2><<<grid, 256, 0, stream>>>(out, ucb, ld, n_cand, front, n_front, ref[0], ref[1], 0.0);
  else
    hvi_kernel<3><<<grid, 256, 0, stream>>>(out, ucb, ld, n_cand, front, n_front, ref[0], ref[1], ref[2]);
  BO_LAUNCH_CHECK("hvi_kernel");
  return BO_OK;
}

}  // namespace bo
