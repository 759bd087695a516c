// hvi.cuh -- exact hypervolume improvement of one UCB vector against a PREPARED front (device functions shared by
// the fused scoring epilogue in score.cu and the stand-alone fused pass in hvi.cu).
//
// The reference's "HVI" is sum-UCB (acquisition.py:89-108; reference_point is allocated but unused,
// bayesian_optimization.py:425); the exact mode is the opt-in extension BASELINE.json's north star asks for:
// "UCB and 2-objective and 3-objective HVI become one fused per-candidate kernel against a sorted Pareto front".
// Specification: oracle/gp_oracle.py exact_hvi (parity unpinned -- nothing in the reference to compare with).
//
// Prepared front (hvi_prepare in hvi.cu; all on the device): points clipped to the reference point, dominated
// points removed, sorted by objective 0 DESCENDING (ties: objective 1 descending), `cap` = allocated points,
// *n_front = live points P.
//   m = 2:  f0[cap] | h[cap] | S[cap] | tab[HVI2_TAB]     h = objective 1 (ascending along the staircase),
//                                          S[i] = sum_{k<=i} (f0[k] - r0) (h[k] - h[k-1]),  h[-1] = r1
//           HVI(u) needs two SHORT searches (bucket tables, below) and O(1) arithmetic -- no cap on P (tables in
//           shared memory up to 4096 points, read through L1 beyond that):
//             a = #{f0 >= u0},  b = first index with h >= u1;  a > b  =>  u is dominated  =>  0
//             covered = (u0-r0)(h[a-1]-r1) + (S[b-1]-S[a-1]) + (f0[b]-r0)(u1-h[b-1])
//   m = 3:  f0[cap] | f1[cap] | f2[cap] | zlev[cap+1] | rank2[cap] | slabs[hvi3_slab_doubles(cap)]   (ints as doubles)
//           zlev[s] = s-th largest objective 2 (zlev[P] = r2); slab s spans (zlev[s], min(zlev[s-1], u2)] and is
//           covered by the points with rank2 < s; inside a slab the covered area of the box [r, u] is the
//           f0-descending sweep: O(P^2) per candidate.  Fronts of up to 1024 points use per-slab staircases prepared
//           once instead (hvi3_eval_slabs below: two binary searches per slab, O(P log P)); larger fronts fall back
//           to the sweep, read from global memory through L1 (no cap on P).
#pragma once
#include "common.cuh"

namespace bo {

struct HviSpec {
  const double* prepared = nullptr;  // nullptr: acquisition = sum-UCB (the reference's behaviour)
  const int* n_front = nullptr;      // device: live points
  int cap = 0;
  double ref[3] = {0.0, 0.0, 0.0};
};

#ifdef __CUDACC__
// m = 2 bucket tables (HVI2_TAB doubles after S): two searches of log2 P probes each dominated the instruction count
// of the fused pass, so the prepared front carries, for HVI2_NB equal-width buckets over the front's range of each
// objective, the search result at every bucket edge:
//   tab[k]              = A[k] = #{f0 >= f0_min + k w0},  k = 0..NB   (A[NB] forced to 0)
//   tab[NB + 1 + k]     = B[k] = #{h  <  h_min  + k w1},  k = 0..NB   (B[NB] forced to P)
//   tab[2 NB + 2 .. +5] = f0_min, 1 / w0, h_min, 1 / w1                (1 / w = 0 for a degenerate range)
// a(u0) = #{f0 >= u0} then lies in [A[k+2], A[k-1]] for the bucket k of u0 (one bucket of slack on each side absorbs
// the rounding of the bucket index) and is found by a binary search over that short range -- typically 1-2 probes.
constexpr int HVI2_NB = 256;
constexpr int HVI2_TAB = 2 * (HVI2_NB + 1) + 4;

// f0 / h / S / tab may point to shared or global memory
__device__ __forceinline__ double hvi2_eval(double u0, double u1, const double* __restrict__ f0,
                                            const double* __restrict__ h, const double* __restrict__ S,
                                            const double* __restrict__ tab, int P, double r0, double r1) {
  const double w0 = u0 - r0, w1 = u1 - r1;
  if (!(w0 > 0.0 && w1 > 0.0)) return 0.0;
  const double* sc = tab + 2 * HVI2_NB + 2;
  const int ka = max(0, min(HVI2_NB - 1, (int)fmin(fmax((u0 - sc[0]) * sc[1], 0.0), (double)HVI2_NB)));
  const int kb = max(0, min(HVI2_NB - 1, (int)fmin(fmax((u1 - sc[2]) * sc[3], 0.0), (double)HVI2_NB)));
  int lo = (int)tab[min(ka + 2, HVI2_NB)], hi = (int)tab[max(ka - 1, 0)];
  while (lo < hi) {  // a = #{f0 >= u0}: "f0[j-1] >= u0" holds exactly for j <= a (f0 descending)
    const int mid = (lo + hi + 1) >> 1;
    if (f0[mid - 1] >= u0) lo = mid;
    else hi = mid - 1;
  }
  const int a = lo;
  lo = (int)tab[HVI2_NB + 1 + max(kb - 1, 0)];
  hi = (int)tab[HVI2_NB + 1 + min(kb + 2, HVI2_NB)];
  while (lo < hi) {  // b = #{h < u1} = first index with h >= u1: "h[j-1] < u1" holds exactly for j <= b (h ascending)
    const int mid = (lo + hi + 1) >> 1;
    if (h[mid - 1] < u1) lo = mid;
    else hi = mid - 1;
  }
  const int b = lo;
  if (a > b) return 0.0;  // some point has f0 >= u0 and h >= u1: u adds nothing
  const double h_a = a > 0 ? h[a - 1] : r1;
  double covered = w0 * (h_a - r1);
  if (b > a) covered += S[b - 1] - (a > 0 ? S[a - 1] : 0.0);
  if (b < P) covered += (f0[b] - r0) * (u1 - (b > 0 ? h[b - 1] : r1));
  return w0 * w1 - covered;
}

// f0/f1/zlev/rank2 may point to shared or global memory; P live points
__device__ __forceinline__ double hvi3_eval(double u0, double u1, double u2, const double* f0, const double* f1,
                                            const double* zlev, const double* rank2, int P, double r0, double r1,
                                            double r2) {
  const double w0 = u0 - r0, w1 = u1 - r1;
  if (!(w0 > 0.0 && w1 > 0.0 && u2 > r2)) return 0.0;
  const double box = w0 * w1;
  double total = 0.0;
  for (int s = 0; s <= P; ++s) {
    const double z_hi = (s == 0) ? u2 : fmin(zlev[s - 1], u2);
    const double z_lo = zlev[s];
    const double thick = z_hi - z_lo;
    if (!(thick > 0.0)) continue;
    double covered = 0.0, best1 = r1;
    const double sd = (double)s;
    for (int p = 0; p < P; ++p) {
      if (rank2[p] < sd) {
        const double a0 = fmin(f0[p], u0), a1 = fmin(f1[p], u1);
        if (a1 > best1) {
          covered += (a0 - r0) * (a1 - best1);
          best1 = a1;
        }
      }
    }
    total += thick * (box - covered);
  }
  return total;
}

// m = 3, fronts of up to HVI3_SLAB_FRONT points: per-slab staircases prepared once (hvi_slabs3_kernel).  Slab s
// (1 <= s <= P, z in (zlev[s], zlev[s-1]]) is covered by the s points of largest objective 2; their 2-D staircase in
// (objective 0, objective 1) -- the candidate-independent part of the sweep above -- is stored with its prefix areas
// at offset s (s - 1) / 2 of slab_f0 / slab_h / slab_S, slab_len[s] entries.  A candidate then needs two binary
// searches per slab instead of a sweep over all points: O(P log P) instead of O(P^2).
constexpr int HVI3_SLAB_FRONT = 1024;
__host__ __device__ inline long long hvi3_slab_doubles(int cap) {  // slab_len[PS+1] + 3 * PS (PS + 1) / 2
  const long long ps = cap < HVI3_SLAB_FRONT ? cap : HVI3_SLAB_FRONT;
  return (ps + 1) + 3 * (ps * (ps + 1) / 2);
}

__device__ __forceinline__ double hvi3_eval_slabs(double u0, double u1, double u2, const double* __restrict__ zlev,
                                                  const double* __restrict__ slabs, int cap, int P, double r0,
                                                  double r1, double r2) {
  const double w0 = u0 - r0, w1 = u1 - r1;
  if (!(w0 > 0.0 && w1 > 0.0 && u2 > r2)) return 0.0;
  const double box = w0 * w1;
  const long long ps = cap < HVI3_SLAB_FRONT ? cap : HVI3_SLAB_FRONT;
  const long long T = ps * (ps + 1) / 2;
  const double* slab_len = slabs;
  const double* slab_f0 = slabs + ps + 1;
  const double* slab_h = slab_f0 + T;
  const double* slab_S = slab_h + T;
  // first slab that reaches below u2: zlev is descending, slabs above u2 have no thickness
  int lo = 0, hi = P;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (zlev[mid] >= u2) lo = mid + 1;
    else hi = mid;
  }
  double total = 0.0;
  for (int s = lo; s <= P; ++s) {
    const double z_hi = (s == 0) ? u2 : fmin(zlev[s - 1], u2);
    const double thick = z_hi - zlev[s];
    if (!(thick > 0.0)) continue;
    double covered = 0.0;
    if (s > 0) {
      const long long off = (long long)s * (s - 1) / 2;
      const int len = (int)__ldg(slab_len + s);
      const double* f0 = slab_f0 + off;
      const double* h = slab_h + off;
      int a0 = 0, a1 = len;  // a = #{f0 >= u0} (descending)
      while (a0 < a1) {
        const int mid = (a0 + a1 + 1) >> 1;
        if (__ldg(f0 + mid - 1) >= u0) a0 = mid;
        else a1 = mid - 1;
      }
      int b0 = 0, b1 = len;  // b = #{h < u1} (ascending)
      while (b0 < b1) {
        const int mid = (b0 + b1 + 1) >> 1;
        if (__ldg(h + mid - 1) < u1) b0 = mid;
        else b1 = mid - 1;
      }
      const int a = a0, b = b0;
      if (a > b) {
        covered = box;  // an active point dominates (u0, u1): the slab adds nothing
      } else {
        const double* S = slab_S + off;
        covered = w0 * ((a > 0 ? __ldg(h + a - 1) : r1) - r1);
        if (b > a) covered += __ldg(S + b - 1) - (a > 0 ? __ldg(S + a - 1) : 0.0);
        if (b < len) covered += (__ldg(f0 + b) - r0) * (u1 - (b > 0 ? __ldg(h + b - 1) : r1));
      }
    }
    total += thick * (box - covered);
  }
  return total;
}
#endif

size_t hvi_front_doubles(int n_points, int m);
size_t hvi_workspace_bytes(int n_points, int m);
// raw points (n, ld) -> prepared front (see above); asynchronous, no host synchronisation
int hvi_prepare(double* prepared, int* n_front, const double* points, long long ld, int n_points, int m,
                const double* ref, void* workspace, size_t workspace_bytes, cudaStream_t stream);
// stand-alone fused pass: standardise + UCB + exact HVI from (m, ld) mu / var arrays; outputs may be NULL
int acquisition_hvi(double* smu, double* svar, double* ucb, double* hvi_out, const double* mu, const double* var,
                    long long ld, long long n_cand, int m, const ObjParams& hp, const HviSpec& spec,
                    cudaStream_t stream);

}  // namespace bo
