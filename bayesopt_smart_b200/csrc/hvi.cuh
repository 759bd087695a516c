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
//   m = 2:  f0[cap] | h[cap] | S[cap]      h = objective 1 (ascending along the staircase),
//                                          S[i] = sum_{k<=i} (f0[k] - r0) (h[k] - h[k-1]),  h[-1] = r1
//           HVI(u) needs two binary searches and O(1) arithmetic -- no cap on P (tables in shared memory up to
//           4096 points, read through L1 beyond that):
//             a = #{f0 >= u0},  b = first index with h >= u1;  a > b  =>  u is dominated  =>  0
//             covered = (u0-r0)(h[a-1]-r1) + (S[b-1]-S[a-1]) + (f0[b]-r0)(u1-h[b-1])
//   m = 3:  f0[cap] | f1[cap] | f2[cap] | zlev[cap+1] | rank2[cap]   (rank2 stored as doubles)
//           zlev[s] = s-th largest objective 2 (zlev[P] = r2); slab s spans (zlev[s], min(zlev[s-1], u2)] and is
//           covered by the points with rank2 < s; inside a slab the covered area of the box [r, u] is the
//           f0-descending sweep.  O(P^2) per candidate; fronts up to 1024 points are swept from shared memory,
//           larger ones from global memory through L1 (no cap on P).
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
// largest power of two <= P (0 for an empty front): first stride of the branch-free searches below
__device__ __forceinline__ int hvi_top_stride(int P) { return P > 0 ? 1 << (31 - __clz(P)) : 0; }

// f0 / h / S may point to shared or global memory.  The two searches are branch-free binary searches with a fixed
// trip count (log2 P + 1 probes each: index arithmetic, one load, one compare, one select per probe).
__device__ __forceinline__ double hvi2_eval(double u0, double u1, const double* __restrict__ f0,
                                            const double* __restrict__ h, const double* __restrict__ S, int P,
                                            int top, double r0, double r1) {
  const double w0 = u0 - r0, w1 = u1 - r1;
  if (!(w0 > 0.0 && w1 > 0.0)) return 0.0;
  int a = 0, b = 0;  // a = #{f0 >= u0} (f0 descending);  b = #{h < u1} = first index with h >= u1 (h ascending)
  for (int step = top; step > 0; step >>= 1) {
    const int ja = a + step, jb = b + step;
    const double fa = f0[min(ja, P) - 1], hb = h[min(jb, P) - 1];
    a = (ja <= P && fa >= u0) ? ja : a;
    b = (jb <= P && hb < u1) ? jb : b;
  }
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
