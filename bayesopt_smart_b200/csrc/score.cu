// score.cu -- the per-candidate hot path: K* -> mean / variance -> standardise -> UCB -> sum-UCB.
// Reference: update_k_star, update_mean, update_variance, standardize_objectives (numba_kernels.py:406-570),
// update_ucb, update_hypervolume_improvement (acquisition.py:55-108).
//
// Candidates are processed in chunks of `chunk_tiles` x 128.  Per chunk, three kernels:
//   1. kstar_pack_kernel   RBF cross kernel written ONCE as 16 KB tiles in DMMA B-fragment order (so K* is
//                          generated with N exps per candidate-objective, not N * N/256), fused with the
//                          posterior-mean dot product k*.alpha (warp-shuffle reduction).
//   2. trmm_sumsq_kernel   V = W K* on FP64 tensor cores (DMMA.8x8x4), persistent (one CTA per SM), W tiles and
//                          K* tiles streamed by cp.async.bulk (UBLKCP) through a 4-stage mbarrier ring; V never
//                          leaves registers: the epilogue reduces sum_i V[i,c]^2 per 128-row block with shuffles.
//   3. finalize_kernel     var = max(var0 - sum, min_var), mu, standardise, UCB, acq; coalesced writes.
// All reductions have a fixed order, so a candidate's result does not depend on the chunking or on how the
// candidate set is sharded across GPUs (bit-identical top-k for any rank count).
#include "score.cuh"

#include <stdlib.h>

#include "rbf.cuh"

namespace bo {

namespace {

// ------------------------------------------------------------------------------------------- K* tiles
// B tile (kt, c): 16 k x 128 candidates -> [wn(4)][j(4)][sp(2)][lane(32)][q(2)]
//   element = K*[kt*16 + (sp*2+q)*4 + t][c*128 + wn*32 + j*8 + g],  lane = g*4 + t
// A CTA (4 warps) produces half a tile column: warp w owns columns wn = 2*half + (w>>1),
// j in {2*(w&1), 2*(w&1)+1} for ALL k, so the mean dot product needs no cross-warp reduction.
// Training rows (x, alpha) are staged through shared memory in blocks of KS_ROWS rows; D is the padded
// dimension count (coordinates beyond d are zero on both sides, so the inner loop has no predicates).
// Rows >= n need no masking: the packed W has zero rows/columns there, and alpha is zero.
constexpr int KS_ROWS = 256;
__host__ __device__ constexpr int ks_row_stride(int d, int m) {
  const int even = (d + m + 1) & ~1;
  return (even % 8 == 0) ? even + 2 : even;
}

// 4 CTAs/SM (<= 128 registers); 6 CTAs/SM was measured and is not faster: the kernel is bound by its
// write stream (8*m*npad bytes per candidate), see DESIGN.md.
template <typename CT, int D, int MOBJ>
__global__ void __launch_bounds__(128, 4)
    kstar_pack_kernel(double* __restrict__ Kp, double* __restrict__ meandot, const CT* __restrict__ cand, int ldc,
                      long long cand0, long long n_cand, int chunk_tiles, long long ld_chunk,
                      const double* __restrict__ x, int ldx, int n, int npad, int d, const double* __restrict__ alpha,
                      ObjParams hp) {
  __shared__ double exp_tab[64];
  // row: D coordinates then MOBJ alphas, read as 16-byte pairs (D is even).  The row stride is even (16-byte
  // alignment) but not a multiple of 8 doubles, so the 4 rows a warp reads at a time (t = 0..3) start 4 banks apart
  // at least and their 16-byte accesses do not collide.
  constexpr int RS = ks_row_stride(D, MOBJ);
  __shared__ __align__(16) double xs[KS_ROWS][RS];
  const int tid = threadIdx.x;
  if (tid < 64) exp_tab[tid] = kExp2Tab[tid];
  const int c = blockIdx.x >> 1;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wn = (blockIdx.x & 1) * 2 + (warp >> 1), j0 = (warp & 1) * 2;
  const int nkt = npad / TK;

  double cc[2][D];
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    long long ci = cand0 + (long long)c * TN + wn * 32 + (j0 + jj) * 8 + g;
    if (ci >= n_cand) ci = n_cand - 1;  // tail tile: computed, never stored by finalize
#pragma unroll
    for (int k = 0; k < D; ++k) cc[jj][k] = (k < d) ? (double)cand[ci * ldc + k] : 0.0;
  }
  double macc[MOBJ][2];
  double coef[MOBJ], pvar[MOBJ];
#pragma unroll
  for (int o = 0; o < MOBJ; ++o) {
    macc[o][0] = macc[o][1] = 0.0;
    coef[o] = hp.neg_half_inv_ls2[o];
    pvar[o] = hp.prior_var[o];
  }

  for (int r0 = 0; r0 < npad; r0 += KS_ROWS) {
    __syncthreads();  // previous block fully consumed (also orders the exp_tab fill)
    for (int e = tid; e < KS_ROWS * (D + MOBJ); e += 128) {
      const int r = e / (D + MOBJ), k = e - r * (D + MOBJ);
      const int row = r0 + r;
      double v = 0.0;
      if (row < npad) {
        const int rr = row < n ? row : n - 1;  // padded rows reuse the last real point (finite values)
        if (k < D) v = (k < d) ? x[(long long)rr * ldx + k] : 0.0;
        else v = alpha[(long long)(k - D) * npad + row];
      }
      xs[r][k] = v;
    }
    __syncthreads();
    const int kt_end = min(nkt, (r0 + KS_ROWS) / TK);
    for (int kt = r0 / TK; kt < kt_end; ++kt) {
#pragma unroll
      for (int sp = 0; sp < 2; ++sp) {
        // the two training rows of this k-phase, fetched once as 16-byte pairs (coordinates, then alphas)
        double xa[RS], xb[RS];
        {
          const double2* pa = reinterpret_cast<const double2*>(xs[kt * TK - r0 + (sp * 2 + 0) * 4 + t]);
          const double2* pb = reinterpret_cast<const double2*>(xs[kt * TK - r0 + (sp * 2 + 1) * 4 + t]);
#pragma unroll
          for (int k = 0; k < (D + MOBJ + 1) / 2; ++k) {
            const double2 a2 = pa[k], b2 = pb[k];
            xa[2 * k] = a2.x; xa[2 * k + 1] = a2.y;
            xb[2 * k] = b2.x; xb[2 * k + 1] = b2.y;
          }
        }
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          double sa = 0.0, sb = 0.0;
#pragma unroll
          for (int k = 0; k < D; ++k) {
            const double da = xa[k] - cc[jj][k];
            const double db = xb[k] - cc[jj][k];
            sa = fma(da, da, sa);
            sb = fma(db, db, sb);
          }
#pragma unroll
          for (int o = 0; o < MOBJ; ++o) {
            const double va = pvar[o] * rbf_exp(sa * coef[o], exp_tab);
            const double vb = pvar[o] * rbf_exp(sb * coef[o], exp_tab);
            macc[o][jj] = fma(va, xa[D + o], macc[o][jj]);
            macc[o][jj] = fma(vb, xb[D + o], macc[o][jj]);
            double* dst = Kp + (((long long)o * chunk_tiles + c) * nkt + kt) * TILE_DOUBLES +
                          ((wn * 4 + j0 + jj) * 2 + sp) * 64 + lane * 2;
            *reinterpret_cast<double2*>(dst) = make_double2(va, vb);
          }
        }
      }
    }
  }
  // mean: reduce the 4 k-phases (t) of each column with warp shuffles
#pragma unroll
  for (int o = 0; o < MOBJ; ++o) {
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      double s = macc[o][jj];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (t == 0) meandot[(long long)o * ld_chunk + (long long)c * TN + wn * 32 + (j0 + jj) * 8 + g] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------- TRMM + sum of squares
constexpr int TR_STAGES = 4;
constexpr int TR_CONSUMER_WARPS = 8;  // 2 (rows) x 4 (candidates), warp tile 64 x 32
constexpr int TR_THREADS = (TR_CONSUMER_WARPS + 1) * 32;
constexpr size_t TR_SMEM = (size_t)TR_STAGES * 2 * TILE_DOUBLES * sizeof(double)  // A + B tiles
                           + 2 * 2 * TN * sizeof(double)                           // epilogue exchange
                           + 2 * TR_STAGES * sizeof(uint64_t);

// unit index -> (objective o, row-block pair pr, candidate tile c).  Candidate tiles are taken in groups of `group`:
// within a group every pair of every tile is scheduled back to back (pair slower, tile fastest), so the K* tiles of
// the group (group * npad * 128 * 8 bytes) are fetched from HBM once and re-read by the other pairs from L2.  With
// group = live_tiles this is the plain (o, pr, c) order, in which a K* tile comes from HBM once per pair.
__device__ __forceinline__ void trmm_decode(int u, int live_tiles, int npairs, int group, int& o, int& pr, int& c) {
  const int per_obj = live_tiles * npairs;
  o = u / per_obj;
  const int rem = u - o * per_obj;
  const int cg = rem / (group * npairs);
  const int r2 = rem - cg * group * npairs;
  const int in_group = min(group, live_tiles - cg * group);
  pr = r2 / in_group;
  c = cg * group + (r2 - pr * in_group);
}

// Persistent: the grid is one CTA per SM; CTA b walks work units b, b + gridDim.x, ...  A unit is
// (objective o, row-block pair pr, candidate tile c) with the candidate tile fastest, so the CTAs resident at
// any time stream the same W tiles (L2 hits).  The producer lane runs ahead across unit boundaries: the tiles
// of the next unit are already landing while the consumers reduce the current one.
__global__ void __launch_bounds__(TR_THREADS, 1)
    trmm_sumsq_kernel(double* __restrict__ part, long long ld_chunk, const double* __restrict__ Wp,
                      long long strideWp, const double* __restrict__ Kp, int nb, int chunk_tiles, int live_tiles,
                      int total_units, int group) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sA = reinterpret_cast<double*>(smem_raw);
  double* sB = sA + TR_STAGES * TILE_DOUBLES;
  double* red = sB + TR_STAGES * TILE_DOUBLES;  // [2 buffers][2 wm][128]
  uint64_t* full = reinterpret_cast<uint64_t*>(red + 2 * 2 * TN);
  uint64_t* empty = full + TR_STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npairs = (nb + 1) / 2;
  const int nkt_total = nb * KT_PER_BLOCK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TR_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], TR_CONSUMER_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == TR_CONSUMER_WARPS) {
    // ===== producer: one lane streams tiles with the bulk-copy engine =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        // grid decode uses the live tile count of this chunk; the K* / part layouts use chunk_tiles
        int o, pr, c;
        trmm_decode(u, live_tiles, npairs, group, o, pr, c);
        const int ib_first = nb - 1 - pr;  // heavy block first, its light partner second: nb + 1 k-blocks
        const int n_rb = (pr == ib_first) ? 1 : 2;
        const double* Wo = Wp + (long long)o * strideWp;
        const double* Bo = Kp + ((long long)o * chunk_tiles + c) * nkt_total * TILE_DOUBLES;
        for (int rb = 0; rb < n_rb; ++rb) {
          const int ib = rb == 0 ? ib_first : pr;
          const int nkt = (ib + 1) * KT_PER_BLOCK;
          const double* At = Wo + wpack_tile_offset(ib) * TILE_DOUBLES;
          for (int kt = 0; kt < nkt; ++kt) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], 2 * TILE_DOUBLES * sizeof(double));
            bulk_g2s(sA + stage * TILE_DOUBLES, At + (long long)kt * TILE_DOUBLES, TILE_DOUBLES * sizeof(double),
                     &full[stage]);
            bulk_g2s(sB + stage * TILE_DOUBLES, Bo + (long long)kt * TILE_DOUBLES, TILE_DOUBLES * sizeof(double),
                     &full[stage]);
            if (++stage == TR_STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
    return;
  }

  // ===== consumers =====
  const int wm = warp >> 2, wn = warp & 3;
  const int g = lane >> 2, t = lane & 3;
  int stage = 0;
  uint32_t phase = 0;
  int epi = 0;  // epilogue counter: alternates the exchange buffer
  for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
    int o, pr, c;
    trmm_decode(u, live_tiles, npairs, group, o, pr, c);
    const int ib_first = nb - 1 - pr;
    const int n_rb = (pr == ib_first) ? 1 : 2;
    for (int rb = 0; rb < n_rb; ++rb, ++epi) {
      const int ib = rb == 0 ? ib_first : pr;
      const int nkt = (ib + 1) * KT_PER_BLOCK;
      // Inside the diagonal 128x128 block W is lower triangular: for this warp's rows (wm*64 + 8i + g) the
      // k-tile kd (16 wide, counted from the start of the diagonal block) only meets non-zeros for row
      // atoms i >= 2*(kd - 4*wm); tiles with kd - 4*wm >= 4 are skipped entirely.
      const int kt_diag = ib * KT_PER_BLOCK;
      double acc[8][4][2];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

      for (int kt = 0; kt < nkt; ++kt) {
        mbar_wait(&full[stage], phase);
        const double* a_base = sA + stage * TILE_DOUBLES + wm * 1024 + lane * 2;
        const double* b_base = sB + stage * TILE_DOUBLES + wn * 512 + lane * 2;
        const int i_min = 2 * (kt - kt_diag - 4 * wm);  // <= 0: every row atom is live
        if (i_min <= 0) {
#pragma unroll
          for (int sp = 0; sp < 2; ++sp) {
            double2 b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = lds128(b_base + (j * 2 + sp) * 64);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const double2 a = lds128(a_base + (i * 2 + sp) * 64);
#pragma unroll
              for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a.x, b[j].x);
#pragma unroll
              for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a.y, b[j].y);
            }
          }
        } else if (i_min < 8) {
#pragma unroll
          for (int sp = 0; sp < 2; ++sp) {
            double2 b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = lds128(b_base + (j * 2 + sp) * 64);
#pragma unroll
            for (int i = 2; i < 8; ++i) {
              if (i >= i_min) {  // warp-uniform
                const double2 a = lds128(a_base + (i * 2 + sp) * 64);
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a.x, b[j].x);
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a.y, b[j].y);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == TR_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }

      // epilogue: sum over this warp's 64 rows of V^2, per candidate column
      double* rbuf = red + (epi & 1) * 2 * TN;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          double s = 0.0;
#pragma unroll
          for (int i = 0; i < 8; ++i) s = fma(acc[i][j][r], acc[i][j][r], s);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          s += __shfl_xor_sync(0xffffffffu, s, 8);
          s += __shfl_xor_sync(0xffffffffu, s, 16);
          if (g == 0) rbuf[wm * TN + wn * 32 + j * 8 + 2 * t + r] = s;
        }
      }
      // consumer warps only.  The exchange buffer alternates, and a thread can only reach the barrier of
      // epilogue e+1 after its reads of epilogue e, so buffer e&1 is free again when epilogue e+2 writes it.
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x < TN)
        part[((long long)o * nb + ib) * ld_chunk + (long long)c * TN + threadIdx.x] =
            rbuf[threadIdx.x] + rbuf[TN + threadIdx.x];
    }
  }
}

// ------------------------------------------------------------------------------------------- finalize
// a4..a8 for one candidate from the partial sums; returns sum-UCB and leaves the UCB vector in u[]
__device__ __forceinline__ double finalize_candidate(double* __restrict__ mu_out, double* __restrict__ var_out,
                                                     double* __restrict__ smu_out, double* __restrict__ svar_out,
                                                     double* __restrict__ ucb_out, long long ld_out, long long gi,
                                                     long long li, const double* __restrict__ part,
                                                     const double* __restrict__ meandot, long long ld_chunk, int nb,
                                                     int m, const ObjParams& hp, double min_variance, double* u) {
  double acq = 0.0;  // sequential sum from 0.0 (acquisition.py:108)
  for (int o = 0; o < m; ++o) {
    double q = 0.0;
    for (int ib = 0; ib < nb; ++ib) q += part[((long long)o * nb + ib) * ld_chunk + li];
    const double var = fmax(hp.prior_var[o] - q, min_variance);               // numba_kernels.py:532-535
    const double mu = hp.prior_mean[o] + meandot[(long long)o * ld_chunk + li];  // :486-488
    const double smu = (mu - hp.prior_mean[o]) / sqrt(hp.prior_var[o]);       // :563-565
    const double svar = var / hp.prior_var[o];                                // :568-570
    const double ucb = smu + hp.beta[o] * sqrt(fabs(svar));                   // acquisition.py:52
    acq = acq + ucb;
    u[o] = ucb;
    if (mu_out) mu_out[o * ld_out + gi] = mu;
    if (var_out) var_out[o * ld_out + gi] = var;
    if (smu_out) smu_out[o * ld_out + gi] = smu;
    if (svar_out) svar_out[o * ld_out + gi] = svar;
    if (ucb_out) ucb_out[o * ld_out + gi] = ucb;
  }
  return acq;
}

__global__ void finalize_kernel(double* __restrict__ mu_out, double* __restrict__ var_out,
                                double* __restrict__ smu_out, double* __restrict__ svar_out,
                                double* __restrict__ ucb_out, double* __restrict__ acq_out, long long ld_out,
                                long long cand0, long long n_cand, const double* __restrict__ part,
                                const double* __restrict__ meandot, long long ld_chunk, int chunk_cands, int nb, int m,
                                ObjParams hp, double min_variance) {
  const long long li = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long gi = cand0 + li;
  if (li >= chunk_cands || gi >= n_cand) return;
  double u[BO_MAX_OBJECTIVES];
  const double acq = finalize_candidate(mu_out, var_out, smu_out, svar_out, ucb_out, ld_out, gi, li, part, meandot,
                                        ld_chunk, nb, m, hp, min_variance, u);
  if (acq_out) acq_out[gi] = acq;
}

// opt-in exact mode: the same epilogue, but the acquisition value is the exact hypervolume improvement of the UCB
// vector against the prepared front (hvi.cuh) -- UCB and HVI in ONE per-candidate pass, the UCB array is not
// re-read from HBM.  m = 3 fronts of up to 1024 points are staged in shared memory.
template <int MOBJ>
__global__ void __launch_bounds__(256)
    finalize_hvi_kernel(double* __restrict__ mu_out, double* __restrict__ var_out, double* __restrict__ smu_out,
                        double* __restrict__ svar_out, double* __restrict__ ucb_out, double* __restrict__ acq_out,
                        long long ld_out, long long cand0, long long n_cand, const double* __restrict__ part,
                        const double* __restrict__ meandot, long long ld_chunk, int chunk_cands, int nb,
                        ObjParams hp, double min_variance, HviSpec spec) {
  constexpr int SM = MOBJ == 3 ? 1024 : 1;
  __shared__ double sf0[SM], sf1[SM], szl[SM + 1], srk[SM];
  const int P = *spec.n_front;
  const int cap = spec.cap;
  const double* f0 = spec.prepared;
  const double* f1 = spec.prepared + cap;
  const double* zlev = spec.prepared + 3LL * cap;
  const double* rank2 = spec.prepared + 4LL * cap + 1;
  if (MOBJ == 3 && P <= SM) {
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
      sf0[p] = f0[p];
      sf1[p] = f1[p];
      srk[p] = rank2[p];
      szl[p] = zlev[p];
    }
    if (threadIdx.x == 0) szl[P] = zlev[P];
    __syncthreads();
    f0 = sf0; f1 = sf1; zlev = szl; rank2 = srk;
  }
  const long long li = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long gi = cand0 + li;
  if (li >= chunk_cands || gi >= n_cand) return;
  double u[BO_MAX_OBJECTIVES];
  finalize_candidate(mu_out, var_out, smu_out, svar_out, ucb_out, ld_out, gi, li, part, meandot, ld_chunk, nb, MOBJ, hp,
                     min_variance, u);
  double v;
  if (MOBJ == 2)
    v = hvi2_eval(u[0], u[1], f0, f1, spec.prepared + 2LL * cap, spec.prepared + 3LL * cap, P, spec.ref[0], spec.ref[1]);
  else if (P <= HVI3_SLAB_FRONT)
    v = hvi3_eval_slabs(u[0], u[1], u[MOBJ - 1], zlev, spec.prepared + 5LL * cap + 1, cap, P, spec.ref[0], spec.ref[1],
                        spec.ref[2]);
  else v = hvi3_eval(u[0], u[1], u[MOBJ - 1], f0, f1, zlev, rank2, P, spec.ref[0], spec.ref[1], spec.ref[2]);
  if (acq_out) acq_out[gi] = v;
}

// stand-alone a6..a8 on existing arrays (HBM bound: reads 2m, writes up to 3m+1 doubles per candidate).
// Two candidates per thread with 16-byte accesses; VEC = false is the unaligned / odd-length fallback.
template <int MOBJ, bool VEC>
__global__ void __launch_bounds__(256)
    acquisition_kernel(double* __restrict__ smu_out, double* __restrict__ svar_out, double* __restrict__ ucb_out,
                       double* __restrict__ acq_out, const double* __restrict__ mu_in,
                       const double* __restrict__ var_in, long long ld, long long n_cand, ObjParams hp) {
  double sd[MOBJ];
#pragma unroll
  for (int o = 0; o < MOBJ; ++o) sd[o] = sqrt(hp.prior_var[o]);  // numba_kernels.py:565 (one sqrt per objective)
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long npair = VEC ? n_cand / 2 : n_cand;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npair; p += stride) {
    if (VEC) {
      const long long i = 2 * p;
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int o = 0; o < MOBJ; ++o) {
        const double2 mu = *reinterpret_cast<const double2*>(mu_in + o * ld + i);
        const double2 va = *reinterpret_cast<const double2*>(var_in + o * ld + i);
        const double2 smu = make_double2((mu.x - hp.prior_mean[o]) / sd[o], (mu.y - hp.prior_mean[o]) / sd[o]);
        const double2 svar = make_double2(va.x / hp.prior_var[o], va.y / hp.prior_var[o]);
        const double2 ucb = make_double2(smu.x + hp.beta[o] * sqrt(fabs(svar.x)), smu.y + hp.beta[o] * sqrt(fabs(svar.y)));
        a0 = a0 + ucb.x;
        a1 = a1 + ucb.y;
        if (smu_out) *reinterpret_cast<double2*>(smu_out + o * ld + i) = smu;
        if (svar_out) *reinterpret_cast<double2*>(svar_out + o * ld + i) = svar;
        if (ucb_out) *reinterpret_cast<double2*>(ucb_out + o * ld + i) = ucb;
      }
      if (acq_out) *reinterpret_cast<double2*>(acq_out + i) = make_double2(a0, a1);
    } else {
      const long long i = p;
      double acq = 0.0;
#pragma unroll
      for (int o = 0; o < MOBJ; ++o) {
        const double smu = (mu_in[o * ld + i] - hp.prior_mean[o]) / sd[o];
        const double svar = var_in[o * ld + i] / hp.prior_var[o];
        const double ucb = smu + hp.beta[o] * sqrt(fabs(svar));
        acq = acq + ucb;
        if (smu_out) smu_out[o * ld + i] = smu;
        if (svar_out) svar_out[o * ld + i] = svar;
        if (ucb_out) ucb_out[o * ld + i] = ucb;
      }
      if (acq_out) acq_out[i] = acq;
    }
  }
}

template <typename CT, int D>
int launch_kstar_m(int m, dim3 grid, cudaStream_t st, double* Kp, double* meandot, const CT* cand, int ldc,
                   long long cand0, long long n_cand, int chunk_tiles, long long ld_chunk, const double* x, int ldx,
                   int n, int npad, int d, const double* alpha, const ObjParams& hp) {
#define BO_KS(MO)                                                                                              \
  kstar_pack_kernel<CT, D, MO><<<grid, 128, 0, st>>>   (Kp, meandot, cand, ldc, cand0, n_cand, chunk_tiles,    \
                                                        ld_chunk, x, ldx, n, npad, d, alpha, hp)
  switch (m) {
    case 1: BO_KS(1); break;
    case 2: BO_KS(2); break;
    case 3: BO_KS(3); break;
    default: BO_KS(4); break;
  }
#undef BO_KS
  BO_LAUNCH_CHECK("kstar_pack_kernel");
  return BO_OK;
}

template <typename CT>
int launch_kstar(int m, int d, dim3 grid, cudaStream_t st, double* Kp, double* meandot, const CT* cand, int ldc,
                 long long cand0, long long n_cand, int chunk_tiles, long long ld_chunk, const double* x, int ldx,
                 int n, int npad, const double* alpha, const ObjParams& hp) {
#define BO_KD(DD)                                                                                                  \
  return launch_kstar_m<CT, DD>(m, grid, st, Kp, meandot, cand, ldc, cand0, n_cand, chunk_tiles, ld_chunk, x, ldx, \
                                n, npad, d, alpha, hp)
  if (d <= 2) BO_KD(2);
  if (d <= 4) BO_KD(4);
  if (d <= 6) BO_KD(6);
  if (d <= 8) BO_KD(8);
  if (d <= 10) BO_KD(10);
  if (d <= 12) BO_KD(12);
  BO_KD(16);
#undef BO_KD
}

}  // namespace

// =========================================================================================== host driver
namespace {

// Helper stream (lowest priority) + events used to generate K* for chunk i+1 while TRMM(i) runs on the
// caller's stream.  One set per device, created lazily; purely an execution resource (no data state).
struct Overlap {
  cudaStream_t aux = nullptr;
  cudaEvent_t inputs_ready = nullptr;
  cudaEvent_t kstar_done[2] = {nullptr, nullptr};
  cudaEvent_t buffer_free[2] = {nullptr, nullptr};
  bool ok = false;
};

Overlap* overlap_for_current_device() {
  static Overlap table[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  Overlap& o = table[dev];
  if (!o.ok) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);  // lo = least urgent
    if (cudaStreamCreateWithPriority(&o.aux, cudaStreamNonBlocking, lo) != cudaSuccess) return nullptr;
    bool good = cudaEventCreateWithFlags(&o.inputs_ready, cudaEventDisableTiming) == cudaSuccess;
    for (int b = 0; b < 2; ++b) {
      good = good && cudaEventCreateWithFlags(&o.kstar_done[b], cudaEventDisableTiming) == cudaSuccess;
      good = good && cudaEventCreateWithFlags(&o.buffer_free[b], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!good) return nullptr;
    o.ok = true;
  }
  return &o;
}

}  // namespace

int finalize_chunk(const ScoreOutputs& out, long long cand0, long long n_cand, const double* part,
                   const double* meandot, long long ld_chunk, int chunk_cands, int nb, int m, const ObjParams& hp,
                   double min_variance, cudaStream_t stream) {
  const int n_out = (out.mu != nullptr) + (out.var != nullptr) + (out.std_mu != nullptr) + (out.std_var != nullptr) +
                    (out.ucb != nullptr);
  // bytes per candidate: (nb + 1) m partials / mean dots read, n_out m + 1 doubles written
  ProfileScope prof_scope(stream, BO_PROF_FINALIZE,
                          (double)chunk_cands * 8.0 * ((double)(nb + 1) * m + (double)n_out * m + (out.acq ? 1 : 0)));
  const unsigned grid = (unsigned)((chunk_cands + 255) / 256);
  if (out.hvi.prepared && m == 2) {
    finalize_hvi_kernel<2><<<grid, 256, 0, stream>>>(out.mu, out.var, out.std_mu, out.std_var, out.ucb, out.acq, out.ld,
                                                     cand0, n_cand, part, meandot, ld_chunk, chunk_cands, nb, hp,
                                                     min_variance, out.hvi);
    BO_LAUNCH_CHECK("finalize_hvi_kernel");
  } else if (out.hvi.prepared && m == 3) {
    finalize_hvi_kernel<3><<<grid, 256, 0, stream>>>(out.mu, out.var, out.std_mu, out.std_var, out.ucb, out.acq, out.ld,
                                                     cand0, n_cand, part, meandot, ld_chunk, chunk_cands, nb, hp,
                                                     min_variance, out.hvi);
    BO_LAUNCH_CHECK("finalize_hvi_kernel");
  } else {
    finalize_kernel<<<grid, 256, 0, stream>>>(out.mu, out.var, out.std_mu, out.std_var, out.ucb, out.acq, out.ld, cand0,
                                              n_cand, part, meandot, ld_chunk, chunk_cands, nb, m, hp, min_variance);
    BO_LAUNCH_CHECK("finalize_kernel");
  }
  return BO_OK;
}

ScorePlan make_score_plan(int n, int m, long long n_cand) {
  ScorePlan p;
  p.npad = round_up(n, TM);
  p.nb = p.npad / TM;
  const long long tiles = (n_cand + TN - 1) / TN;
  // four CTAs' worth of tiles per SM per (objective, row pair): every TRMM wave is full and the launch
  // tail is amortised over 32 waves
  long long ct = 4LL * device_sm_count();
  // cap each of the two K* staging buffers near 1.5 GB
  const long long bytes_per_tile = (long long)m * p.npad * TN * sizeof(double);
  const long long cap = (3LL << 29) / bytes_per_tile;
  if (ct > cap) ct = cap < 1 ? 1 : cap;
  if (ct > tiles) ct = tiles;
  if (ct < 1) ct = 1;
  p.chunk_tiles = (int)ct;
  p.ld_chunk = ct * TN;
  p.nbuf = (tiles > ct) ? 2 : 1;  // double buffering only pays when there is more than one chunk
  p.kp_doubles = (size_t)m * ct * p.npad * TN;
  p.part_doubles = (size_t)m * p.nb * p.ld_chunk;
  p.mean_doubles = (size_t)m * p.ld_chunk;
  return p;
}

size_t score_workspace_bytes(const ScorePlan& p) {
  return p.nbuf * (align256(p.kp_doubles * 8) + align256(p.mean_doubles * 8)) + align256(p.part_doubles * 8);
}

int score_candidates(const ScoreOutputs& out, const void* cand, int cand_kind, int ldc, long long n_cand,
                     const double* x, int ldx, int n, int d, int m, const double* wpack, const double* alpha,
                     const ObjParams& hp, double min_variance, void* workspace, size_t workspace_bytes,
                     cudaStream_t stream) {
  if (n_cand <= 0) return BO_OK;
  const ScorePlan p = make_score_plan(n, m, n_cand);
  if (workspace_bytes < score_workspace_bytes(p)) {
    set_error("score workspace too small: %zu < %zu", workspace_bytes, score_workspace_bytes(p));
    return BO_ERR_WORKSPACE;
  }
  {
    const int rc_attr = ensure_dynamic_smem(trmm_sumsq_kernel, TR_SMEM);
    if (rc_attr) return rc_attr;
  }
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  double* Kp[2];
  double* meandot[2];
  size_t off = 0;
  for (int b = 0; b < p.nbuf; ++b) {
    Kp[b] = reinterpret_cast<double*>(ws + off);
    off += align256(p.kp_doubles * 8);
    meandot[b] = reinterpret_cast<double*>(ws + off);
    off += align256(p.mean_doubles * 8);
  }
  double* part = reinterpret_cast<double*>(ws + off);
  const long long strideWp = (long long)wpack_tile_offset(p.nb) * TILE_DOUBLES;
  const int npairs = (p.nb + 1) / 2;
  const long long n_chunks = (n_cand + p.ld_chunk - 1) / p.ld_chunk;

  // Generating K*(i+1) on the helper stream while TRMM(i) runs is OPT-IN (BO_SCORE_OVERLAP=1): DMMA and
  // DFMA share one FP64 datapath on B200, so there is nothing to hide behind, and co-resident K* CTAs
  // delay the launch of the 48K-register TRMM CTAs (measured: 71.4 ms overlapped vs 68.3 ms serial).
  static const bool want_overlap = [] {
    const char* e = getenv("BO_SCORE_OVERLAP");
    return e && e[0] == '1';
  }();
  Overlap* ov = (p.nbuf == 2 && want_overlap) ? overlap_for_current_device() : nullptr;
  cudaStream_t ks_stream = ov ? ov->aux : stream;
  if (ov) {
    BO_CUDA(cudaEventRecord(ov->inputs_ready, stream));  // factor / candidates produced on the caller's stream
    BO_CUDA(cudaStreamWaitEvent(ov->aux, ov->inputs_ready, 0));
  }

  auto chunk_tiles_of = [&](long long ci) {
    const long long cand0 = ci * p.ld_chunk;
    const long long remaining = n_cand - cand0;
    return (int)(((remaining < p.ld_chunk ? remaining : p.ld_chunk) + TN - 1) / TN);
  };
  auto launch_kstar_chunk = [&](long long ci) -> int {
    const int b = (int)(ci % p.nbuf);
    const long long cand0 = ci * p.ld_chunk;
    const int tiles = chunk_tiles_of(ci);
    if (ov && ci >= 2) BO_CUDA(cudaStreamWaitEvent(ov->aux, ov->buffer_free[b], 0));  // TRMM(ci-2) has drained it
    // algorithmic bytes of this launch: the K* staging it writes (8 m npad per candidate) + the candidates it reads
    ProfileScope prof_scope(ks_stream, BO_PROF_KSTAR, (double)tiles * TN * ((double)m * p.npad * 8.0 + 8.0 * d));
    int rc;
    if (cand_kind == BO_CAND_I64)
      rc = launch_kstar<long long>(m, d, dim3(2 * tiles), ks_stream, Kp[b], meandot[b],
                                   static_cast<const long long*>(cand), ldc, cand0, n_cand, p.chunk_tiles,
                                   p.ld_chunk, x, ldx, n, p.npad, alpha, hp);
    else
      rc = launch_kstar<double>(m, d, dim3(2 * tiles), ks_stream, Kp[b], meandot[b], static_cast<const double*>(cand),
                                ldc, cand0, n_cand, p.chunk_tiles, p.ld_chunk, x, ldx, n, p.npad, alpha, hp);
    if (rc) return rc;
    if (ov) BO_CUDA(cudaEventRecord(ov->kstar_done[b], ov->aux));
    return BO_OK;
  };

  // software pipeline: K*(0); for each chunk: K*(i+1) on the helper stream, then TRMM(i) + finalize(i)
  int rc = launch_kstar_chunk(0);
  if (rc) return rc;
  for (long long ci = 0; ci < n_chunks; ++ci) {
    const int b = (int)(ci % p.nbuf);
    const long long cand0 = ci * p.ld_chunk;
    const long long remaining = n_cand - cand0;
    const int tiles = chunk_tiles_of(ci);
    if (ov && ci + 1 < n_chunks) {
      rc = launch_kstar_chunk(ci + 1);
      if (rc) return rc;
    }
    if (ov) BO_CUDA(cudaStreamWaitEvent(stream, ov->kstar_done[b], 0));
    // grid: candidate tile fastest so that concurrently resident CTAs stream the same W tiles (L2 hits)
    const int units = m * npairs * tiles;
    const unsigned grid = (unsigned)(units < device_sm_count() ? units : device_sm_count());
    const bool prof = profile_enabled();
    if (prof) profile_begin(stream);
    // K* reuse in L2: groups of grid / npairs tiles, provided one objective's packed W and the group's K* tiles fit
    // in ~80 MB of the 126 MB L2 together; otherwise the plain order (BO_TRMM_GROUP=0 forces it)
    int group = tiles;
    {
      static const bool no_group = [] {
        const char* e = getenv("BO_TRMM_GROUP");
        return e && e[0] == '0';
      }();
      const long long tile_bytes = (long long)p.npad * TN * sizeof(double);
      const int g = (int)grid / npairs;
      if (!no_group && g >= 1 && g < tiles && strideWp * 8 + (long long)g * tile_bytes <= (80LL << 20)) group = g;
    }
    trmm_sumsq_kernel<<<grid, TR_THREADS, TR_SMEM, stream>>>(part, p.ld_chunk, wpack, strideWp, Kp[b], p.nb,
                                                             p.chunk_tiles, tiles, units, group);
    // algorithmic work of this launch: m * N^2 flops per live candidate
    if (prof) {
      const long long live = (remaining < p.ld_chunk ? remaining : p.ld_chunk);
      profile_end(stream, (double)live * m * (double)n * (double)n);
    }
    BO_LAUNCH_CHECK("trmm_sumsq_kernel");
    rc = finalize_chunk(out, cand0, n_cand, part, meandot[b], p.ld_chunk, tiles * TN, p.nb, m, hp, min_variance,
                        stream);
    if (rc) return rc;
    // recorded AFTER finalize: K*(ci+2) on the helper stream rewrites meandot[b], which finalize(ci) still reads
    if (ov) BO_CUDA(cudaEventRecord(ov->buffer_free[b], stream));
    if (!ov && ci + 1 < n_chunks) {
      rc = launch_kstar_chunk(ci + 1);
      if (rc) return rc;
    }
  }
  return BO_OK;
}

int acquisition_only(double* smu, double* svar, double* ucb, double* acq, const double* mu, const double* var,
                     long long ld, long long n_cand, int m, const ObjParams& hp, cudaStream_t stream) {
  if (n_cand <= 0) return BO_OK;
  auto aligned = [](const void* p) { return p == nullptr || (((uintptr_t)p) & 15) == 0; };
  const bool vec = (ld % 2 == 0) && (n_cand % 2 == 0) && aligned(smu) && aligned(svar) && aligned(ucb) &&
                   aligned(acq) && aligned(mu) && aligned(var);
  const long long items = vec ? n_cand / 2 : n_cand;
  long long blocks = (items + 255) / 256;
  const long long cap = 32LL * device_sm_count();
  if (blocks > cap) blocks = cap;
#define BO_ACQ(MO)                                                                                              \
  do {                                                                                                          \
    if (vec)                                                                                                    \
      acquisition_kernel<MO, true><<<(unsigned)blocks, 256, 0, stream>>>(smu, svar, ucb, acq, mu, var, ld, n_cand, hp); \
    else                                                                                                        \
      acquisition_kernel<MO, false><<<(unsigned)blocks, 256, 0, stream>>>(smu, svar, ucb, acq, mu, var, ld, n_cand, hp); \
  } while (0)
  switch (m) {
    case 1: BO_ACQ(1); break;
    case 2: BO_ACQ(2); break;
    case 3: BO_ACQ(3); break;
    default: BO_ACQ(4); break;
  }
#undef BO_ACQ
  BO_LAUNCH_CHECK("acquisition_kernel");
  return BO_OK;
}

}  // namespace bo
