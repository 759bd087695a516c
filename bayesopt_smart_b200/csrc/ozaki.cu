// ozaki.cu -- INT8 tensor-core engine for the posterior variance: sum_i (W k*)_i^2 by error-free splitting.
// Reference semantics: update_variance (numba_kernels.py:491-535) -- same quantity as score.cu's DMMA path,
// computed on tcgen05.mma.kind::i8 instead of DMMA.
//
// Number format.  Row i of W is scaled by 2^-e_i (|w| 2^-e_i <= 126/128) and rounded to a 48-bit integer
// q = rint(w 2^(47-e_i)); a K* entry kt = exp(..) in [0,1] becomes q = rint(kt 2^46).  q is written in BALANCED
// base-256 digits q = sum_s d_s 256^(5-s), d_s in [-128,127] (top digit of W in [-127,127], of K* in [0,65]):
// adding 0x80 to each of the five low bytes turns the balanced digits into the plain bytes of the sum, so the
// digits are the bytes of (q + 0x8080808080) with the low five XORed by 0x80.  With balanced digits the dropped
// pairs (s + t >= 6) are zero-mean, which is worth two digits of accuracy over unsigned slices.
//   V_i = pvar 2^(e_i-13) sum_{g=0..5} 256^-g acc_g,   acc_g = sum_{s+t=g} sum_k dW_s[i,k] dK_t[k,c]   (exact int32)
//       = pvar 2^(e_i-29) (b_0 + 2^-24 b_1),  b_j = (256 acc_3j + acc_3j+1) 256 + acc_3j+2          (exact int64)
//
// Layout.  Both operands are stored in HBM exactly as the UMMA canonical K-major / no-swizzle shared-memory
// image of one k-step (32 k), all six planes of a (row block | candidate tile) contiguous:
//   [plane][row group of 8][k half of 16][row in group][16 bytes]
// so one cp.async.bulk per operand per k-step fills a pipeline stage, and the smem descriptor is
// (start, LBO = 128 B between the two k halves, SBO = 256 B between row groups).
//
// Kernel (oz_sumsq_kernel).  One CTA of 11 warps per SM, persistent; a work unit is (candidate tile of 64,
// objective, row split) and the CTA walks all row blocks of the unit.
//   warp 0        bulk-copy producer, 5 stages of 36 KB (one k-step of W planes + K* planes)
//   warps 1, 10   the two MMA issuers, alternating k-steps: 21 tcgen05.mma.kind::i8 128x64x32 per k-step into
//                 6 TMEM accumulators (384 columns); warp 1 also allocates / frees TMEM
//   warps 2-9     during the k loop: A-operand feeders; at the end of a row block: epilogue, two warps per TMEM
//                 lane quarter with 32 columns each (tcgen05.ld, integer recombination in triples, two exact
//                 int64 -> FP64 conversions, row scale, square, running per-thread sums over all row blocks of
//                 the unit; one transposed shuffle reduction per unit)
// The W planes are the MMA's A operand and are read from TENSOR MEMORY: with A in shared memory a 128 x N x 32
// kind::i8 MMA takes N/2 + 43 cycles on B200 (the 4 KB A read is not hidden: tools/umma_probe.cu), from TMEM it
// takes N/2.  Per k-step the feeder warps move the six 4 KB A planes shared memory -> registers -> TMEM
// (ld.shared.v4 + tcgen05.st.32x32b.x8, two alternating 48-column slots).  tcgen05.cp from the issuing thread
// was measured and rejected: it shares the in-order pipe with the MMAs (939 vs 672 cycles per k-step,
// tools/umma_probe2.cu; the register path: 845 with the handshake, tools/umma_probe3.cu).
#include "ozaki.cuh"

#include <stdlib.h>

#include "rbf.cuh"

namespace bo {

namespace {

constexpr int OZ_STAGES = 5;
constexpr int OZ_THREADS = 352;     // producer warp, MMA warp, 8 feeder / epilogue warps, second MMA warp
constexpr int OZ_EC = OZ_TN / 2;     // accumulator columns per epilogue warp
constexpr int OZ_STAGE_BYTES = OZ_A_STAGE + OZ_B_STAGE;  // 36864
constexpr int OZ_ASLOT_COL = OZ_PLANES * OZ_TN;          // first TMEM column of the two A-operand slots
constexpr int OZ_ASLOT_COLS = OZ_PLANES * (OZ_KS / 4);   // 48 columns: six planes of 128 lanes x 32 bytes
constexpr int OZ_TMEM_COLS = 512;
constexpr size_t OZ_SMEM = (size_t)OZ_STAGES * OZ_STAGE_BYTES + 4 * OZ_TN * sizeof(double) +
                           (2 * OZ_STAGES + 6) * sizeof(uint64_t) + 16;

constexpr unsigned long long OZ_BIAS = 0x0000008080808080ull;  // 0x80 in each of the five low digit bytes
constexpr double OZ_MAGIC = 6755399441055744.0;                // 1.5 * 2^52

// bytes 0..5 of the result, with 0x80 XORed onto bytes 0..4, are the balanced digits (least significant first) of
// rint(v * scale); the XOR is applied to whole packed words by digit_planes (one instruction per four digits)
__device__ __forceinline__ unsigned long long biased_digits(double v, double scale) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(fma(v, scale, OZ_MAGIC));
  return bits + OZ_BIAS;
}

// 16 biased digit words -> the six 16-byte plane rows (rows[p] = plane p, most significant first; byte b of a row
// belongs to word b).  Per group of four words a 4x4 byte transpose of the low halves (8 byte-permutes) and a 2x4
// one of the high halves (4), then the bias is removed from the five low planes with one XOR per packed word.
__device__ __forceinline__ void digit_planes(const unsigned long long (&w)[16], uint4 (&rows)[OZ_PLANES]) {
  uint32_t pw[OZ_PLANES][4];  // [byte index 0..5][group]
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const uint32_t a = (uint32_t)w[4 * g], b = (uint32_t)w[4 * g + 1], c = (uint32_t)w[4 * g + 2],
                   d = (uint32_t)w[4 * g + 3];
    const uint32_t ab_lo = __byte_perm(a, b, 0x5140), ab_hi = __byte_perm(a, b, 0x7362);
    const uint32_t cd_lo = __byte_perm(c, d, 0x5140), cd_hi = __byte_perm(c, d, 0x7362);
    pw[0][g] = __byte_perm(ab_lo, cd_lo, 0x5410) ^ 0x80808080u;
    pw[1][g] = __byte_perm(ab_lo, cd_lo, 0x7632) ^ 0x80808080u;
    pw[2][g] = __byte_perm(ab_hi, cd_hi, 0x5410) ^ 0x80808080u;
    pw[3][g] = __byte_perm(ab_hi, cd_hi, 0x7632) ^ 0x80808080u;
    const uint32_t ah = (uint32_t)(w[4 * g] >> 32), bh = (uint32_t)(w[4 * g + 1] >> 32),
                   ch = (uint32_t)(w[4 * g + 2] >> 32), dh = (uint32_t)(w[4 * g + 3] >> 32);
    const uint32_t abh = __byte_perm(ah, bh, 0x5140), cdh = __byte_perm(ch, dh, 0x5140);
    pw[4][g] = __byte_perm(abh, cdh, 0x5410) ^ 0x80808080u;
    pw[5][g] = __byte_perm(abh, cdh, 0x7632);  // the top digit carries no bias
  }
#pragma unroll
  for (int p = 0; p < OZ_PLANES; ++p)
    rows[p] = make_uint4(pw[OZ_PLANES - 1 - p][0], pw[OZ_PLANES - 1 - p][1], pw[OZ_PLANES - 1 - p][2],
                         pw[OZ_PLANES - 1 - p][3]);
}

// ------------------------------------------------------------------------------------------- W digits
// element (r, k) of the DMMA-packed W (common.cuh / factor.cu pack_w)
__device__ __forceinline__ long long wpack_index(int r, int k) {
  const int ib = r >> 7, rr = r & 127, kt = k >> 4, kk = k & 15;
  const int wm = rr >> 6, i = (rr >> 3) & 7, g = rr & 7;
  const int sp = kk >> 3, q = (kk >> 2) & 1, t = kk & 3;
  return (wpack_tile_offset(ib) + kt) * TILE_DOUBLES + (((wm * 8 + i) * 2 + sp) * 32 + (4 * g + t)) * 2 + q;
}

// one warp per (objective, row): e_i with max|w| * 128/126 <= 2^e_i
__global__ void oz_rowscale_kernel(double* __restrict__ wscale, double* __restrict__ qscale,
                                   const double* __restrict__ wpack, long long strideWp, int n, int npad) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int o = blockIdx.y, lane = threadIdx.x & 31;
  if (row >= npad) return;
  double mx = 0.0;
  if (row < n) {
    const double* W = wpack + (long long)o * strideWp;
    for (int k = lane; k <= row; k += 32) mx = fmax(mx, fabs(W[wpack_index(row, k)]));
  }
#pragma unroll
  for (int s = 16; s; s >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, s));
  if (lane == 0) {
    double ws = 0.0, qs = 0.0;
    if (mx > 0.0 && mx < 1e300) {
      int e;
      frexp(mx * (128.0 / 126.0), &e);  // value = f 2^e, f in [0.5, 1)  =>  value <= 2^e
      ws = ldexp(1.0, e - 29);  // 2^(e-13) for sum_g 256^-g acc_g, times 2^-16 for the recombination in triples
      qs = ldexp(1.0, 47 - e);
    }
    wscale[(long long)o * npad + row] = ws;
    qscale[(long long)o * npad + row] = qs;
  }
}

// one thread per 16-byte row of one k-step block: (o, ib, ks, rg, kc, r) -> six plane rows
__global__ void oz_wdigits_kernel(unsigned char* __restrict__ wq, long long strideWq,
                                  const double* __restrict__ qscale, const double* __restrict__ wpack,
                                  long long strideWp, int n, int npad, int nb) {
  const long long blocks = 2LL * nb * (nb + 1);  // k-step blocks per objective
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int o = blockIdx.y;
  if (tid >= blocks * 256) return;
  const long long blk = tid >> 8;
  const int within = (int)(tid & 255);
  const int rg = within >> 4, kc = (within >> 3) & 1, r = within & 7;
  // blk = 2 ib (ib + 1) + ks
  int ib = (int)((sqrt(1.0 + 2.0 * (double)blk) - 1.0) * 0.5);
  while (2LL * ib * (ib + 1) > blk) --ib;
  while (2LL * (ib + 1) * (ib + 2) <= blk) ++ib;
  const int ks = (int)(blk - 2LL * ib * (ib + 1));
  const int row = ib * OZ_TM + rg * 8 + r;
  const int k0 = ks * OZ_KS + kc * 16;
  const double* W = wpack + (long long)o * strideWp;
  const double qs = qscale[(long long)o * npad + row];
  unsigned long long dg[16];
#pragma unroll
  for (int b = 0; b < 16; ++b) {
    const int k = k0 + b;
    const double w = (row < n && k <= row) ? W[wpack_index(row, k)] : 0.0;
    dg[b] = biased_digits(w, qs);
  }
  unsigned char* dst = wq + (long long)o * strideWq + blk * OZ_A_STAGE + rg * 256 + kc * 128 + r * 16;
  uint4 rows[OZ_PLANES];
  digit_planes(dg, rows);
#pragma unroll
  for (int p = 0; p < OZ_PLANES; ++p) *reinterpret_cast<uint4*>(dst + p * OZ_A_PLANE) = rows[p];
}

// ------------------------------------------------------------------------------------------- K* digits
// One CTA (128 threads) per candidate tile of 64: thread = (candidate cl = tid % 64, k half kc = tid / 64).
// Per k-step the thread evaluates its 16 kernel entries per objective, turns them into digits and writes the
// six 16-byte plane rows; the posterior-mean dot product k*.alpha is accumulated from the unquantised values.
constexpr int OZK_ROWS = 256;
constexpr int OZK_THREADS = 2 * OZ_TN;
constexpr int OZK_MIN_CTAS = 3;  // 168 registers: 2 and 4 CTAs per SM were measured 0-8 % slower
// row of the staged training block: D coordinates (D is even) then MOBJ alphas, padded to whole 16-byte pairs --
// every lane of a warp reads the same row (broadcast), so the rows can be read with 16-byte loads
__host__ __device__ constexpr int ozk_row_doubles(int d, int m) { return (d + m + 1) & ~1; }

template <typename CT, int D, int MOBJ>
__global__ void __launch_bounds__(OZK_THREADS, OZK_MIN_CTAS)
    oz_kstar_digits_kernel(unsigned char* __restrict__ kq, double* __restrict__ meandot, const CT* __restrict__ cand,
                           int ldc, long long cand0, long long n_cand, int chunk_tiles, long long ld_chunk,
                           const double* __restrict__ x, int ldx, int n, int npad, int d,
                           const double* __restrict__ alpha, int alpha_ld, ObjParams hp) {
  __shared__ double exp_tab[64];
  constexpr int RS = ozk_row_doubles(D, MOBJ);
  __shared__ __align__(16) double xs[OZK_ROWS][RS];
  __shared__ double mred[MOBJ][OZ_TN];
  const int tid = threadIdx.x;
  if (tid < 64) exp_tab[tid] = kExp2Tab[tid];
  const int ct = blockIdx.x;
  const int cl = tid % OZ_TN, kc = tid / OZ_TN;
  const int rg = cl >> 3, r = cl & 7;
  const int nk_tot = npad / OZ_KS;

  double cc[D];
  {
    long long ci = cand0 + (long long)ct * OZ_TN + cl;
    if (ci >= n_cand) ci = n_cand - 1;  // tail tile: computed, never stored by finalize
#pragma unroll
    for (int k = 0; k < D; ++k) cc[k] = (k < d) ? (double)cand[ci * ldc + k] : 0.0;
  }
  double macc[MOBJ], coef[MOBJ], pvar[MOBJ];
#pragma unroll
  for (int o = 0; o < MOBJ; ++o) {
    macc[o] = 0.0;
    coef[o] = hp.neg_half_inv_ls2[o];
    pvar[o] = hp.prior_var[o];
  }
  const double kscale = 70368744177664.0;  // 2^46

  for (int r0 = 0; r0 < npad; r0 += OZK_ROWS) {
    __syncthreads();
    {
      // all loads of this thread first, then the stores: the trip count is a compile-time constant, so the global
      // loads are in flight together instead of one load -> store round trip per element
      constexpr int PER_THREAD = OZK_ROWS * RS / OZK_THREADS;
      double stage[PER_THREAD];
#pragma unroll
      for (int i = 0; i < PER_THREAD; ++i) {
        const int e = tid + i * OZK_THREADS;
        const int rr = e / RS, k = e - rr * RS;
        const int row = r0 + rr;
        double v = 0.0;
        if (row < npad) {
          const int rc = row < n ? row : n - 1;  // padded rows reuse the last real point (finite values)
          if (k < D) v = (k < d) ? x[(long long)rc * ldx + k] : 0.0;
          else if (k < D + MOBJ) v = (row < n) ? alpha[(long long)(k - D) * alpha_ld + row] : 0.0;
        }
        stage[i] = v;
      }
#pragma unroll
      for (int i = 0; i < PER_THREAD; ++i) {
        const int e = tid + i * OZK_THREADS;
        xs[e / RS][e % RS] = stage[i];
      }
    }
    __syncthreads();
    const int ks_end = min(nk_tot, (r0 + OZK_ROWS) / OZ_KS);
    for (int ks = r0 / OZ_KS; ks < ks_end; ++ks) {
      const int kb = ks * OZ_KS + kc * 16 - r0;
      double sq[16];
#pragma unroll
      for (int b = 0; b < 16; ++b) {
        const double2* xr = reinterpret_cast<const double2*>(xs[kb + b]);
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; k += 2) {  // same order of accumulation as the scalar loop: k = 0..D-1
          const double2 xv = xr[k >> 1];
          const double d0 = xv.x - cc[k], d1 = xv.y - cc[k + 1];
          s = fma(d0, d0, s);
          s = fma(d1, d1, s);
        }
        sq[b] = s;
      }
#pragma unroll
      for (int o = 0; o < MOBJ; ++o) {
        unsigned long long dg[16];
#pragma unroll
        for (int b = 0; b < 16; ++b) {
          const double e = rbf_exp<false>(sq[b] * coef[o], exp_tab);
          macc[o] = fma(pvar[o] * e, xs[kb + b][D + o], macc[o]);
          dg[b] = biased_digits(e, kscale);
        }
        unsigned char* dst = kq + (((long long)o * chunk_tiles + ct) * nk_tot + ks) * OZ_B_STAGE + rg * 256 +
                             kc * 128 + r * 16;
        uint4 rows[OZ_PLANES];
        digit_planes(dg, rows);
#pragma unroll
        for (int p = 0; p < OZ_PLANES; ++p) *reinterpret_cast<uint4*>(dst + p * OZ_B_PLANE) = rows[p];
      }
    }
  }
  // mean: the two k halves of a candidate, fixed order
  if (kc == 1) {
#pragma unroll
    for (int o = 0; o < MOBJ; ++o) mred[o][cl] = macc[o];
  }
  __syncthreads();
  if (kc == 0) {
#pragma unroll
    for (int o = 0; o < MOBJ; ++o)
      meandot[(long long)o * ld_chunk + (long long)ct * OZ_TN + cl] = macc[o] + mred[o][cl];
  }
}

// ------------------------------------------------------------------------------------------- tcgen05 helpers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1), K-major, no swizzle: 8-row x 16-byte
// core matrices; LBO = 128 B between the two k halves of one MMA, SBO = 256 B between consecutive 8-row groups.
// Built from a precomputed low word ((address >> 4) | LBO << 16); the high word is constant.
__device__ __forceinline__ uint64_t umma_desc_lo(uint32_t lo) {
  return (uint64_t)lo | ((uint64_t)((256u >> 4) | (1u << 14)) << 32);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = S32, A = B = signed 8 bit, both K-major, N, M
constexpr uint32_t OZ_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_TN >> 3) << 17) |
                              ((uint32_t)(OZ_TM >> 4) << 24);

// A operand (128 lanes x 8 columns = 128 rows x 32 bytes) from tensor memory
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(OZ_IDESC), "r"(accumulate)
      : "memory");
}

// arrives on the mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, int (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr));
}
// the loaded registers are operands of the wait, so no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait(int (&a)[OZ_PLANES][4]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0][0]), "+r"(a[0][1]), "+r"(a[0][2]), "+r"(a[0][3]), "+r"(a[1][0]), "+r"(a[1][1]),
                 "+r"(a[1][2]), "+r"(a[1][3]), "+r"(a[2][0]), "+r"(a[2][1]), "+r"(a[2][2]), "+r"(a[2][3]),
                 "+r"(a[3][0]), "+r"(a[3][1]), "+r"(a[3][2]), "+r"(a[3][3]), "+r"(a[4][0]), "+r"(a[4][1]),
                 "+r"(a[4][2]), "+r"(a[4][3]), "+r"(a[5][0]), "+r"(a[5][1]), "+r"(a[5][2]), "+r"(a[5][3])
               :
               : "memory");
}

// exact conversion of (a*256 + b)*256 + c (|.| < 2^48) to FP64: build the integer next to 1.5*2^52 and subtract
__device__ __forceinline__ double triple_to_double(int a, int b, int c) {
  const long long v = ((long long)a * 256 + (long long)b) * 256 + (long long)c;
  return __longlong_as_double(v + 0x4338000000000000ll) - OZ_MAGIC;
}

// row blocks of one candidate tile are dealt to `nsplit` CTAs in serpentine order (equal k-step totals +-1 block)
__device__ __forceinline__ int oz_row_block(int t, int r, int nsplit) {
  return t * nsplit + ((t & 1) ? (nsplit - 1 - r) : r);
}

// ------------------------------------------------------------------------------------------- the MMA kernel
// position of a role in the persistent schedule: work unit u -> its row blocks (serpentine order) -> k-steps
struct OzCursor {
  int u, t, ib, nk, ks;
  bool valid;
};
__device__ __forceinline__ void oz_cursor_init(OzCursor& c, int first_unit, int total_units, int nsplit) {
  c.u = first_unit;
  c.t = 0;
  c.ks = 0;
  c.valid = first_unit < total_units;
  c.ib = c.valid ? oz_row_block(0, c.u % nsplit, nsplit) : 0;
  c.nk = 4 * (c.ib + 1);
}
// next row block (returns true when that crosses into a new work unit)
__device__ __forceinline__ bool oz_cursor_next_block(OzCursor& c, int stride, int total_units, int nsplit, int nb) {
  bool new_unit = false;
  ++c.t;
  int ib = oz_row_block(c.t, c.u % nsplit, nsplit);
  if (ib >= nb) {
    c.u += stride;
    c.t = 0;
    new_unit = true;
    c.valid = c.u < total_units;
    ib = c.valid ? oz_row_block(0, c.u % nsplit, nsplit) : 0;
  }
  c.ib = ib;
  c.nk = 4 * (ib + 1);
  c.ks = 0;
  return new_unit;
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint4& lo, const uint4& hi) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(lo.x),
               "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
               : "memory");
}

// (A variant in which clusters of 2 or 4 CTAs multicast the W stages into each other's shared memory was measured:
// no gain -- the pass is not L2-bound -- and 4-CTA clusters lose SMs to GPC fragmentation.  It is not kept.)
__global__ void __launch_bounds__(OZ_THREADS, 1)
    oz_sumsq_kernel(double* __restrict__ part, long long ld_chunk, const unsigned char* __restrict__ wq,
                    long long strideWq, const double* __restrict__ wscale, const unsigned char* __restrict__ kq,
                    int npad, int nb, int nk_tot, int chunk_tiles, int nsplit, int m, int total_units,
                    ObjParams hp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* sA = smem_raw;                              // [stage][plane][4096]
  unsigned char* sB = smem_raw + OZ_STAGES * OZ_A_STAGE;     // [stage][plane][2048]
  double* red = reinterpret_cast<double*>(smem_raw + (size_t)OZ_STAGES * OZ_STAGE_BYTES);  // [4][64]
  uint64_t* full = reinterpret_cast<uint64_t*>(red + 4 * OZ_TN);
  uint64_t* empty = full + OZ_STAGES;
  uint64_t* a_ready = empty + OZ_STAGES;  // [2]: the A slot holds the planes of its k-step
  uint64_t* turn = a_ready + 2;  // [2]: issue token of the two MMA warps
  uint64_t* tmem_full = turn + 2;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < OZ_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&a_ready[0], 8);
    mbar_init(&a_ready[1], 8);
    mbar_init(&turn[0], 1);
    mbar_init(&turn[1], 1);
    mbar_init(tmem_full, 2);  // one commit from each MMA warp (their last k-steps of the row block)
    mbar_init(tmem_empty, 8);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)OZ_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int first_unit = blockIdx.x, unit_stride = gridDim.x;

  if (warp == 0) {
    // ===== producer: one lane streams k-step stages with the bulk-copy engine =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      OzCursor c;
      for (oz_cursor_init(c, first_unit, total_units, nsplit); c.valid;
           oz_cursor_next_block(c, unit_stride, total_units, nsplit, nb)) {
        const int o = (c.u / nsplit) % m;
        const int ct = c.u / (nsplit * m);
        const unsigned char* At = wq + (long long)o * strideWq + 2LL * c.ib * (c.ib + 1) * OZ_A_STAGE;
        const unsigned char* Ko = kq + ((long long)o * chunk_tiles + ct) * nk_tot * OZ_B_STAGE;
        for (int ks = 0; ks < c.nk; ++ks) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], OZ_STAGE_BYTES);
          bulk_g2s(sA + stage * OZ_A_STAGE, At + (long long)ks * OZ_A_STAGE, OZ_A_STAGE, &full[stage]);
          bulk_g2s(sB + stage * OZ_B_STAGE, Ko + (long long)ks * OZ_B_STAGE, OZ_B_STAGE, &full[stage]);
          if (++stage == OZ_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ===== two MMA issuers: warp 1 takes the even k-steps (A slot 0), warp 10 the odd ones (slot 1).  The
    // tensor pipe buffers about one instruction, so whatever an issuing thread does between two batches (the
    // a_ready wait, the fence, the commit: ~170 cycles) is a bubble -- unless the other issuer's batch is running
    // meanwhile.  Issue is handed back and forth by a token after 17 of the 21 MMAs of a batch.  Correctness does
    // not depend on how the pipe orders the two threads' MMAs: the only non-commutative MMAs are the six that
    // overwrite (accumulate = 0 on the first k-step of a row block, always warp 1 because nk is a multiple of 4)
    // and warp 10 waits for THEIR completion before its first batch of the block; integer accumulation commutes;
    // and both issuers commit to tmem_full (count 2), each for its own last k-step of the block.
    // The whole warp walks the schedule (warp-uniform control flow keeps the descriptors in uniform registers);
    // one elected lane issues. =====
    const uint32_t me = warp == 1 ? 0u : 1u;
    int stage = 0;
    uint32_t acc_phase = 0, g = 0;  // g: k-steps of the schedule so far
    const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | ((128u >> 4) << 16);
    const uint32_t a_tm = tmem_base + OZ_ASLOT_COL + me * OZ_ASLOT_COLS;
    OzCursor c;
    for (oz_cursor_init(c, first_unit, total_units, nsplit); c.valid;
         oz_cursor_next_block(c, unit_stride, total_units, nsplit, nb)) {
      if (me == 0) mbar_wait(tmem_empty, acc_phase ^ 1);  // epilogue has drained the accumulators
      for (int ks = 0; ks < c.nk; ++ks, ++g) {
        if ((g & 1u) == me) {
          mbar_wait(&a_ready[me], (g >> 1) & 1);  // feeders saw full[stage] and filled this A slot
          tc_fence_after();
          if (g > 0) mbar_wait(&turn[me], ((g - 1) >> 1) & 1);  // the other issuer is 17 MMAs into k-step g-1
          if (ks == 1) {
            // second k-step of a row block: the overwriting MMAs of the first one (other issuer) must have
            // COMPLETED, not merely been issued -- from here on every MMA of the block accumulates, and
            // accumulation commutes, so nothing below depends on how the pipe orders the two threads' MMAs
            mbar_wait(&empty[(g - 1) % OZ_STAGES], ((g - 1) / OZ_STAGES) & 1);
            tc_fence_after();
          }
          if (elect_one()) {
            const uint32_t b_lo = b_lo0 + stage * (OZ_B_STAGE >> 4);
            const uint32_t first = ks > 0 ? 1u : 0u;
#pragma unroll
            for (int gg = 0; gg < OZ_PLANES; ++gg)  // s = 0: the MMAs that start an accumulator
              umma_i8_ts(tmem_base + gg * OZ_TN, a_tm, umma_desc_lo(b_lo + gg * (OZ_B_PLANE >> 4)), first);
            int cnt = OZ_PLANES;
#pragma unroll
            for (int gg = 1; gg < OZ_PLANES; ++gg) {
#pragma unroll
              for (int sp = 1; sp <= gg; ++sp) {
                umma_i8_ts(tmem_base + gg * OZ_TN, a_tm + sp * (OZ_KS / 4),
                           umma_desc_lo(b_lo + (gg - sp) * (OZ_B_PLANE >> 4)), 1u);
                if (++cnt == 17) mbar_arrive(&turn[me ^ 1u]);
              }
            }
            umma_commit(&empty[stage]);  // frees the smem stage and this A slot once these MMAs have read them
            // each issuer reports the completion of its own last k-step of the row block
            if (ks >= c.nk - 2) umma_commit(tmem_full);
          }
          __syncwarp();
        }
        if (++stage == OZ_STAGES) stage = 0;
      }
      acc_phase ^= 1;
    }
  } else {
    // ===== warps 2..9: A-operand feeders during the k loop, epilogue at the end of every row block.
    // TMEM lane quarter = warp % 4 (row = quarter*32 + lane); half = (warp - 2) / 4 picks planes 3*half..3*half+2
    // when feeding and accumulator columns half*32..half*32+31 when draining. =====
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t acc_addr = lane_base + half * OZ_EC;
    const int row = quarter * 32 + lane;
    const unsigned char* a_src = sA + (half * 3) * OZ_A_PLANE + (row >> 3) * 256 + (row & 7) * 16;
    const int et = threadIdx.x - 64;  // 0..255 among these threads
    uint32_t acc_phase = 0;
    uint32_t fed = 0;        // k-steps fed so far (global count: stage = fed % STAGES, slot = fed & 1)
    uint32_t drain_end = 0;  // global index one past the last k-step of the row block being drained
    OzCursor fc, dc;         // feeding runs up to two k-steps ahead of draining
    oz_cursor_init(fc, first_unit, total_units, nsplit);
    oz_cursor_init(dc, first_unit, total_units, nsplit);
    double ssum[OZ_EC];
#pragma unroll
    for (int c = 0; c < OZ_EC; ++c) ssum[c] = 0.0;
    while (dc.valid) {
      drain_end += dc.nk;
      // ---- feed every k-step of this row block and the first two of the next (their slots are free as soon as
      // the last MMAs of this block have completed, i.e. while the accumulators are being drained)
      while (fc.valid && fed < drain_end + 2) {
        const int stage = fed % OZ_STAGES;
        mbar_wait(&full[stage], (fed / OZ_STAGES) & 1);  // TMA has landed the stage
        // the planes go to registers first: only the store into the A slot has to wait for the slot
        const unsigned char* src = a_src + stage * OZ_A_STAGE;
        uint4 lo[3], hi[3];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          lo[p] = *reinterpret_cast<const uint4*>(src + p * OZ_A_PLANE);        // k 0..15 of the row
          hi[p] = *reinterpret_cast<const uint4*>(src + p * OZ_A_PLANE + 128);  // k 16..31
        }
        if (fed >= 2) {
          const uint32_t j = fed - 2;  // the k-step that used this A slot: its MMAs signal empty[its stage]
          mbar_wait(&empty[j % OZ_STAGES], (j / OZ_STAGES) & 1);
          tc_fence_after();
        }
        const uint32_t dst = lane_base + OZ_ASLOT_COL + (fed & 1) * OZ_ASLOT_COLS + (half * 3) * (OZ_KS / 4);
#pragma unroll
        for (int p = 0; p < 3; ++p) tmem_st8(dst + p * (OZ_KS / 4), lo[p], hi[p]);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[fed & 1]);
        ++fed;
        if (++fc.ks == fc.nk) oz_cursor_next_block(fc, unit_stride, total_units, nsplit, nb);
      }
      // ---- drain the accumulators of row block dc
      const int r = dc.u % nsplit;
      const int o = (dc.u / nsplit) % m;
      const int ct = dc.u / (nsplit * m);
      const double f = wscale[(long long)o * npad + dc.ib * OZ_TM + row];
      const double f24 = f * (1.0 / 16777216.0);
      mbar_wait(tmem_full, acc_phase);
      tc_fence_after();
      // software pipeline over chunks of 4 columns: the tcgen05.ld of chunk c+1 is in flight while chunk c is
      // recombined, and the accumulators are handed back to the MMA warps as soon as the last load has landed
      // (before the arithmetic of the last chunk)
      int a[2][OZ_PLANES][4];
#pragma unroll
      for (int gq = 0; gq < OZ_PLANES; ++gq) tmem_ld4(acc_addr + gq * OZ_TN, a[0][gq]);
#pragma unroll
      for (int c0 = 0; c0 < OZ_EC; c0 += 4) {
        const int cur = (c0 >> 2) & 1;
        tmem_ld_wait(a[cur]);
        if (c0 + 4 < OZ_EC) {
#pragma unroll
          for (int gq = 0; gq < OZ_PLANES; ++gq) tmem_ld4(acc_addr + gq * OZ_TN + c0 + 4, a[cur ^ 1][gq]);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // V = f (b0 + 2^-24 b1),  b0 = (a0 256 + a1) 256 + a2,  b1 = (a3 256 + a4) 256 + a5   (exact int64)
          const double b0 = triple_to_double(a[cur][0][j], a[cur][1][j], a[cur][2][j]);
          const double b1 = triple_to_double(a[cur][3][j], a[cur][4][j], a[cur][5][j]);
          const double v = fma(b1, f24, b0 * f);
          ssum[c0 + j] = fma(v, v, ssum[c0 + j]);
        }
      }
      acc_phase ^= 1;
      const bool unit_done = oz_cursor_next_block(dc, unit_stride, total_units, nsplit, nb);
      if (unit_done) {
        // sum over the 32 rows of this warp: transposed butterfly, 32 -> 16 -> 8 -> 4 -> 2 -> 1 values per lane;
        // lane L ends up with the sum of column half*32 + L
#pragma unroll
        for (int w = OZ_EC / 2; w >= 1; w >>= 1) {
#pragma unroll
          for (int c = 0; c < w; ++c) {
            const bool up = lane & w;
            const double send = up ? ssum[c] : ssum[c + w];
            const double keep = up ? ssum[c + w] : ssum[c];
            ssum[c] = keep + __shfl_xor_sync(0xffffffffu, send, w);
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // previous unit's readers of `red` are done
        red[quarter * OZ_TN + half * OZ_EC + lane] = ssum[0];
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et < OZ_TN) {
          const double q = ((red[et] + red[OZ_TN + et]) + red[2 * OZ_TN + et]) + red[3 * OZ_TN + et];
          const double pv = hp.prior_var[o];
          part[((long long)o * nsplit + r) * ld_chunk + (long long)ct * OZ_TN + et] = q * pv * pv;
        }
#pragma unroll
        for (int c = 0; c < OZ_EC; ++c) ssum[c] = 0.0;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)OZ_TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------- INT8 peak probe
// The roofline denominator of oz_sumsq_kernel, measured on the chip like cuBLAS DGEMM is for the DMMA path: one
// CTA per SM issues nothing but the kernel's own MMA batch (21 x tcgen05.mma.kind::i8 128x64x32, A from TMEM, B
// from shared memory, 6 accumulators) back to back on resident operands -- no loads, no handshakes, no epilogue.
__global__ void __launch_bounds__(64, 1) oz_peak_probe_kernel(int iters) {
  __shared__ __align__(128) unsigned char sBp[OZ_B_STAGE];
  __shared__ uint64_t done;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < OZ_B_STAGE; i += blockDim.x) sBp[i] = (unsigned char)((i * 2654435761u) >> 13);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&done, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)),
                 "r"((uint32_t)OZ_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> async proxy (MMA)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t b_lo = ((smem_u32(sBp) & 0x3FFFFu) >> 4) | ((128u >> 4) << 16);
      for (int it = 0; it < iters; ++it) {
        const uint32_t a_tm = tm + OZ_ASLOT_COL + (it & 1) * OZ_ASLOT_COLS;  // operand values are irrelevant
#pragma unroll
        for (int g = 0; g < OZ_PLANES; ++g) {
#pragma unroll
          for (int sp = 0; sp <= g; ++sp)
            umma_i8_ts(tm + g * OZ_TN, a_tm + sp * (OZ_KS / 4), umma_desc_lo(b_lo + (g - sp) * (OZ_B_PLANE >> 4)),
                       (it | sp) ? 1u : 0u);
        }
      }
      umma_commit(&done);
      mbar_wait(&done, 0);
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"((uint32_t)OZ_TMEM_COLS)
                 : "memory");
  }
}

template <typename CT, int D>
int launch_oz_kstar_m(int m, dim3 grid, cudaStream_t st, unsigned char* kq, double* meandot, const CT* cand, int ldc,
                      long long cand0, long long n_cand, int chunk_tiles, long long ld_chunk, const double* x,
                      int ldx, int n, int npad, int d, const double* alpha, int alpha_ld, const ObjParams& hp) {
#define BO_OZK(MO)                                                                                                 \
  oz_kstar_digits_kernel<CT, D, MO><<<grid, OZK_THREADS, 0, st>>>(kq, meandot, cand, ldc, cand0, n_cand,           \
                                                                   chunk_tiles, ld_chunk, x, ldx, n, npad, d, alpha, \
                                                                   alpha_ld, hp)
  switch (m) {
    case 1: BO_OZK(1); break;
    case 2: BO_OZK(2); break;
    case 3: BO_OZK(3); break;
    default: BO_OZK(4); break;
  }
#undef BO_OZK
  BO_LAUNCH_CHECK("oz_kstar_digits_kernel");
  return BO_OK;
}

template <typename CT>
int launch_oz_kstar(int m, int d, dim3 grid, cudaStream_t st, unsigned char* kq, double* meandot, const CT* cand,
                    int ldc, long long cand0, long long n_cand, int chunk_tiles, long long ld_chunk, const double* x,
                    int ldx, int n, int npad, const double* alpha, int alpha_ld, const ObjParams& hp) {
#define BO_OZD(DD)                                                                                                 \
  return launch_oz_kstar_m<CT, DD>(m, grid, st, kq, meandot, cand, ldc, cand0, n_cand, chunk_tiles, ld_chunk, x,   \
                                   ldx, n, npad, d, alpha, alpha_ld, hp)
  if (d <= 2) BO_OZD(2);
  if (d <= 4) BO_OZD(4);
  if (d <= 6) BO_OZD(6);
  if (d <= 8) BO_OZD(8);
  if (d <= 10) BO_OZD(10);
  if (d <= 12) BO_OZD(12);
  BO_OZD(16);
#undef BO_OZD
}

}  // namespace

// =========================================================================================== host side
size_t oz_wq_bytes(int n) {
  const long long nb = round_up(n, OZ_TM) / OZ_TM;
  return (size_t)(2 * nb * (nb + 1)) * OZ_A_STAGE;
}

int oz_quantize_w(unsigned char* wq, double* wscale, const double* wpack, int n, int m, cudaStream_t st) {
  const int npad = round_up(n, OZ_TM), nb = npad / OZ_TM;
  const long long strideWp = (long long)wpack_tile_offset(nb) * TILE_DOUBLES;
  // the quantisation scale 2^(47 - e_i) lives in the second half of the caller's wscale array
  double* qscale = wscale + (long long)m * npad;
  oz_rowscale_kernel<<<dim3((npad + 7) / 8, m), 256, 0, st>>>(wscale, qscale, wpack, strideWp, n, npad);
  BO_LAUNCH_CHECK("oz_rowscale_kernel");
  const long long threads = 2LL * nb * (nb + 1) * 256;
  oz_wdigits_kernel<<<dim3((unsigned)((threads + 255) / 256), m), 256, 0, st>>>(wq, (long long)oz_wq_bytes(n), qscale,
                                                                                 wpack, strideWp, n, npad, nb);
  BO_LAUNCH_CHECK("oz_wdigits_kernel");
  return BO_OK;
}

int oz_kstar_digits(unsigned char* kq, double* meandot, const void* cand, int cand_kind, int ldc, long long cand0,
                    long long n_cand, int tiles, int chunk_tiles, const double* x, int ldx, int n, int d, int m,
                    const double* alpha, const ObjParams& hp, cudaStream_t st) {
  const int npad = round_up(n, OZ_TM);
  const long long ld_chunk = (long long)chunk_tiles * OZ_TN;
  if (cand_kind == BO_CAND_I64)
    return launch_oz_kstar<long long>(m, d, dim3(tiles), st, kq, meandot, static_cast<const long long*>(cand), ldc,
                                      cand0, n_cand, chunk_tiles, ld_chunk, x, ldx, n, npad, alpha, npad, hp);
  return launch_oz_kstar<double>(m, d, dim3(tiles), st, kq, meandot, static_cast<const double*>(cand), ldc, cand0,
                                 n_cand, chunk_tiles, ld_chunk, x, ldx, n, npad, alpha, npad, hp);
}

int oz_sumsq(double* part, long long ld_chunk, const unsigned char* wq, const double* wscale,
             const unsigned char* kq, int n, int m, int tiles, int chunk_tiles, int nsplit, const ObjParams& hp,
             cudaStream_t st) {
  {
    const int rc_attr = ensure_dynamic_smem(oz_sumsq_kernel, OZ_SMEM);
    if (rc_attr) return rc_attr;
  }
  const int npad = round_up(n, OZ_TM), nb = npad / OZ_TM;
  if (nsplit > nb) nsplit = nb;
  const int units = tiles * m * nsplit;
  const unsigned grid = (unsigned)(units < device_sm_count() ? units : device_sm_count());
  oz_sumsq_kernel<<<grid, OZ_THREADS, OZ_SMEM, st>>>(part, ld_chunk, wq, (long long)oz_wq_bytes(n), wscale, kq, npad,
                                                     nb, npad / OZ_KS, chunk_tiles, nsplit, m, units, hp);
  BO_LAUNCH_CHECK("oz_sumsq_kernel");
  return BO_OK;
}

int oz_peak_tops(double* tops, double seconds, cudaStream_t st) {
  cudaEvent_t e0, e1;
  BO_CUDA(cudaEventCreate(&e0));
  BO_CUDA(cudaEventCreate(&e1));
  const int sms = device_sm_count();
  const double ops_per_iter = 21.0 * 2.0 * OZ_TM * OZ_TN * OZ_KS;
  oz_peak_probe_kernel<<<sms, 64, 0, st>>>(200);  // warm-up
  BO_LAUNCH_CHECK("oz_peak_probe_kernel");
  // ~0.36 us per iteration at full clocks
  int iters = (int)(seconds / 0.36e-6);
  if (iters < 1000) iters = 1000;
  BO_CUDA(cudaEventRecord(e0, st));
  oz_peak_probe_kernel<<<sms, 64, 0, st>>>(iters);
  BO_LAUNCH_CHECK("oz_peak_probe_kernel");
  BO_CUDA(cudaEventRecord(e1, st));
  BO_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  BO_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tops = (double)sms * iters * ops_per_iter / (ms * 1e-3) / 1e12;
  return BO_OK;
}

// ------------------------------------------------------------------------------------------- sampled guard
namespace {

__device__ __forceinline__ long long guard_pos(long long j, long long stride, long long n_cand) {
  const unsigned h = (unsigned)j * 2654435761u;  // hashed offset inside the window: no aliasing with grid periods
  long long p = j * stride + (long long)((h >> 8) % (unsigned long long)stride);
  return p < n_cand ? p : n_cand - 1;
}

template <typename CT>
__global__ void guard_gather_kernel(double* __restrict__ out, const CT* __restrict__ cand, int ldc, long long n_cand,
                                    long long stride, int n_sample, int d) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_sample) return;
  const long long p = guard_pos(j, stride, n_cand);
  for (int k = 0; k < d; ++k) out[(long long)j * d + k] = (double)cand[p * ldc + k];
}

// wnorm[o] += sum of squares of this CTA's share of the packed W_o  (= trace(K_o^-1) once all CTAs have added)
__global__ void __launch_bounds__(256)
    guard_wnorm_kernel(double* __restrict__ wnorm, const double* __restrict__ wpack, long long strideWp) {
  __shared__ double red[256];
  const int o = blockIdx.y;
  const double* W = wpack + (long long)o * strideWp;
  double s = 0.0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < strideWp; e += (long long)gridDim.x * blockDim.x)
    s = fma(W[e], W[e], s);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(&wnorm[o], red[0]);
}

// kinf[o] = max_i sum_j |K_o(x_i, x_j) + jitter delta_ij| = |K_o + jitter I|_inf >= lambda_max  (one CTA per row; RBF
// entries are positive).  For short length scales this is far below the trace n (var0 + jitter).
__global__ void __launch_bounds__(256)
    guard_rowsum_kernel(double* __restrict__ kinf, const double* __restrict__ x, int ldx, int n, int d, ObjParams hp,
                        double jitter) {
  __shared__ double red[256];
  const int i = blockIdx.x, o = blockIdx.y;
  double s = 0.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    double sq = 0.0;
    for (int k = 0; k < d; ++k) {
      const double diff = x[(long long)i * ldx + k] - x[(long long)j * ldx + k];
      sq = fma(diff, diff, sq);
    }
    s += hp.prior_var[o] * exp(sq * hp.neg_half_inv_ls2[o]);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0)  // positive doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long*>(&kinf[o]),
              (unsigned long long)__double_as_longlong(red[0] + jitter));
}

// res[0] = max over the sample and the objectives of |var_int8 - var_fp64| / prior_variance (NaN counts as +inf)
// res[1] = the tolerance it is held to: max(tol, 10 eps cond_upper) with the rigorous upper bound
//          cond(K + jitter I) <= |K + jitter I|_inf * trace((K + jitter I)^-1) = kinf * |W|_F^2
// res[2] = 1 if some objective exceeded its tolerance
__global__ void guard_compare_kernel(double* __restrict__ res, const double* __restrict__ va,
                                     const double* __restrict__ vb, int n_sample, int m, ObjParams hp, double tol,
                                     const double* __restrict__ wnorm, const double* __restrict__ kinf) {
  __shared__ double red[256];
  double worst_all = 0.0, tau_all = 0.0, bad = 0.0;
  for (int o = 0; o < m; ++o) {
    double mx = 0.0;
    for (int e = threadIdx.x; e < n_sample; e += blockDim.x) {
      const double dv = fabs(va[(long long)o * n_sample + e] - vb[(long long)o * n_sample + e]) / hp.prior_var[o];
      mx = (dv != dv) ? __longlong_as_double(0x7ff0000000000000ll) : fmax(mx, dv);
    }
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int s = 128; s; s >>= 1) {
      if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
      __syncthreads();
    }
    const double worst = red[0];
    __syncthreads();
    const double cond_upper = kinf[o] * wnorm[o];
    const double tau = fmax(tol, 10.0 * 2.220446049250313e-16 * cond_upper);
    worst_all = fmax(worst_all, worst);
    tau_all = fmax(tau_all, tau);
    if (!(worst <= tau)) bad = 1.0;
  }
  if (threadIdx.x == 0) {
    res[0] = worst_all;
    res[1] = tau_all;
    res[2] = bad;
  }
}

int guard_samples(long long n_cand, long long stride) {
  long long s = (n_cand + stride - 1) / stride;
  if (s > 65536) s = 65536;
  return (int)(s < 1 ? 1 : s);
}

}  // namespace

size_t oz_guard_workspace_bytes(int n, int m, int d, long long n_cand, long long stride) {
  const int S = guard_samples(n_cand, stride);
  return align256((size_t)S * d * 8) + 2 * align256((size_t)m * S * 8) + 256 +
         align256(oz_workspace_bytes(make_oz_plan(n, m, S))) + align256(score_workspace_bytes(make_score_plan(n, m, S)));
}

int oz_guard(double* worst_host, double* tau_host, const void* cand, int cand_kind, int ldc, long long n_cand,
             long long stride, const double* x, int ldx, int n, int d, int m, const unsigned char* wq,
             const double* wscale, const double* wpack, const double* alpha, const ObjParams& hp, double jitter,
             double min_variance, double tol, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (workspace_bytes < oz_guard_workspace_bytes(n, m, d, n_cand, stride)) {
    set_error("int8 guard workspace too small");
    return BO_ERR_WORKSPACE;
  }
  const int S = guard_samples(n_cand, stride);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  size_t off = 0;
  double* sample = reinterpret_cast<double*>(ws + off);  off += align256((size_t)S * d * 8);
  double* va = reinterpret_cast<double*>(ws + off);      off += align256((size_t)m * S * 8);
  double* vb = reinterpret_cast<double*>(ws + off);      off += align256((size_t)m * S * 8);
  double* res = reinterpret_cast<double*>(ws + off);     off += 256;  // [worst, tau, bad, -, wnorm[4], kinf[4]]
  double* wnorm = res + 4;
  double* kinf = res + 8;
  void* ws_i8 = ws + off;
  const size_t ws_i8_bytes = oz_workspace_bytes(make_oz_plan(n, m, S));
  off += align256(ws_i8_bytes);
  void* ws_f64 = ws + off;
  const size_t ws_f64_bytes = score_workspace_bytes(make_score_plan(n, m, S));
  if (cand_kind == BO_CAND_I64)
    guard_gather_kernel<long long><<<(S + 255) / 256, 256, 0, stream>>>(sample, static_cast<const long long*>(cand), ldc,
                                                                        n_cand, stride, S, d);
  else
    guard_gather_kernel<double><<<(S + 255) / 256, 256, 0, stream>>>(sample, static_cast<const double*>(cand), ldc,
                                                                     n_cand, stride, S, d);
  BO_LAUNCH_CHECK("guard_gather_kernel");
  ScoreOutputs oa, ob;
  oa.var = va; oa.ld = S;
  ob.var = vb; ob.ld = S;
  // a candidate's result does not depend on the chunking, so these are bit for bit the numbers of the main pass
  int rc = oz_score_candidates(oa, sample, BO_CAND_F64, d, S, x, ldx, n, d, m, wq, wscale, alpha, hp, min_variance,
                               ws_i8, ws_i8_bytes, stream);
  if (rc) return rc;
  rc = score_candidates(ob, sample, BO_CAND_F64, d, S, x, ldx, n, d, m, wpack, alpha, hp, min_variance, ws_f64,
                        ws_f64_bytes, stream);
  if (rc) return rc;
  const long long strideWp = (long long)wpack_tile_offset(round_up(n, TM) / TM) * TILE_DOUBLES;
  BO_CUDA(cudaMemsetAsync(wnorm, 0, 8 * sizeof(double), stream));
  guard_wnorm_kernel<<<dim3(device_sm_count(), m), 256, 0, stream>>>(wnorm, wpack, strideWp);
  BO_LAUNCH_CHECK("guard_wnorm_kernel");
  guard_rowsum_kernel<<<dim3(n, m), 256, 0, stream>>>(kinf, x, ldx, n, d, hp, jitter);
  BO_LAUNCH_CHECK("guard_rowsum_kernel");
  guard_compare_kernel<<<1, 256, 0, stream>>>(res, va, vb, S, m, hp, tol, wnorm, kinf);
  BO_LAUNCH_CHECK("guard_compare_kernel");
  double r[3] = {0.0, 0.0, 0.0};
  BO_CUDA(cudaMemcpyAsync(r, res, sizeof(r), cudaMemcpyDeviceToHost, stream));
  BO_CUDA(cudaStreamSynchronize(stream));
  if (worst_host) *worst_host = r[0];
  if (tau_host) *tau_host = r[1];
  if (r[2] != 0.0) {
    set_error("int8 variance engine guard: |var_int8 - var_fp64| / prior_variance = %.3e exceeds %.3e on a sample of "
              "%d candidates (no fallback: use variance_engine=\"dmma\")", r[0], r[1], S);
    return BO_ERR_GUARD;
  }
  return BO_OK;
}

OzPlan make_oz_plan(int n, int m, long long n_cand) {
  OzPlan p;
  p.npad = round_up(n, OZ_TM);
  p.nb = p.npad / OZ_TM;
  p.nk_tot = p.npad / OZ_KS;
  const long long tiles = (n_cand + OZ_TN - 1) / OZ_TN;
  const int sms = device_sm_count();
  // K* tiles that are live at the same time must fit in L2 next to W: one tile is 6 * 64 * npad bytes
  const long long tile_bytes = (long long)OZ_PLANES * OZ_TN * p.npad;
  long long ns = ((long long)sms * tile_bytes + (64LL << 20) - 1) / (64LL << 20);
  if (ns < 1) ns = 1;
  if (ns > p.nb) ns = p.nb;
  p.nsplit = (int)ns;
  long long ct = 8LL * sms;
  const long long cap = (3LL << 29) / (tile_bytes * m);  // staging buffer near 1.5 GB
  if (ct > cap) ct = cap < 1 ? 1 : cap;
  if (ct > tiles) ct = tiles;
  if (ct < 1) ct = 1;
  p.chunk_tiles = (int)ct;
  p.ld_chunk = ct * OZ_TN;
  p.kq_bytes = (size_t)m * ct * tile_bytes;
  p.part_doubles = (size_t)m * p.nsplit * p.ld_chunk;
  p.mean_doubles = (size_t)m * p.ld_chunk;
  return p;
}

size_t oz_workspace_bytes(const OzPlan& p) {
  return align256(p.kq_bytes) + align256(p.mean_doubles * 8) + align256(p.part_doubles * 8);
}

int oz_score_candidates(const ScoreOutputs& out, const void* cand, int cand_kind, int ldc, long long n_cand,
                        const double* x, int ldx, int n, int d, int m, const unsigned char* wq,
                        const double* wscale, const double* alpha, const ObjParams& hp, double min_variance,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (n_cand <= 0) return BO_OK;
  const OzPlan p = make_oz_plan(n, m, n_cand);
  if (p.npad > OZ_MAX_N) {
    set_error("int8 engine: n = %d exceeds %d (int32 accumulator bound)", n, OZ_MAX_N);
    return BO_ERR_INVALID;
  }
  if (workspace_bytes < oz_workspace_bytes(p)) {
    set_error("int8 score workspace too small: %zu < %zu", workspace_bytes, oz_workspace_bytes(p));
    return BO_ERR_WORKSPACE;
  }
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  unsigned char* kq = ws;
  double* meandot = reinterpret_cast<double*>(ws + align256(p.kq_bytes));
  double* part = reinterpret_cast<double*>(ws + align256(p.kq_bytes) + align256(p.mean_doubles * 8));
  const long long n_chunks = (n_cand + p.ld_chunk - 1) / p.ld_chunk;
  for (long long ci = 0; ci < n_chunks; ++ci) {
    const long long cand0 = ci * p.ld_chunk;
    const long long remaining = n_cand - cand0;
    const long long live = remaining < p.ld_chunk ? remaining : p.ld_chunk;
    const int tiles = (int)((live + OZ_TN - 1) / OZ_TN);
    int rc;
    {
      // algorithmic bytes: six digit planes per K* entry (6 m npad per candidate) + the candidates read
      ProfileScope prof_scope(stream, BO_PROF_KSTAR, (double)tiles * OZ_TN * ((double)m * p.npad * 6.0 + 8.0 * d));
      rc = oz_kstar_digits(kq, meandot, cand, cand_kind, ldc, cand0, n_cand, tiles, p.chunk_tiles, x, ldx, n, d, m,
                           alpha, hp, stream);
    }
    if (rc) return rc;
    const bool prof = profile_enabled();
    if (prof) profile_begin(stream);
    rc = oz_sumsq(part, p.ld_chunk, wq, wscale, kq, n, m, tiles, p.chunk_tiles, p.nsplit, hp, stream);
    if (prof) profile_end(stream, (double)live * m * (double)n * (double)n);
    if (rc) return rc;
    rc = finalize_chunk(out, cand0, n_cand, part, meandot, p.ld_chunk, tiles * OZ_TN, p.nsplit, m, hp, min_variance,
                        stream);
    if (rc) return rc;
  }
  return BO_OK;
}

}  // namespace bo
