// rbf.cuh -- exp() for the RBF kernel entries, shared by gram_kernel, kstar_dense_kernel and kstar_pack_kernel.
//
// On B200 DMMA and DFMA issue to the same FP64 units (tools/pipe_probe.cu), so every FP64 instruction spent
// on exp() is taken from the triangular product.  This version needs 10 FP64 ops (library exp: ~17):
//   n = rint(x * 64/ln2),  r = x - n*ln2/64 (two-term Cody-Waite, |r| <= ln2/128),
//   exp(x) = 2^(n>>6) * T[n & 63] * (1 + r + r^2/2 + r^3/6 + r^4/24 + r^5/120),   T[j] = 2^(j/64)
// Truncation error r^6/720 < 3.6e-17; total error ~1 ulp (the reference's own exp runs under fastmath).
// Domain: x <= 0 (x = -0.5 * sq / ls^2).  Results below exp(-706) are flushed to zero.
#pragma once

namespace bo {

// correctly rounded 2^(j/64) (generated with 200-bit arithmetic)
static __device__ const double kExp2Tab[64] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951,
};

// Coefficients live in the constant bank so that DFMA takes them as direct c[][] operands; as literals the
// compiler rebuilt each 64-bit constant with two UMOVs per use (17 % of the K* kernel's issue slots).
//   [0] 64/ln2   [1] -ln2/64 high part (low 32 mantissa bits zero: n*hi exact)   [2] -(ln2/64 - hi)
//   [3] 1/120    [4] 1/24    [5] 1/6
static __device__ __constant__ double kExpCoef[6] = {
    92.33248261689366, -0.01083042469326756, -2.9815858269852933e-12,
    1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0};

// tab: 64 doubles 2^(j/64), in shared memory (hot kernel) or kExp2Tab itself (L1-cached global).
// FLUSH = false skips the final flush to zero: inputs below -706 then return exp(-706) = 2.5e-307 instead of 0,
// which is the same thing to a caller that quantises the result (ozaki.cu) and saves four instructions per call.
template <bool FLUSH = true>
__device__ __forceinline__ double rbf_exp(double x, const double* __restrict__ tab) {
  const double kMagic = 6755399441055744.0;  // 1.5 * 2^52: adds round-to-nearest-integer
  const double xc = fmax(x, -706.0);         // keeps n inside int range and the result normal
  const double t = fma(xc, kExpCoef[0], kMagic);
  const int n = __double2loint(t);
  const double nd = t - kMagic;
  double r = fma(nd, kExpCoef[1], xc);
  r = fma(nd, kExpCoef[2], r);
  double p = fma(r, kExpCoef[3], kExpCoef[4]);
  p = fma(p, r, kExpCoef[5]);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double v = tab[n & 63] * p;
  const double scaled = __hiloint2double(__double2hiint(v) + ((n >> 6) << 20), __double2loint(v));
  if (!FLUSH) return scaled;
  return (x < -706.0) ? 0.0 : scaled;  // exp(x) < 2.5e-307 is flushed to zero (also maps NaN-free inputs only)
}

}  // namespace bo
