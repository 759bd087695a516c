// gemm.cu -- general FP64 GEMM on DMMA (mma.sync.m8n8k4.f64) for the factorisation side of the path:
// Cholesky panel/trailing updates, triangular inverse, W^T W, and the dense drop-ins.
// C = beta*C + alpha * opA * opB, row-major, batched over blockIdx.z.  64x64x16 tiles, 4 warps (2x2),
// register-prefetched double buffering.  The scoring hot loop does NOT use this kernel (see trmm.cu).
#include "gemm.cuh"

namespace bo {

namespace {

constexpr int GB = 64;    // tile rows / cols
constexpr int GK = 16;    // tile depth
constexpr int LDK = 20;   // [row][k] layout, padded: fragment LDS.64 is conflict free (4g+t distinct mod 16)
constexpr int LDR = 68;   // [k][row] layout, padded

// MODE 0: operand is K-contiguous  (A[i*ld + k] or B[j*ld + k])  -> smem [row][k]
// MODE 1: operand is row-contiguous (A[k*ld + i] or B[k*ld + j]) -> smem [k][row]
template <int MODE>
struct TileLoader {
  double2 v[4];
  // row0: first row (i or j) of the tile, k0: first k; R = valid rows of operand, K = valid depth
  __device__ __forceinline__ void load(const double* __restrict__ P, long long ld, int row0, int k0, int R, int K,
                                       bool fast, int tid) {
    if (MODE == 0) {
      const int kk = (tid & 7) * 2;
      const int r = tid >> 3;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int row = row0 + r + 16 * p;
        const int k = k0 + kk;
        if (fast) {
          v[p] = *reinterpret_cast<const double2*>(P + (long long)row * ld + k);
        } else {
          double a = 0.0, b = 0.0;
          if (row < R) {
            if (k < K) a = P[(long long)row * ld + k];
            if (k + 1 < K) b = P[(long long)row * ld + k + 1];
          }
          v[p] = make_double2(a, b);
        }
      }
    } else {
      const int rr = (tid & 31) * 2;
      const int kq = tid >> 5;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int k = k0 + kq + 4 * p;
        const int row = row0 + rr;
        if (fast) {
          v[p] = *reinterpret_cast<const double2*>(P + (long long)k * ld + row);
        } else {
          double a = 0.0, b = 0.0;
          if (k < K) {
            if (row < R) a = P[(long long)k * ld + row];
            if (row + 1 < R) b = P[(long long)k * ld + row + 1];
          }
          v[p] = make_double2(a, b);
        }
      }
    }
  }
  __device__ __forceinline__ void store(double* S, int tid) const {
    if (MODE == 0) {
      const int kk = (tid & 7) * 2;
      const int r = tid >> 3;
#pragma unroll
      for (int p = 0; p < 4; ++p) *reinterpret_cast<double2*>(S + (r + 16 * p) * LDK + kk) = v[p];
    } else {
      const int rr = (tid & 31) * 2;
      const int kq = tid >> 5;
#pragma unroll
      for (int p = 0; p < 4; ++p) *reinterpret_cast<double2*>(S + (kq + 4 * p) * LDR + rr) = v[p];
    }
  }
  static constexpr int smem_doubles = (MODE == 0) ? GB * LDK : GK * LDR;
  // element (row r, depth k) of the staged tile
  __device__ __forceinline__ static double at(const double* S, int r, int k) {
    return (MODE == 0) ? S[r * LDK + k] : S[k * LDR + r];
  }
};

template <int AMODE, int BMODE>
__global__ void __launch_bounds__(128) gemm64_kernel(GemmArgs g) {
  // lower_only: the grid is the LIST of tiles on or below the diagonal (no empty CTAs: a 60 x 60-tile trailing update
  // of 128 matrices would otherwise dispatch 230 000 CTAs that return at once).  Tile rows by < NX hold by + 1 tiles,
  // the rows below hold NX tiles each.
  int tile_x = blockIdx.x, tile_y = blockIdx.y;
  if (g.lower_only) {
    const int NX = (g.N + GB - 1) / GB;
    const int t = blockIdx.x, tri = NX * (NX + 1) / 2;
    if (t < tri) {
      int by = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);  // estimate; the two loops make it exact
      while (by * (by + 1) / 2 > t) --by;
      while ((by + 1) * (by + 2) / 2 <= t) ++by;
      tile_y = by;
      tile_x = t - by * (by + 1) / 2;
    } else {
      tile_y = NX + (t - tri) / NX;
      tile_x = (t - tri) % NX;
    }
  }
  __shared__ __align__(16) double As[2][TileLoader<AMODE>::smem_doubles];
  __shared__ __align__(16) double Bs[2][TileLoader<BMODE>::smem_doubles];

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int g8 = lane >> 2, t4 = lane & 3;
  const int row0 = tile_y * GB, col0 = tile_x * GB;
  const double* A = g.A + (long long)blockIdx.z * g.strideA;
  const double* B = g.B + (long long)blockIdx.z * g.strideB;
  double* C = g.C + (long long)blockIdx.z * g.strideC;

  // optional k-range restriction for triangular operands (exact: the skipped products are zeros)
  int kbeg = 0, kend = g.K;
  if (g.k_limit_rows) kend = min(g.K, row0 + GB);       // A lower triangular: A(i,k)=0 for k>i
  if (g.k_start_cols) kbeg = (col0 / GK) * GK;          // B lower triangular (NN): B(k,j)=0 for k<j
  const bool fast = g.fast != 0;

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  TileLoader<AMODE> la;
  TileLoader<BMODE> lb;
  int buf = 0;
  if (kbeg < kend) {
    la.load(A, g.lda, row0, kbeg, g.M, g.K, fast, tid);
    lb.load(B, g.ldb, col0, kbeg, g.N, g.K, fast, tid);
    la.store(As[0], tid);
    lb.store(Bs[0], tid);
  }
  __syncthreads();
  for (int k0 = kbeg; k0 < kend; k0 += GK) {
    const bool more = (k0 + GK) < kend;
    if (more) {
      la.load(A, g.lda, row0, k0 + GK, g.M, g.K, fast, tid);
      lb.load(B, g.ldb, col0, k0 + GK, g.N, g.K, fast, tid);
    }
    const double* as = As[buf];
    const double* bs = Bs[buf];
#pragma unroll
    for (int kk = 0; kk < GK / 4; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = TileLoader<AMODE>::at(as, wm * 32 + i * 8 + g8, kk * 4 + t4);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = TileLoader<BMODE>::at(bs, wn * 32 + j * 8 + g8, kk * 4 + t4);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    if (more) {
      la.store(As[buf ^ 1], tid);
      lb.store(Bs[buf ^ 1], tid);
    }
    __syncthreads();
    buf ^= 1;
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + wm * 32 + i * 8 + g8;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + wn * 32 + j * 8 + 2 * t4;
      if (r < g.M) {
        double* p = C + (long long)r * g.ldc + c;
        if (c < g.N) p[0] = (g.beta == 0.0 ? 0.0 : g.beta * p[0]) + g.alpha * acc[i][j][0];
        if (c + 1 < g.N) p[1] = (g.beta == 0.0 ? 0.0 : g.beta * p[1]) + g.alpha * acc[i][j][1];
      }
    }
  }
}

}  // namespace

int gemm(const GemmArgs& a_in, int amode, int bmode, cudaStream_t stream) {
  GemmArgs a = a_in;
  if (a.M <= 0 || a.N <= 0 || a.batch <= 0) return BO_OK;
  // vectorised unguarded path: whole tiles, even leading dimensions, 16 B aligned bases
  bool fast = (a.M % GB == 0) && (a.N % GB == 0) && (a.K % GK == 0) && (a.lda % 2 == 0) && (a.ldb % 2 == 0) &&
              (((uintptr_t)a.A & 15) == 0) && (((uintptr_t)a.B & 15) == 0) && (a.strideA % 2 == 0) &&
              (a.strideB % 2 == 0);
  a.fast = fast ? 1 : 0;
  const int NX = (a.N + GB - 1) / GB, NY = (a.M + GB - 1) / GB;
  dim3 grid(NX, NY, a.batch);
  if (a.lower_only) {  // tiles with tile_x <= tile_y only, as a flat list
    const long long live = NY >= NX ? (long long)NX * (NX + 1) / 2 + (long long)(NY - NX) * NX
                                    : (long long)NY * (NY + 1) / 2;
    grid = dim3((unsigned)live, 1, a.batch);
  }
  if (amode == 0 && bmode == 0)
    gemm64_kernel<0, 0><<<grid, 128, 0, stream>>>(a);
  else if (amode == 0 && bmode == 1)
    gemm64_kernel<0, 1><<<grid, 128, 0, stream>>>(a);
  else if (amode == 1 && bmode == 1)
    gemm64_kernel<1, 1><<<grid, 128, 0, stream>>>(a);
  else if (amode == 1 && bmode == 0)
    gemm64_kernel<1, 0><<<grid, 128, 0, stream>>>(a);
  BO_LAUNCH_CHECK("gemm64_kernel");
  return BO_OK;
}

}  // namespace bo
