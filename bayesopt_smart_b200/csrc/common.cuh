// common.cuh -- shared host/device helpers for libbo_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/bo_b200.h"

namespace bo {

// ---------------------------------------------------------------- error state
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define BO_CUDA(call)                                            \
  do {                                                           \
    cudaError_t _e = (call);                                     \
    if (_e != cudaSuccess) return ::bo::cuda_fail(_e, #call);    \
  } while (0)
#define BO_LAUNCH_CHECK(name)                                    \
  do {                                                           \
    ::bo::count_launch();                                        \
    cudaError_t _e = cudaGetLastError();                         \
    if (_e != cudaSuccess) return ::bo::cuda_fail(_e, name);     \
  } while (0)
#define BO_REQUIRE(cond, msg)                                    \
  do {                                                           \
    if (!(cond)) {                                               \
      ::bo::set_error("invalid argument: %s", msg);              \
      return BO_ERR_INVALID;                                     \
    }                                                            \
  } while (0)

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
static inline long long round_up_ll(long long a, long long b) { return (a + b - 1) / b * b; }
static inline size_t align256(size_t b) { return (b + 255) & ~size_t(255); }

// per-objective hyper-parameters travel by value in kernel parameters
struct ObjParams {
  double prior_mean[BO_MAX_OBJECTIVES];
  double prior_var[BO_MAX_OBJECTIVES];
  double neg_half_inv_ls2[BO_MAX_OBJECTIVES];  // -0.5 / ls^2
  double beta[BO_MAX_OBJECTIVES];
};

int device_sm_count();  // of the CURRENT device (cached per device)
void count_launch();
// cudaFuncSetAttribute(fn, MaxDynamicSharedMemorySize, bytes), remembered per (function, device): a process that
// drives several GPUs sets the attribute once on each of them (the attribute is per device / context)
int ensure_dynamic_smem_impl(const void* fn, size_t bytes);
template <typename F>
inline int ensure_dynamic_smem(F* fn, size_t bytes) {
  return ensure_dynamic_smem_impl(reinterpret_cast<const void*>(fn), bytes);
}
// live CUDA-event timing of the dominant kernel (see bo_profile_enable in bo_b200.h)
bool profile_enabled();
void profile_begin(cudaStream_t st, int kernel = BO_PROF_CONTRACTION);  // kernel: a BO_PROF_* class
void profile_end(cudaStream_t st, double work);  // work: algorithmic flops (contraction) or bytes (the others)
// brackets one launch (or a short launch sequence) when profiling is on
struct ProfileScope {
  cudaStream_t st;
  double work;
  bool on;
  ProfileScope(cudaStream_t s, int kernel, double w) : st(s), work(w), on(profile_enabled()) {
    if (on) profile_begin(st, kernel);
  }
  ~ProfileScope() {
    if (on) profile_end(st, work);
  }
};

// ---------------------------------------------------------------- geometry of the packed operands
// W (= L^-1, lower triangular, npad x npad) and K* (npad x candidates) are stored as 16 KB tiles in
// the register-fragment order of mma.sync.m8n8k4.f64 so that one cp.async.bulk brings a whole tile
// and every fragment load is a conflict-free LDS.128.
constexpr int TM = 128;             // rows of W per tile / row block
constexpr int TN = 128;             // candidates per tile
constexpr int TK = 16;              // k depth per tile
constexpr int TILE_DOUBLES = TM * TK;  // 2048 doubles = 16 KB (A and B tiles have the same size)
constexpr int KT_PER_BLOCK = TM / TK;  // 8 k-tiles per 128-row block

// number of W tiles stored before row block ib (only k-tiles up to the diagonal block are kept)
__host__ __device__ inline long long wpack_tile_offset(int ib) {
  return (long long)KT_PER_BLOCK * ib * (ib + 1) / 2;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  // D(8x8) += A(8x4) * B(4x8); lowers to one DMMA.8x8x4 on sm_100a
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP); bytes % 16 == 0, 16 B aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ double2 lds128(const double* p) { return *reinterpret_cast<const double2*>(p); }

// order-preserving key: larger double -> larger uint64; NaN -> 0 (sorts last); -0.0 == +0.0
__device__ __forceinline__ unsigned long long order_key(double v) {
  if (v != v) return 0ull;
  if (v == 0.0) v = 0.0;
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
#endif

}  // namespace bo
