// Shared helpers for libmgs.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>

#include "../../include/mgs.h"

namespace mgs {

// ---- error / bookkeeping (defined in common.cu) -------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launch_count;
int sm_count();

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MGS_ERR_CUDA;
  }
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return MGS_OK;
}

#define MGS_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::mgs::set_error(__VA_ARGS__);           \
      return MGS_ERR_INVALID_ARGUMENT;         \
    }                                          \
  } while (0)

#define MGS_CUDA(expr)                                              \
  do {                                                              \
    cudaError_t e__ = (expr);                                       \
    if (e__ != cudaSuccess) {                                       \
      ::mgs::set_error("%s: %s", #expr, cudaGetErrorString(e__));   \
      return MGS_ERR_CUDA;                                          \
    }                                                               \
  } while (0)

// Grid for a grid-stride loop over `total` work items: enough CTAs to cover the work, capped at
// a whole number of waves (148 SMs x `ctas_per_sm`) so the tail wave is full.
inline int grid_for(int64_t total, int threads, int ctas_per_sm) {
  int64_t need = (total + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// Grid for the row-streaming kernels (a warp owns a contiguous range of rows and walks it 32 at a time): one warp
// per 32 rows when that fills two 8-warp CTAs per SM, otherwise fewer rows per warp (down to 4) -- a 64-molecule
// batch is ~2000 rows = 63 warps of 32 rows on 148 SMs, each row a ~3 us dependent chain.
inline int grid_for_rows(int64_t rows, int warps_per_cta, int ctas_per_sm) {
  const int64_t want = (int64_t)sm_count() * ctas_per_sm * warps_per_cta;
  int rpw = 32;
  while (rpw > 4 && (rows + rpw - 1) / rpw < want) rpw >>= 1;
  int64_t ctas = ((rows + rpw - 1) / rpw + warps_per_cta - 1) / warps_per_cta;
  const int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  if (ctas < 1) ctas = 1;
  return (int)(ctas < cap ? ctas : cap);
}

// widest vector width (in floats) usable for rows starting at p with leading dimension ld and nf columns
inline int vec_width(const void* p, int64_t ld, int64_t nf) {
  uintptr_t a = (uintptr_t)p;
  if (a % 16 == 0 && ld % 4 == 0 && nf % 4 == 0) return 4;
  if (a % 8 == 0 && ld % 2 == 0 && nf % 2 == 0) return 2;
  return 1;
}
inline int min_int(int a, int b) { return a < b ? a : b; }

// ---- small device vector abstraction -------------------------------------------------------
template <int V> struct Vec;
template <> struct Vec<1> {
  float v[1];
  __device__ __forceinline__ static Vec load(const float* p) { Vec r; r.v[0] = __ldg(p); return r; }
  __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
};
template <> struct Vec<2> {
  float v[2];
  __device__ __forceinline__ static Vec load(const float* p) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p)); Vec r; r.v[0] = t.x; r.v[1] = t.y; return r;
  }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
};
template <> struct Vec<4> {
  float v[4];
  __device__ __forceinline__ static Vec load(const float* p) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p)); Vec r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
  }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <int V> __device__ __forceinline__ Vec<V> vzero() {
  Vec<V> r;
#pragma unroll
  for (int k = 0; k < V; ++k) r.v[k] = 0.f;
  return r;
}

// ---- exact division by a row count -------------------------------------------------------------
// a / b rounded to nearest even, bit-identical to __fdiv_rn(a, b), for a positive normal b (an in-degree)
// and rc = __frcp_rn(b).  nvcc's IEEE division costs ~10 issue slots per element and drops into a ~100
// instruction subroutine whenever the dividend is zero or denormal -- and half of a post-ReLU activation
// matrix IS zero: on B200 the SAGE aggregation ran 2x (forward) to 2.7x (backward) slower on real
// activations than on dense random rows for that reason alone.  One reciprocal per row and the classical
// residual correction (q' = q + (a - b q) rc, correctly rounded when rc is, Markstein 1990) give the same
// bits in 3 FMAs; dividends outside [2^-100, 2^100] (where the residual may be inexact) and zeros (to keep
// the sign of zero) leave through a cold branch.  Checked exhaustively against __fdiv_rn by
// mgs_selftest_div (tests/test_gpu_kernels.py).
__device__ __forceinline__ float div_by_count(float a, float b, float rc) {
  const float q = __fmul_rn(a, rc);
  const float r = __fmaf_rn(-b, q, a);
  const float aa = fabsf(a);
  if (!(aa >= 0x1p-100f && aa <= 0x1p100f)) return aa == 0.f ? a : __fdiv_rn(a, b);
  return __fmaf_rn(r, rc, q);
}

// ---- warp-per-row mapping ---------------------------------------------------------------------
// A warp owns one row of `chunks` V-wide chunks; lane l owns chunks l, l+32, ... (ITERS of them, the
// tail predicated off).  Consecutive lanes touch consecutive 8/16-byte chunks: every load/store
// instruction of the warp is one fully coalesced 256/512-byte access, row-level metadata (row
// pointers, neighbour ids, attention weights) is loaded once per warp instead of once per thread,
// and ITERS x (edges in flight) independent loads per lane give the memory-level parallelism.
inline int iters_for(int chunks) {
  const int need = (chunks + 31) / 32;
  if (need <= 1) return 1;
  if (need <= 2) return 2;
  if (need <= 4) return 4;
  if (need <= 6) return 6;
  if (need <= 8) return 8;
  return 0;  // too wide: caller falls back to the flat (thread per chunk) kernel
}

#define MGS_DISPATCH_V_ITERS(V, IT, LAUNCH)                   \
  do {                                                        \
    if ((V) == 4) {                                           \
      if ((IT) == 1) { LAUNCH(4, 1); } else if ((IT) == 2) { LAUNCH(4, 2); } else if ((IT) == 4) { LAUNCH(4, 4); } \
      else if ((IT) == 6) { LAUNCH(4, 6); } else { LAUNCH(4, 8); }                                              \
    } else if ((V) == 2) {                                    \
      if ((IT) == 1) { LAUNCH(2, 1); } else if ((IT) == 2) { LAUNCH(2, 2); } else if ((IT) == 4) { LAUNCH(2, 4); } \
      else if ((IT) == 6) { LAUNCH(2, 6); } else { LAUNCH(2, 8); }                                              \
    } else {                                                  \
      if ((IT) == 1) { LAUNCH(1, 1); } else if ((IT) == 2) { LAUNCH(1, 2); } else if ((IT) == 4) { LAUNCH(1, 4); } \
      else if ((IT) == 6) { LAUNCH(1, 6); } else { LAUNCH(1, 8); }                                              \
    }                                                         \
  } while (0)

}  // namespace mgs
