// K4 (baseline variant) -- fp32 FFMA GEMM for the dense projections and the readout MLP
// (SURVEY.md section 8 rows a3/a4/a7/a9): forward C = A W^T (+ A2 W2^T) + bias, dgrad dA = G W,
// wgrad dW = G^T A (split over the long M dimension, deterministic reduction), column sums for
// bias gradients.  Plain IEEE fp32 multiply-add: meets the 1e-5 logit tolerance without TF32
// splitting.  The tcgen05 3xTF32 path for the K >= 256 projections replaces this kernel per
// shape in tc_linear.cu; this file stays the path for the HBM-bound K = 35 projections.
//
// Tiling: 128 x 128 x 16 CTA tile, 256 threads, 8 x 8 register tile per thread (two 4-wide strips
// per dimension so shared-memory reads are conflict-free 128-bit loads), double-buffered shared
// memory with the next tile's global loads in flight during the FFMA loop.
#include "common.cuh"
#include "tc_linear.cuh"
#include "tc_tma.cuh"
#include "tc_wgrad.cuh"

#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdlib>

namespace mgs {
namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int kThreads = 256;
constexpr int LDS = BM + 4;  // padded shared-memory row (floats); keeps rows 16-byte aligned

struct Operand {
  const float* p;
  int64_t ld;
  int vec;  // widest aligned vector (1/2/4 floats) along the contiguous dimension
};

struct Segment {  // one (A, B) pair contracted over K
  Operand a, b;
  int K;
};

// Tile of an operand whose CONTRACTION index is contiguous in memory: elem(r, k) = p[r * ld + k].
// Thread owns row (tid & 127), k range [(tid >> 7) * 8, +8).  Stored transposed: s[k][r].
__device__ __forceinline__ void load_kc(const Operand& op, int r0, int rows, int k0, int kend, float (&reg)[8]) {
  const int r = r0 + (threadIdx.x & 127);
  const int k = k0 + (threadIdx.x >> 7) * 8;
#pragma unroll
  for (int u = 0; u < 8; ++u) reg[u] = 0.f;
  if (r >= rows) return;
  const float* src = op.p + (int64_t)r * op.ld + k;
  if (op.vec == 4 && k + 8 <= kend) {
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(src));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
    reg[0] = v0.x; reg[1] = v0.y; reg[2] = v0.z; reg[3] = v0.w;
    reg[4] = v1.x; reg[5] = v1.y; reg[6] = v1.z; reg[7] = v1.w;
  } else if (op.vec >= 2 && k + 8 <= kend) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(src) + u);
      reg[2 * u] = v.x; reg[2 * u + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (k + u < kend) reg[u] = __ldg(src + u);
  }
}
__device__ __forceinline__ void store_kc(float* s, const float (&reg)[8]) {
  const int r = threadIdx.x & 127;
  const int k = (threadIdx.x >> 7) * 8;
#pragma unroll
  for (int u = 0; u < 8; ++u) s[(k + u) * LDS + r] = reg[u];
}

// Tile of an operand whose NON-contraction index is contiguous: elem(r, k) = p[k * ld + r].
// Thread owns k = tid >> 4, r range [(tid & 15) * 8, +8).  Stored as is: s[k][r].
__device__ __forceinline__ void load_mn(const Operand& op, int r0, int rows, int k0, int kend, float (&reg)[8]) {
  const int k = k0 + (threadIdx.x >> 4);
  const int r = r0 + (threadIdx.x & 15) * 8;
#pragma unroll
  for (int u = 0; u < 8; ++u) reg[u] = 0.f;
  if (k >= kend) return;
  const float* src = op.p + (int64_t)k * op.ld + r;
  if (op.vec == 4 && r + 8 <= rows) {
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(src));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
    reg[0] = v0.x; reg[1] = v0.y; reg[2] = v0.z; reg[3] = v0.w;
    reg[4] = v1.x; reg[5] = v1.y; reg[6] = v1.z; reg[7] = v1.w;
  } else if (op.vec >= 2 && r + 8 <= rows) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(src) + u);
      reg[2 * u] = v.x; reg[2 * u + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (r + u < rows) reg[u] = __ldg(src + u);
  }
}
__device__ __forceinline__ void store_mn(float* s, const float (&reg)[8]) {
  const int k = threadIdx.x >> 4;
  const int r = (threadIdx.x & 15) * 8;
  *reinterpret_cast<float4*>(s + k * LDS + r) = make_float4(reg[0], reg[1], reg[2], reg[3]);
  *reinterpret_cast<float4*>(s + k * LDS + r + 4) = make_float4(reg[4], reg[5], reg[6], reg[7]);
}

// C[m][n] = sum over segments, k of A(m,k) * B(k,n)  (+ bias[n]) (ReLU)
//   A_KC: A(m,k) = a[m*lda + k]   else a[k*lda + m]
//   B_KC: B(k,n) = b[n*ldb + k]   else b[k*ldb + n]
// blockIdx.z = K-split of segment 0 (segment 1 must be empty when gridDim.z > 1): partial results go to
// c + z * split_stride.
template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(kThreads, 2)
gemm_kernel(Segment s0, Segment s1, int M, int N, float* __restrict__ c, int64_t ldc, int c_vec,
            const float* __restrict__ bias, int relu, int k_per_split, int64_t split_stride) {
  __shared__ __align__(16) float As[2][BK * LDS];
  __shared__ __align__(16) float Bs[2][BK * LDS];

  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int k_lo = 0, k_hi = s0.K;
  if (gridDim.z > 1) {
    k_lo = blockIdx.z * k_per_split;
    k_hi = min(s0.K, k_lo + k_per_split);
    c += (int64_t)blockIdx.z * split_stride;
  }
  const int nt0 = k_hi > k_lo ? (k_hi - k_lo + BK - 1) / BK : 0;
  const int nt1 = (s1.K + BK - 1) / BK;
  const int nt = nt0 + nt1;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[8], rb[8];
  auto fetch = [&](int t) {
    const bool first = t < nt0;
    const Segment& s = first ? s0 : s1;
    const int k0 = first ? k_lo + t * BK : (t - nt0) * BK;
    const int kend = first ? k_hi : s1.K;
    if (A_KC) load_kc(s.a, m0, M, k0, kend, ra); else load_mn(s.a, m0, M, k0, kend, ra);
    if (B_KC) load_kc(s.b, n0, N, k0, kend, rb); else load_mn(s.b, n0, N, k0, kend, rb);
  };
  auto stash = [&](int buf) {
    if (A_KC) store_kc(As[buf], ra); else store_mn(As[buf], ra);
    if (B_KC) store_kc(Bs[buf], rb); else store_mn(Bs[buf], rb);
  };

  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  if (nt > 0) {
    fetch(0);
    stash(0);
  }
  __syncthreads();
  for (int t = 0; t < nt; ++t) {
    const int buf = t & 1;
    if (t + 1 < nt) fetch(t + 1);
    const float* as = As[buf];
    const float* bs = Bs[buf];
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(as + kk * LDS + ty * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(as + kk * LDS + 64 + ty * 4);
      const float4 b0 = *reinterpret_cast<const float4*>(bs + kk * LDS + tx * 4);
      const float4 b1 = *reinterpret_cast<const float4*>(bs + kk * LDS + 64 + tx * 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (t + 1 < nt) stash(buf ^ 1);
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + jh * 64 + tx * 4;
      if (n >= N) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = acc[i][jh * 4 + j];
        if (bias != nullptr && n + j < N) v[j] += __ldg(bias + n + j);
        if (relu) v[j] = v[j] <= 0.f ? 0.f : v[j];
      }
      float* dst = c + (int64_t)m * ldc + n;
      if (c_vec == 4 && n + 4 <= N) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      } else if (c_vec >= 2 && n + 4 <= N) {
        *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
        *reinterpret_cast<float2*>(dst + 2) = make_float2(v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < N) dst[j] = v[j];
      }
    }
  }
}

__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ part, int splits, int64_t split_stride, int M, int N,
                     float* __restrict__ out, int64_t ldo, const float* __restrict__ bias = nullptr, int relu = 0,
                     int ldp = 0 /* leading dimension of a partial tile; 0: N */) {
  const int64_t total = (int64_t)M * N;
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
    const int m = (int)(t / N);
    const int n = (int)(t - (int64_t)m * N);
    const int64_t tp = ldp > 0 ? (int64_t)m * ldp + n : t;
    float s = 0.f;
    int z = 0;
    for (; z + 8 <= splits; z += 8) {                         // eight partials in flight, added in split order (same bits)
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = part[(int64_t)(z + u) * split_stride + tp];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; z < splits; ++z) s += part[(int64_t)z * split_stride + tp];
    if (bias != nullptr) s += __ldg(bias + n);
    if (relu) s = s <= 0.f ? 0.f : s;
    out[(int64_t)m * ldo + n] = s;
  }
}

// ---- single-output layer (the readout's last nn.Linear(128, 1), train.py:111, ablation/model1.py:64) --------------------------
// As a 128 x 128-tile GEMM with a split contraction this layer cost ~45 us per step (fwd + both gradients + reductions) for
// half a megabyte of data: a dot product per row, an outer product, and a weighted column sum.
__global__ void __launch_bounds__(256)
gemv_rows_kernel(const float* __restrict__ a, int64_t lda, int M, int K, const float* __restrict__ w,
                 const float* __restrict__ bias, int relu, float* __restrict__ c, int64_t ldc) {
  const int lane = threadIdx.x & 31;
  for (int m = blockIdx.x * 8 + (threadIdx.x >> 5); m < M; m += gridDim.x * 8) {
    const float* row = a + (int64_t)m * lda;
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s = fmaf(__ldg(row + k), __ldg(w + k), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      if (bias != nullptr) s += __ldg(bias);
      if (relu) s = s <= 0.f ? 0.f : s;
      c[(int64_t)m * ldc] = s;
    }
  }
}
// da[m, k] = g[m] * w[k]
__global__ void __launch_bounds__(256)
outer_rows_kernel(const float* __restrict__ g, int64_t ldg, int M, int K, const float* __restrict__ w,
                  float* __restrict__ da, int64_t ldda) {
  const int64_t total = (int64_t)M * K;
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
    const int m = (int)(t / K), k = (int)(t - (int64_t)m * K);
    da[(int64_t)m * ldda + k] = __ldg(g + (int64_t)m * ldg) * __ldg(w + k);
  }
}
// part[cta][k] = sum over the CTA's rows m (m = cta, cta + grid, ...) of g[m] * a[m, k]; summed by colsum_final_kernel
__global__ void __launch_bounds__(256)
wsum_partial_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ a, int64_t lda, int M, int K,
                    float* __restrict__ part) {
  for (int k = threadIdx.x; k < K; k += 256) {
    float s = 0.f;
    int m = blockIdx.x;
    for (; m + 7 * (int)gridDim.x < M; m += 8 * gridDim.x) {
      float gv[8], av[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        gv[u] = __ldg(g + (int64_t)(m + u * (int)gridDim.x) * ldg);
        av[u] = __ldg(a + (int64_t)(m + u * (int)gridDim.x) * lda + k);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) s = fmaf(gv[u], av[u], s);
    }
    for (; m < M; m += gridDim.x) s = fmaf(__ldg(g + (int64_t)m * ldg), __ldg(a + (int64_t)m * lda + k), s);
    part[(int64_t)blockIdx.x * K + k] = s;
  }
}
constexpr int kWsumParts = 128;

constexpr int kColParts = 592;   // 4 CTAs per SM on 148 SMs
constexpr int kColWarps = 8;

// Column sums, stage 1.  Warp-per-row mapping (common.cuh): each warp streams whole rows with coalesced
// vector loads, 4 rows in flight, and keeps ITERS x V partial sums per lane; the 8 warps of a CTA are
// combined through shared memory in fixed order, so the result is deterministic.  Reads g exactly once.
template <int V, int ITERS>
__global__ void __launch_bounds__(kColWarps * 32)
colsum_partial_row_kernel(const float* __restrict__ g, int64_t ldg, int M, int N, int chunks,
                          float* __restrict__ part) {
  extern __shared__ float sm[];  // [kColWarps][N]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Vec<V> acc[ITERS];
#pragma unroll
  for (int t = 0; t < ITERS; ++t) acc[t] = vzero<V>();
  const int stride = gridDim.x * kColWarps;
  int r = blockIdx.x * kColWarps + warp;
  for (; r + 3 * stride < M; r += 4 * stride) {
    Vec<V> v[4][ITERS];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int t = 0; t < ITERS; ++t)
        if (lane + 32 * t < chunks) v[k][t] = Vec<V>::load(g + (int64_t)(r + k * stride) * ldg + (lane + 32 * t) * V);
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int t = 0; t < ITERS; ++t)
        if (lane + 32 * t < chunks)
#pragma unroll
          for (int u = 0; u < V; ++u) acc[t].v[u] += v[k][t].v[u];
  }
  for (; r < M; r += stride) {
#pragma unroll
    for (int t = 0; t < ITERS; ++t)
      if (lane + 32 * t < chunks) {
        Vec<V> v = Vec<V>::load(g + (int64_t)r * ldg + (lane + 32 * t) * V);
#pragma unroll
        for (int u = 0; u < V; ++u) acc[t].v[u] += v.v[u];
      }
  }
#pragma unroll
  for (int t = 0; t < ITERS; ++t)
    if (lane + 32 * t < chunks)
#pragma unroll
      for (int u = 0; u < V; ++u) sm[warp * N + (lane + 32 * t) * V + u] = acc[t].v[u];
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += kColWarps * 32) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kColWarps; ++w) s += sm[w * N + n];
    part[(int64_t)blockIdx.x * N + n] = s;
  }
}

// fallback for very wide matrices: thread per column
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ g, int64_t ldg, int M, int N, float* __restrict__ part) {
  for (int n = threadIdx.x; n < N; n += 256) {
    float s = 0.f;
    int m = blockIdx.x;
    for (; m + 7 * (int)gridDim.x < M; m += 8 * gridDim.x) {   // eight loads in flight (a plain loop exposes one round trip each)
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(g + (int64_t)(m + u * (int)gridDim.x) * ldg + n);
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; m < M; m += gridDim.x) s += __ldg(g + (int64_t)m * ldg + n);
    part[(int64_t)blockIdx.x * N + n] = s;
  }
}
// stage 2: out[n] = sum_p part[p][n].  32 columns x 8 partial-slices per CTA (the first version had one thread
// walk all 592 partials of a column: 30 us of pure load latency for 350 columns).  Fixed order: deterministic.
// (Round 2: 32 slices of 1024 threads and eight loads in flight per thread -- the 8-slice version was a chain of 74
// dependent round trips, 9 us per launch, four launches per step.)
__global__ void __launch_bounds__(1024)
colsum_final_kernel(const float* __restrict__ part, int parts, int N, float* __restrict__ out) {
  __shared__ float sm[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (n < N) {
    int p = ty;
    for (; p + 7 * 32 < parts; p += 8 * 32) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = part[(int64_t)(p + 32 * u) * N + n];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; p < parts; p += 32) s += part[(int64_t)p * N + n];
  }
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += sm[k][tx];
    out[n] = t;
  }
}

Operand make_operand(const float* p, int64_t ld, bool contiguous_is_k, int64_t extent_contig) {
  Operand o;
  o.p = p;
  o.ld = ld;
  uintptr_t a = (uintptr_t)p;
  o.vec = 1;
  if (a % 16 == 0 && ld % 4 == 0) o.vec = 4;
  else if (a % 8 == 0 && ld % 2 == 0) o.vec = 2;
  (void)contiguous_is_k;
  (void)extent_contig;
  return o;
}

int out_vec(const float* c, int64_t ldc) {
  uintptr_t a = (uintptr_t)c;
  if (a % 16 == 0 && ldc % 4 == 0) return 4;
  if (a % 8 == 0 && ldc % 2 == 0) return 2;
  return 1;
}

int wgrad_splits(int64_t M, int32_t Nout, int32_t K) {
  const int64_t tiles = (int64_t)((Nout + BM - 1) / BM) * ((K + BN - 1) / BN);
  int64_t want = (2 * (int64_t)sm_count() + tiles - 1) / tiles;
  int64_t max_by_len = M / (4 * BK);  // at least 4 k-tiles per split
  if (max_by_len < 1) max_by_len = 1;
  if (want > max_by_len) want = max_by_len;
  if (want > 64) want = 64;
  if (want < 1) want = 1;
  return (int)want;
}


// ---- tensor-core path selection -------------------------------------------------------------------
bool tc_enabled() {
  const char* e = std::getenv("MGS_DISABLE_TC");
  return !(e && e[0] && e[0] != '0');
}
bool tc_applicable(int64_t M, int N, int Ktot) { return tc_enabled() && M >= 128 && N >= 16 && Ktot >= 8; }

int tc_pick_bn(int N) {
  if (const char* e = std::getenv("MGS_TC_BN")) {     // tuning experiments
    const int v = std::atoi(e);
    if (v == 128 || v == 176 || v == 256) return v;
  }
  // padded width x measured cost per output column (B200, M = 130k, K = 350: BN = 256 5.3e-4 ms, 176 6.5e-4, 128
  // 7.6e-4 -- the activation path is paid per tile, a wider tile amortises it): N = 350 -> 176, 700 / 1500 -> 256
  int best = 128;
  double best_cost = 1e30;
  const int cand[3] = {128, 176, 256};
  const double per_col[3] = {1.45, 1.24, 1.0};
  for (int i = 0; i < 3; ++i) {
    const double cost = (double)((N + cand[i] - 1) / cand[i] * cand[i]) * per_col[i];
    if (cost <= best_cost) { best_cost = cost; best = cand[i]; }
  }
  return best;
}

tc::Operand tc_operand(const float* p, int64_t ld, bool k_contig) {
  tc::Operand o;
  o.p = p;
  o.ld = ld;
  const uintptr_t a = (uintptr_t)p;
  o.vec = (a % 16 == 0 && ld % 4 == 0) ? 4 : ((a % 8 == 0 && ld % 2 == 0) ? 2 : 1);
  o.k_contig = k_contig ? 1 : 0;
  return o;
}

size_t tc_packed_bytes(int N, int K0, int K1) {
  const int bn = tc_pick_bn(N);
  const int64_t ntiles = (N + bn - 1) / bn;
  const int64_t nkb = (K0 + tc::BK - 1) / tc::BK + (K1 + tc::BK - 1) / tc::BK;
  return (size_t)(ntiles * nkb * 2 * bn * tc::kRowBytes);
}

constexpr int kTcCluster = 4;   // CTAs per cluster sharing (multicasting) one weight tile
// Measured on B200 (model1 SAGE projection): the GEMM is not L2-bandwidth bound, and the cluster-wide stage
// release makes 4 CTAs advance in lockstep (0.57 -> 0.66 ms).  Kept for wider layers, off by default.
constexpr bool kTcUseCluster = false;

template <int BN, bool PACKED, int CL, int VEC>
int tc_launch_one(const tc::Segment& s0, const tc::Segment& s1, const uint8_t* packed, int M, int N, float* c,
                  int64_t ldc, const float* bias, int relu, int splits, int k_per_split, int64_t split_stride,
                  cudaStream_t stream) {
  using C = tc::Cfg<BN, PACKED>;
  const int mtiles = (M + tc::BM - 1) / tc::BM;
  auto kern = tc::tc_gemm_kernel<BN, PACKED, CL, VEC>;
  MGS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((N + BN - 1) / BN, (mtiles + CL - 1) / CL * CL, splits);
  cfg.blockDim = dim3(tc::kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = CL;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  const char* dbg_env = std::getenv("MGS_TC_DEBUG");             // timing experiments only (results are garbage)
  const int dbg = dbg_env ? std::atoi(dbg_env) : 0;
  MGS_CUDA(cudaLaunchKernelEx(&cfg, kern, s0, s1, packed, M, N, c, ldc, out_vec(c, ldc), bias, relu, k_per_split,
                              split_stride, dbg));
  return check_launch("tc_gemm_kernel");
}

// one CTA per SM walking the tile list (tc_gemm_persistent_kernel): forward and dgrad with packed weights
template <int BN, int VEC>
int tc_launch_persistent(const tc::Segment& s0, const tc::Segment& s1, const uint8_t* packed, int M, int N, float* c,
                         int64_t ldc, const float* bias, int relu, cudaStream_t stream) {
  auto kern = tc::tc_gemm_persistent_kernel<BN, VEC>;
  constexpr int smem = tc::Cfg<BN, true>::kSmemBytes;
  MGS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int64_t tiles = (int64_t)((N + BN - 1) / BN) * ((M + tc::BM - 1) / tc::BM);
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  kern<<<grid, tc::kThreads, smem, stream>>>(s0, s1, packed, M, N, c, ldc, out_vec(c, ldc), bias, relu);
  return check_launch("tc_gemm_persistent_kernel");
}

template <int BN, bool PACKED>
int tc_launch_bn(const tc::Segment& s0, const tc::Segment& s1, const uint8_t* packed, int M, int N, float* c,
                 int64_t ldc, const float* bias, int relu, int splits, int k_per_split, int64_t split_stride,
                 cudaStream_t stream) {
  if constexpr (PACKED) {
    // packed kernels require K-contiguous activations; both segments share one compile-time vector width
    int vec = s0.a.vec;
    if (s1.K > 0 && s1.a.vec < vec) vec = s1.a.vec;
#define MGS_GO(CLV, VV) tc_launch_one<BN, true, CLV, VV>(s0, s1, packed, M, N, c, ldc, bias, relu, splits, k_per_split, \
                                                         split_stride, stream)
    if (kTcUseCluster && (M + tc::BM - 1) / tc::BM >= 2 * kTcCluster) {
      return vec == 4 ? MGS_GO(kTcCluster, 4) : vec == 2 ? MGS_GO(kTcCluster, 2) : MGS_GO(kTcCluster, 1);
    }
    const char* pe = std::getenv("MGS_TC_PERSISTENT");            // read per call: tests / probes toggle it
    const bool persistent = !(pe && pe[0] == '0');
    if (persistent && splits == 1 && s0.K > 0 && s0.a.k_contig && (s1.K == 0 || s1.a.k_contig)) {
      return vec == 4 ? tc_launch_persistent<BN, 4>(s0, s1, packed, M, N, c, ldc, bias, relu, stream)
           : vec == 2 ? tc_launch_persistent<BN, 2>(s0, s1, packed, M, N, c, ldc, bias, relu, stream)
                      : tc_launch_persistent<BN, 1>(s0, s1, packed, M, N, c, ldc, bias, relu, stream);
    }
    return vec == 4 ? MGS_GO(1, 4) : vec == 2 ? MGS_GO(1, 2) : MGS_GO(1, 1);
#undef MGS_GO
  } else {
    return tc_launch_one<BN, false, 1, 1>(s0, s1, packed, M, N, c, ldc, bias, relu, splits, k_per_split, split_stride,
                                          stream);
  }
}

// `packed_ws` != nullptr: B is a weight matrix -> pack it once (tc_pack_b_kernel), stream it with bulk copies.
int tc_launch(const tc::Segment& s0, const tc::Segment& s1, void* packed_ws, int M, int N, float* c, int64_t ldc,
              const float* bias, int relu, int splits, int k_per_split, int64_t split_stride, cudaStream_t stream) {
  const int bn = tc_pick_bn(N);
  const uint8_t* packed = (const uint8_t*)packed_ws;
  if (packed_ws != nullptr) {
    const int64_t chunks = (int64_t)tc_packed_bytes(N, s0.K, s1.K) / 32;  // one thread per (hi, lo) chunk pair
    tc::tc_pack_b_kernel<<<grid_for(chunks, 256, 8), 256, 0, stream>>>(s0.b, s0.K, s1.b, s1.K, N, bn,
                                                                       (uint8_t*)packed_ws);
    if (int rc = check_launch("tc_pack_b_kernel")) return rc;
  }
#define MGS_TC(BNV)                                                                                              \
  (packed ? tc_launch_bn<BNV, true>(s0, s1, packed, M, N, c, ldc, bias, relu, splits, k_per_split, split_stride, stream) \
          : tc_launch_bn<BNV, false>(s0, s1, packed, M, N, c, ldc, bias, relu, splits, k_per_split, split_stride, stream))
  switch (bn) {
    case 128: return MGS_TC(128);
    case 176: return MGS_TC(176);
    default:  return MGS_TC(256);
  }
#undef MGS_TC
}

// ---- TMA-fed kernel (tc_tma.cuh): forward / dgrad with 16-byte-aligned activation rows ---------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  // resolved through the runtime: libmgs.so does not link libcuda
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return (EncodeTiledFn)p;
  }();
  return fn;
}

bool tma_enabled() {
  const char* e = std::getenv("MGS_TC_TMA");                      // read per call: tests / probes toggle it
  return !(e && e[0] == '0');
}
bool tma_operand_ok(const float* p, int64_t ld) { return p != nullptr && ((uintptr_t)p & 15u) == 0 && ld % 4 == 0; }

// fp32 [rows, K] row-major with leading dimension ld -> boxes of 16 floats x 128 rows, SWIZZLE_64B, zero fill
bool make_act_map(CUtensorMap* map, const float* p, int64_t ld, int64_t rows, int K) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)tc::BK, (cuuint32_t)tc::BM};
  const cuuint32_t estride[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)p, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

constexpr int64_t kFewRows = 16384;   // readout-MLP class GEMMs (one row per molecule): tile choice by wave count, not width
int tma_pick_bn(int N, int64_t M) {
  if (const char* e = std::getenv("MGS_TMA_BN")) {
    const int v = std::atoi(e);
    if (v == 128 || v == 176 || v == 256) return v;
  }
  // measured at M = 4096 (tools/mlp_gemm_probe.py): [700 -> 1500] 0.067 ms on the cp.async kernel's 256-wide tiles (192 tiles =
  // 1.3 waves), 0.057 on 176-wide TMA tiles (288 = 1.95 waves); [1500 -> 128] 0.052 -> 0.029 with 128-wide tiles x 4 K splits
  if (M <= kFewRows) return N <= 128 ? 128 : 176;
  // tensor-bound kernel: cost ~ padded width, plus the per-N-tile re-read of the activation tile
  int best = 128;
  int64_t best_cost = INT64_MAX;
  const int cand[3] = {128, 176, 256};
  for (int i = 0; i < 3; ++i) {
    const int64_t cost = (int64_t)((N + cand[i] - 1) / cand[i]) * (cand[i] + 48);
    if (cost < best_cost) { best_cost = cost; best = cand[i]; }
  }
  return best;
}

// K split for GEMMs with fewer output tiles than SMs (readout MLP: 4096 x 128 is 32 tiles on 148 SMs)
int tma_splits(int64_t M, int N, int K0, int K1, int bn) {
  if (K1 > 0) return 1;
  // measured on B200 (tools/gemm_probe2.py): the partial-sum round trip and the extra reduction launch cost more than
  // the idle SMs of a 32-tile GEMM (4096 x 1500 -> 128: 0.032 ms unsplit, 0.038-0.051 ms split 2-4) -- off unless asked for
  const int64_t tiles = ((M + tc::BM - 1) / tc::BM) * ((N + bn - 1) / bn);
  const int nb = (K0 + tc::BK - 1) / tc::BK;
  const char* e = std::getenv("MGS_TMA_SPLITS");
  if (e && std::atoi(e) > 0) return std::atoi(e);
  if (!e) {
    // ... except for very long contractions on a handful of tiles (CNNNet.fc1 of train.py:133 at the script's batch of
    // 128: [128, 131072] x [131072, 256] is ONE 256-wide tile walking 8192 K blocks -- 8.2 instead of 4.3 ms per step)
    if (K0 >= 4096 && tiles * 2 <= sm_count()) {
      int64_t s = sm_count() / tiles;
      if (s > nb / 32) s = nb / 32;
      return s < 2 ? 1 : (int)(s > 148 ? 148 : s);
    }
    // ... and for few-row GEMMs that fill a quarter of the machine (4096 x 1500 -> 128 is 32 tiles: 0.052 -> 0.029 ms with 4
    // splits on this kernel with tensor-memory operands; the earlier measurement above was the first TMA kernel)
    if (M <= kFewRows && tiles * 4 <= sm_count() && nb >= 32) {
      int64_t s = sm_count() / tiles;
      if (s > 4) s = 4;
      if (s > nb / 8) s = nb / 8;
      return s < 2 ? 1 : (int)s;
    }
    return 1;
  }
  int64_t s = sm_count() / tiles;
  if (s > nb / 8) s = nb / 8;                    // at least 8 K blocks (128 contraction steps) per split
  if (s > 16) s = 16;
  return s < 2 ? 1 : (int)s;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t tma_workspace_bytes(int64_t M, int N, int K0, int K1) {
  const int bn = tma_pick_bn(N, M);
  const int64_t nkb = (K0 + tc::BK - 1) / tc::BK + (K1 + tc::BK - 1) / tc::BK;
  size_t bytes = align_up((size_t)((N + bn - 1) / bn) * nkb * bn * tc::kRowBytes, 1024);
  const int splits = tma_splits(M, N, K0, K1, bn);
  if (splits > 1) bytes += sizeof(float) * (size_t)splits * M * N;
  return bytes;
}

struct MaskBits {                 // ReLU-backward mask applied by the epilogue (tc_tma.cuh: mask_by_bits); bits == nullptr: none
  const uint32_t* bits;
  int words, v;
  float* colsum_part = nullptr;   // CTA-pair kernel only: per-32-row column sums of the output ([ceil(M / 256) * 8][N])
};

template <int BN, bool TS>
int tma_launch_bn(const CUtensorMap& m0, const CUtensorMap& m1, const tc::Segment& s0, const tc::Segment& s1,
                  const uint8_t* packed, int M, int N,
                  float* c, int64_t ldc, const float* bias, int relu, int splits, int64_t split_stride,
                  cudaStream_t stream, const MaskBits& mb) {
  auto kern = tma::gemm_tma_kernel<BN, TS>;
  constexpr int smem = tma::Cfg<BN, TS>::kSmemBytes;
  MGS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int64_t tiles = (int64_t)((N + BN - 1) / BN) * ((M + tc::BM - 1) / tc::BM) * splits;
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  const char* dbg_env = std::getenv("MGS_TMA_DEBUG");             // timing experiments only (results are garbage)
  const int dbg = dbg_env ? std::atoi(dbg_env) : 0;
  kern<<<grid, tc::kThreads, smem, stream>>>(m0, m1, s0.K, s1.K, s0.a.p, s0.a.ld, s1.K > 0 ? s1.a.p : s0.a.p,
                                             s1.K > 0 ? s1.a.ld : s0.a.ld, packed, M, N, c, ldc, bias, relu, splits,
                                             split_stride, dbg, mb.bits, mb.words, mb.v);
  return check_launch("gemm_tma_kernel");
}

// CTA-pair kernel (tcgen05 cta_group::2): clusters of two CTAs, one 256 x BN tile per pair
template <int BN>
int tma2_launch_bn(const CUtensorMap& m0, const CUtensorMap& m1, const tc::Segment& s0, const tc::Segment& s1,
                   const uint8_t* packed, int M, int N, float* c, int64_t ldc, const float* bias, int relu, int splits,
                   int64_t split_stride, cudaStream_t stream, const MaskBits& mb) {
  auto kern = tma::gemm_tma2_kernel<BN>;
  constexpr int smem = tma::Cfg2<BN>::kSmemBytes;
  MGS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int64_t pairs = (int64_t)((N + BN - 1) / BN) * ((M + 2 * tc::BM - 1) / (2 * tc::BM)) * splits;
  const int64_t max_pairs = sm_count() / 2;
  const int grid = 2 * (int)(pairs < max_pairs ? pairs : max_pairs);
  const char* dbg_env = std::getenv("MGS_TMA_DEBUG");             // timing experiments only
  const int dbg = dbg_env ? std::atoi(dbg_env) : 0;
  kern<<<grid, tc::kThreads, smem, stream>>>(m0, m1, s0.K, s1.K, s0.a.p, s0.a.ld, s1.K > 0 ? s1.a.p : s0.a.p,
                                             s1.K > 0 ? s1.a.ld : s0.a.ld, packed, M, N, c, ldc, bias, relu, splits,
                                             split_stride, dbg, mb.bits, mb.words, mb.v, mb.colsum_part);
  return check_launch("gemm_tma2_kernel");
}

// returns MGS_OK after launching, or -1 when this call cannot take the TMA kernel (caller falls back)
int tma_gemm(const tc::Segment& s0, const tc::Segment& s1, void* workspace, size_t workspace_bytes, int M, int N, float* c,
             int64_t ldc, const float* bias, int relu, cudaStream_t stream, const MaskBits mb = MaskBits{nullptr, 0, 0}) {
  if (!tma_enabled() || !tma_operand_ok(s0.a.p, s0.a.ld) || (s1.K > 0 && !tma_operand_ok(s1.a.p, s1.a.ld))) return -1;
  // Measured on B200 (tools/gemm_probe2.py, profiles/round2_gemm_probe.txt): both kernels are bound by shared-memory
  // bandwidth (UMMA operand reads + staging traffic), not by loads; this kernel wins where 176-wide tiles fit the output
  // (N = 350: 0.350 vs 0.376 ms, [130k,700]->350: 0.391 vs 0.437 ms), the cp.async kernel with its pre-split weights
  // wins on 256-wide tiles (N >= 705: 225 vs 194 TFLOP/s).  MGS_TC_TMA=2 forces this kernel for every shape.
  {
    const char* e = std::getenv("MGS_TC_TMA");
    const int pbn = tma_pick_bn(N, M);
    if (!(e && e[0] == '2') && M > kFewRows && pbn != 176 && tma_splits(M, N, s0.K, s1.K, pbn) == 1) return -1;
  }
  if (workspace == nullptr || workspace_bytes < tma_workspace_bytes(M, N, s0.K, s1.K)) return -1;
  CUtensorMap m0, m1;
  if (!make_act_map(&m0, s0.a.p, s0.a.ld, M, s0.K)) return -1;
  if (s1.K > 0) { if (!make_act_map(&m1, s1.a.p, s1.a.ld, M, s1.K)) return -1; }
  else m1 = m0;
  const int bn = tma_pick_bn(N, M);
  const int64_t nkb = (s0.K + tc::BK - 1) / tc::BK + (s1.K + tc::BK - 1) / tc::BK;
  const size_t packed_bytes = align_up((size_t)((N + bn - 1) / bn) * nkb * bn * tc::kRowBytes, 1024);
  uint8_t* packed = (uint8_t*)workspace;
  tma::pack_b_raw_kernel<<<grid_for((int64_t)packed_bytes / 16, 256, 8), 256, 0, stream>>>(s0.b, s0.K, s1.b, s1.K, N, bn, packed);
  if (int rc = check_launch("pack_b_raw_kernel")) return rc;
  const int splits = tma_splits(M, N, s0.K, s1.K, bn);
  if (mb.bits != nullptr && splits > 1) return -1;              // the mask lives in the GEMM epilogue only
  float* dst = c;
  int64_t dst_ld = ldc, stride = 0;
  if (splits > 1) {
    dst = (float*)(packed + packed_bytes);
    dst_ld = N;
    stride = (int64_t)M * N;
  }
  const float* kb = splits > 1 ? nullptr : bias;
  const int kr = splits > 1 ? 0 : relu;
  int rc;
  const char* ts_env = std::getenv("MGS_TMA_TS");                  // 0: activation operand from shared memory (SS form)
  const bool ts = !(ts_env && ts_env[0] == '0');
  const char* pair_env = std::getenv("MGS_TMA_2CTA");              // 0: one CTA per tile (cta_group::1)
  // (split few-row GEMMs: one CTA per work item spreads 128 items better than 64 pairs: 0.029 vs 0.031 ms)
  const bool pair = ts && !(pair_env && pair_env[0] == '0') && M > tc::BM && !(splits > 1 && M <= kFewRows && !pair_env);
  if (mb.colsum_part != nullptr && !(pair && (bn == 128 || bn == 176) && splits == 1)) return -1;
  if (pair && bn == 128) {
    rc = tma2_launch_bn<128>(m0, m1, s0, s1, packed, M, N, dst, dst_ld, kb, kr, splits, stride, stream, mb);
  } else if (pair && bn == 176) {
    rc = tma2_launch_bn<176>(m0, m1, s0, s1, packed, M, N, dst, dst_ld, kb, kr, splits, stride, stream, mb);
  } else
  switch (bn) {
    case 128:
      rc = ts ? tma_launch_bn<128, true>(m0, m1, s0, s1, packed, M, N, dst, dst_ld, kb, kr, splits, stride, stream, mb)
              : tma_launch_bn<128, false>(m0, m1, s0, s1, packed, M, N, dst, dst_ld, kb, kr, splits, stride, stream, mb);
      break;
    case 176:
      rc = ts ? tma_launch_bn<176, true>(m0, m1, s0, s1, packed, M, N, dst, dst_ld, kb, kr, splits, stride, stream, mb)
              : tma_launch_bn<176, false>(m0, m1, s0, s1, packed, M, N, dst, dst_ld, kb, kr, splits, stride, stream, mb);
      break;
    default:  // two 256-column accumulators fill the tensor memory: shared-memory operands only
      rc = tma_launch_bn<256, false>(m0, m1, s0, s1, packed, M, N, dst, dst_ld, kb, kr, splits, stride, stream, mb);
      break;
  }
  if (rc != MGS_OK || splits == 1) return rc;
  splitk_reduce_kernel<<<grid_for(stride, 256, 8), 256, 0, stream>>>(dst, splits, stride, M, N, c, ldc, bias, relu);
  return check_launch("splitk_reduce_kernel");
}


// ---- weight gradient on the TMA path (tc_wgrad.cuh) -------------------------------------------------------------------
bool make_map_2d(CUtensorMap* map, const float* p, int64_t ld, int64_t inner, int64_t outer, int box_inner, int box_outer,
                 CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  const cuuint32_t estride[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)p, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
             swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct WgradTmaPlan {
  bool ok;
  int bn;
  int splits;
};
// One (tile, split) per CTA, every CTA resident: splits = SMs / tiles, at least 8 K blocks (128 atoms) per split.
WgradTmaPlan wgrad_tma_plan(int64_t M, int32_t Nout, int32_t K) {
  WgradTmaPlan p{false, 176, 1};
  const char* e = std::getenv("MGS_WGRAD_TMA");                   // read per call: tests / probes toggle it
  if (e && e[0] == '0') return p;
  if (!tc_enabled() || M < 1024 || Nout < 32 || K < 32) return p;
  if (K <= 48) {
    p.bn = 48;                                                     // GATConv(35, ...): two 32-channel boxes, N = 48 MMAs
  } else if (const char* b = std::getenv("MGS_WGRAD_BN")) {
    const int v = std::atoi(b);
    p.bn = v == 128 ? 128 : 176;
  } else {
    const int64_t c128 = (int64_t)((K + 127) / 128) * (128 + 48), c176 = (int64_t)((K + 175) / 176) * (176 + 48);
    p.bn = c128 < c176 ? 128 : 176;
  }
  const int64_t tiles = (int64_t)((Nout + tc::BM - 1) / tc::BM) * ((K + p.bn - 1) / p.bn);
  const int64_t nb = (M + tc::BK - 1) / tc::BK;
  int64_t s = sm_count() / tiles;
  // The tensor core's fp32 accumulation truncates: the error grows with the number of accumulations per split (measured
  // at 130 k atoms: 1.1e-5 of max |dw| with 24 splits of 340 K blocks, 5.8e-6 with 48 -- the class of the cp.async kernel's
  // 49).  Long contractions therefore run two rounds of work items per CTA (MGS_WGRAD_WAVES overrides).
  // (up to 8 rounds: the stress shape's 1.5 M atoms would otherwise accumulate 2 600 K blocks per split)
  int64_t waves = s >= 1 ? (nb / s + 255) / 256 : 1;
  if (waves < 1) waves = 1;
  if (waves > 8) waves = 8;
  if (const char* we = std::getenv("MGS_WGRAD_WAVES")) { if (std::atoi(we) > 0) waves = std::atoi(we); }
  s *= waves;
  if (const char* se = std::getenv("MGS_WGRAD_SPLITS")) { if (std::atoi(se) > 0) s = std::atoi(se); }
  if (s > nb / 8) s = nb / 8;
  if (s < 1) s = 1;
  p.splits = (int)s;
  p.ok = true;
  return p;
}

template <int BN>
int wgrad_tma_launch(const CUtensorMap& mg, const CUtensorMap& ma, int M, int Nout, int K, float* c, int64_t ldc, int splits,
                     int64_t stride, cudaStream_t stream) {
  auto kern = tma::gemm_tma_wgrad_kernel<BN>;
  constexpr int smem = tma::WCfg<BN>::kSmemBytes;
  MGS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int64_t tiles = (int64_t)((Nout + tc::BM - 1) / tc::BM) * ((K + BN - 1) / BN) * splits;
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  const char* dbg_env = std::getenv("MGS_TMA_DEBUG");             // timing experiments only (results are garbage)
  const int dbg = dbg_env ? std::atoi(dbg_env) : 0;
  kern<<<grid, tc::kThreads, smem, stream>>>(mg, ma, M, Nout, K, c, ldc, splits, stride, dbg);
  return check_launch("gemm_tma_wgrad_kernel");
}

// returns MGS_OK after launching, -1 when the call cannot take this kernel
int wgrad_tma(const float* g, int64_t ldg, int64_t M, int32_t Nout, const float* a, int64_t lda, int32_t K, float* dst,
              int64_t dst_ld, const WgradTmaPlan& plan, int64_t stride, cudaStream_t stream) {
  if (!tma_operand_ok(g, ldg) || !tma_operand_ok(a, lda)) return -1;
  CUtensorMap mg, ma;
  if (!make_map_2d(&mg, g, ldg, Nout, M, tc::BM, tc::BK, CU_TENSOR_MAP_SWIZZLE_NONE)) return -1;
  if (!make_map_2d(&ma, a, lda, K, M, tma::kBoxN, tc::BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return -1;
  if (plan.bn == 48) return wgrad_tma_launch<48>(mg, ma, (int)M, Nout, K, dst, dst_ld, plan.splits, stride, stream);
  if (plan.bn == 128) return wgrad_tma_launch<128>(mg, ma, (int)M, Nout, K, dst, dst_ld, plan.splits, stride, stream);
  return wgrad_tma_launch<176>(mg, ma, (int)M, Nout, K, dst, dst_ld, plan.splits, stride, stream);
}

struct WgradPlan {
  bool use_tc;
  int splits;
  int k_per_split;
};
// dw[Nout, K] = g^T a: output tiles are few (Nout, K ~ 350), the contraction (M rows) is long -> split it.
WgradPlan wgrad_plan(int64_t M, int32_t Nout, int32_t K) {
  WgradPlan p;
  p.use_tc = tc_applicable(Nout, K, (int)(M > 0x7fffffff ? 0x7fffffff : M)) && M >= 256;
  if (p.use_tc) {
    const int bn = tc_pick_bn(K);
    const int64_t tiles = (int64_t)((Nout + tc::BM - 1) / tc::BM) * ((K + bn - 1) / bn);
    // two waves of one-CTA-per-SM tiles; shorter splits also bound the truncating tensor-core accumulation.
    // Rounded DOWN: 6 tiles x 50 splits = 300 CTAs on 148 SMs ran a third wave of 4 CTAs (0.44 ms instead of 0.30)
    int64_t want = (2 * (int64_t)sm_count()) / tiles;
    int64_t max_by_len = M / (8 * tc::BK);
    if (max_by_len < 1) max_by_len = 1;
    if (want > max_by_len) want = max_by_len;
    if (want > 64) want = 64;
    if (want < 1) want = 1;
    p.splits = (int)want;
    int kps = (int)((M + p.splits - 1) / p.splits);
    p.k_per_split = (kps + tc::BK - 1) / tc::BK * tc::BK;
  } else {
    p.splits = wgrad_splits(M, Nout, K);
    int kps = (int)((M + p.splits - 1) / p.splits);
    p.k_per_split = (kps + BK - 1) / BK * BK;
  }
  return p;
}

// Few-row GEMMs on the FFMA kernel (readout MLP at the reference's own batch of 64 / 128 molecules: M < 128 rows never
// reaches the tensor-core path): 128 x 128 output tiles give 1-12 CTAs that each walk the whole contraction --
// 60-80 us per call on 148 SMs.  Split the contraction so that about two CTAs per SM are busy; partial tiles are
// summed in split order by splitk_reduce_kernel (which also applies the bias), so the result is deterministic.
struct SkinnyPlan {
  int splits;
  int k_per_split;
};
SkinnyPlan skinny_plan(int64_t M, int32_t N, int32_t K, bool single_segment) {
  SkinnyPlan p{1, 0};
  if (!single_segment) return p;
  const int64_t tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  int64_t want = (2 * (int64_t)sm_count()) / tiles;
  const int64_t max_by_len = K / (4 * BK);                 // at least 64 contraction steps per split
  if (want > max_by_len) want = max_by_len;
  if (want > 64) want = 64;
  if (want < 2) return p;
  int kps = (int)((K + want - 1) / want);
  kps = (kps + BK - 1) / BK * BK;
  p.splits = (K + kps - 1) / kps;
  p.k_per_split = kps;
  if (p.splits < 2) p = SkinnyPlan{1, 0};
  return p;
}

}  // namespace
}  // namespace mgs

using namespace mgs;

extern "C" size_t mgs_linear_fwd_workspace_bytes(int64_t M, int32_t K, int32_t Nout, int32_t K2) {
  if (M <= 0 || K <= 0 || Nout <= 0 || K2 < 0) return 0;
  if (tc_applicable(M, Nout, K + K2)) return std::max(tc_packed_bytes(Nout, K, K2), tma_workspace_bytes(M, Nout, K, K2));
  const SkinnyPlan sp = skinny_plan(M, Nout, K, K2 == 0);
  return sp.splits > 1 ? sizeof(float) * (size_t)sp.splits * M * Nout : 0;
}

extern "C" int mgs_linear_fwd(const float* a, int64_t lda, int64_t M, int32_t K, const float* w, int64_t ldw,
                              int32_t Nout, const float* bias, const float* a2, int64_t lda2, int32_t K2,
                              const float* w2, int64_t ldw2, float* c, int64_t ldc, int32_t relu,
                              void* workspace, size_t workspace_bytes, mgs_stream_t stream_) {
  MGS_REQUIRE(M >= 0 && M < 0x7fffffff && K > 0 && Nout > 0, "mgs_linear_fwd: bad sizes");
  MGS_REQUIRE(lda >= K && ldw >= K && ldc >= Nout, "mgs_linear_fwd: leading dimension too small");
  if (M == 0) return MGS_OK;
  MGS_REQUIRE(a && w && c, "mgs_linear_fwd: null pointer");
  if (a2 != nullptr)
    MGS_REQUIRE(w2 && K2 > 0 && lda2 >= K2 && ldw2 >= K2, "mgs_linear_fwd: bad second operand pair");
  const int k2 = a2 ? K2 : 0;
  if (Nout == 1 && a2 == nullptr) {                                  // one dot product per row
    gemv_rows_kernel<<<grid_for(M * 32, 256, 8), 256, 0, (cudaStream_t)stream_>>>(a, lda, (int)M, K, w, bias, relu, c, ldc);
    return check_launch("gemv_rows_kernel");
  }
  if (tc_applicable(M, Nout, K + k2)) {
    const size_t need = tc_packed_bytes(Nout, K, k2);
    if (workspace_bytes < need || !workspace) {
      set_error("mgs_linear_fwd: workspace too small (%zu < %zu)", workspace_bytes, need);
      return MGS_ERR_WORKSPACE_TOO_SMALL;
    }
    tc::Segment t0{tc_operand(a, lda, true), tc_operand(w, ldw, true), K};
    tc::Segment t1{tc::Operand{nullptr, 0, 1, 1}, tc::Operand{nullptr, 0, 1, 1}, 0};
    if (a2 != nullptr) t1 = tc::Segment{tc_operand(a2, lda2, true), tc_operand(w2, ldw2, true), K2};
    if (int rc = tma_gemm(t0, t1, workspace, workspace_bytes, (int)M, Nout, c, ldc, bias, relu, (cudaStream_t)stream_); rc >= 0)
      return rc;
    return tc_launch(t0, t1, workspace, (int)M, Nout, c, ldc, bias, relu, 1, 0, 0, (cudaStream_t)stream_);
  }
  Segment s0{make_operand(a, lda, true, K), make_operand(w, ldw, true, K), K};
  Segment s1{Operand{nullptr, 0, 1}, Operand{nullptr, 0, 1}, 0};
  if (a2 != nullptr) s1 = Segment{make_operand(a2, lda2, true, K2), make_operand(w2, ldw2, true, K2), K2};
  const SkinnyPlan sp = skinny_plan(M, Nout, K, a2 == nullptr);
  if (sp.splits > 1 && workspace && workspace_bytes >= sizeof(float) * (size_t)sp.splits * M * Nout) {
    float* part = (float*)workspace;
    const int64_t stride = (int64_t)M * Nout;
    dim3 grid((Nout + BN - 1) / BN, (unsigned)((M + BM - 1) / BM), sp.splits);
    gemm_kernel<true, true><<<grid, kThreads, 0, (cudaStream_t)stream_>>>(s0, s1, (int)M, Nout, part, Nout,
                                                                          out_vec(part, Nout), nullptr, 0,
                                                                          sp.k_per_split, stride);
    if (int rc = check_launch("gemm_kernel<NT>")) return rc;
    splitk_reduce_kernel<<<grid_for(stride, 256, 8), 256, 0, (cudaStream_t)stream_>>>(part, sp.splits, stride, (int)M,
                                                                                     Nout, c, ldc, bias, relu);
    return check_launch("splitk_reduce_kernel");
  }
  dim3 grid((Nout + BN - 1) / BN, (unsigned)((M + BM - 1) / BM), 1);
  gemm_kernel<true, true><<<grid, kThreads, 0, (cudaStream_t)stream_>>>(s0, s1, (int)M, Nout, c, ldc, out_vec(c, ldc),
                                                                        bias, relu, 0, 0);
  return check_launch("gemm_kernel<NT>");
}

extern "C" size_t mgs_linear_dgrad_workspace_bytes(int64_t M, int32_t Nout, int32_t K) {
  if (M <= 0 || K <= 0 || Nout <= 0) return 0;
  if (tc_applicable(M, K, Nout)) return std::max(tc_packed_bytes(K, Nout, 0), tma_workspace_bytes(M, K, Nout, 0));
  const SkinnyPlan sp = skinny_plan(M, K, Nout, true);
  return sp.splits > 1 ? sizeof(float) * (size_t)sp.splits * M * K : 0;
}

extern "C" int mgs_linear_dgrad(const float* g, int64_t ldg, int64_t M, int32_t Nout, const float* w, int64_t ldw,
                                int32_t K, float* da, int64_t ldda, void* workspace, size_t workspace_bytes,
                                mgs_stream_t stream_) {
  MGS_REQUIRE(M >= 0 && M < 0x7fffffff && K > 0 && Nout > 0, "mgs_linear_dgrad: bad sizes");
  MGS_REQUIRE(ldg >= Nout && ldw >= K && ldda >= K, "mgs_linear_dgrad: leading dimension too small");
  if (M == 0) return MGS_OK;
  MGS_REQUIRE(g && w && da, "mgs_linear_dgrad: null pointer");
  // da[m][k'] = sum_n g[m][n] * w[n][k']  ->  A = g (contraction contiguous), B(k=n, n'=k') = w[n*ldw + k']
  if (Nout == 1) {                                                   // outer product g w
    outer_rows_kernel<<<grid_for(M * K, 256, 8), 256, 0, (cudaStream_t)stream_>>>(g, ldg, (int)M, K, w, da, ldda);
    return check_launch("outer_rows_kernel");
  }
  if (tc_applicable(M, K, Nout)) {
    const size_t need = tc_packed_bytes(K, Nout, 0);
    if (workspace_bytes < need || !workspace) {
      set_error("mgs_linear_dgrad: workspace too small (%zu < %zu)", workspace_bytes, need);
      return MGS_ERR_WORKSPACE_TOO_SMALL;
    }
    tc::Segment t0{tc_operand(g, ldg, true), tc_operand(w, ldw, false), Nout};
    tc::Segment t1{tc::Operand{nullptr, 0, 1, 1}, tc::Operand{nullptr, 0, 1, 1}, 0};
    if (int rc = tma_gemm(t0, t1, workspace, workspace_bytes, (int)M, K, da, ldda, nullptr, 0, (cudaStream_t)stream_); rc >= 0)
      return rc;
    return tc_launch(t0, t1, workspace, (int)M, K, da, ldda, nullptr, 0, 1, 0, 0, (cudaStream_t)stream_);
  }
  Segment s0{make_operand(g, ldg, true, Nout), make_operand(w, ldw, false, K), Nout};
  Segment s1{Operand{nullptr, 0, 1}, Operand{nullptr, 0, 1}, 0};
  const SkinnyPlan sp = skinny_plan(M, K, Nout, true);
  if (sp.splits > 1 && workspace && workspace_bytes >= sizeof(float) * (size_t)sp.splits * M * K) {
    float* part = (float*)workspace;
    const int64_t stride = (int64_t)M * K;
    dim3 grid((K + BN - 1) / BN, (unsigned)((M + BM - 1) / BM), sp.splits);
    gemm_kernel<true, false><<<grid, kThreads, 0, (cudaStream_t)stream_>>>(s0, s1, (int)M, K, part, K, out_vec(part, K),
                                                                           nullptr, 0, sp.k_per_split, stride);
    if (int rc = check_launch("gemm_kernel<NN>")) return rc;
    splitk_reduce_kernel<<<grid_for(stride, 256, 8), 256, 0, (cudaStream_t)stream_>>>(part, sp.splits, stride, (int)M, K,
                                                                                     da, ldda);
    return check_launch("splitk_reduce_kernel");
  }
  dim3 grid((K + BN - 1) / BN, (unsigned)((M + BM - 1) / BM), 1);
  gemm_kernel<true, false><<<grid, kThreads, 0, (cudaStream_t)stream_>>>(s0, s1, (int)M, K, da, ldda,
                                                                         out_vec(da, ldda), nullptr, 0, 0, 0);
  return check_launch("gemm_kernel<NN>");
}

// da = g0 w0 + g1 w1 (both data-gradient form: w_i is [Nout_i, K] row-major, contraction over its rows), optionally masked by
// the bits of a fused ReLU (SAGEConv backward: gx = relu'(x) * (g W_r + (A^T D^-1 g) W_l) as ONE K = 700 GEMM whose
// epilogue applies the mask -- the [N, 700] intermediate of the two-column-block formulation never exists).  TMA kernel only:
// MGS_ERR_UNSUPPORTED when the operands do not qualify (the caller keeps its unfused path).
static size_t dgrad2_colsum_rows(int64_t M) { return (size_t)((M + 255) / 256) * 8; }
extern "C" size_t mgs_linear_dgrad2_workspace_bytes(int64_t M, int32_t N0, int32_t N1, int32_t K) {
  if (M <= 0 || K <= 0 || N0 <= 0 || N1 < 0) return 0;
  return align_up(tma_workspace_bytes(M, K, N0, N1), 1024) + sizeof(float) * dgrad2_colsum_rows(M) * K;
}

extern "C" int mgs_linear_dgrad2(const float* g0, int64_t ldg0, int32_t N0, const float* w0, int64_t ldw0, const float* g1,
                                 int64_t ldg1, int32_t N1, const float* w1, int64_t ldw1, int64_t M, int32_t K, float* da,
                                 int64_t ldda, const uint32_t* relu_bits, int32_t bits_words, int32_t bits_v,
                                 float* colsum_out, void* workspace, size_t workspace_bytes, mgs_stream_t stream_) {
  MGS_REQUIRE(M >= 0 && M < 0x7fffffff && K > 0 && N0 > 0 && N1 >= 0, "mgs_linear_dgrad2: bad sizes");
  MGS_REQUIRE(ldg0 >= N0 && ldw0 >= K && ldda >= K && (N1 == 0 || (ldg1 >= N1 && ldw1 >= K)),
              "mgs_linear_dgrad2: leading dimension too small");
  MGS_REQUIRE(relu_bits == nullptr || ((bits_v == 1 || bits_v == 2 || bits_v == 4) && bits_words > 0),
              "mgs_linear_dgrad2: bad mask layout");
  if (M == 0) return MGS_OK;
  MGS_REQUIRE(g0 && w0 && da && (N1 == 0 || (g1 && w1)), "mgs_linear_dgrad2: null pointer");
  if (!tc_applicable(M, K, N0 + N1)) {
    set_error("mgs_linear_dgrad2: shape not on the tensor-core path");
    return MGS_ERR_UNSUPPORTED;
  }
  tc::Segment t0{tc_operand(g0, ldg0, true), tc_operand(w0, ldw0, false), N0};
  tc::Segment t1{tc::Operand{nullptr, 0, 1, 1}, tc::Operand{nullptr, 0, 1, 1}, 0};
  if (N1 > 0) t1 = tc::Segment{tc_operand(g1, ldg1, true), tc_operand(w1, ldw1, false), N1};
  MaskBits mb{relu_bits, bits_words, bits_v, nullptr};
  const size_t gemm_bytes = align_up(tma_workspace_bytes(M, K, N0, N1), 1024);
  if (colsum_out != nullptr) {
    if (!workspace || workspace_bytes < gemm_bytes + sizeof(float) * dgrad2_colsum_rows(M) * K) {
      set_error("mgs_linear_dgrad2: workspace too small for the column sums");
      return MGS_ERR_WORKSPACE_TOO_SMALL;
    }
    mb.colsum_part = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + gemm_bytes);
  }
  const int rc = tma_gemm(t0, t1, workspace, workspace_bytes < gemm_bytes ? workspace_bytes : gemm_bytes, (int)M, K, da, ldda,
                          nullptr, 0, (cudaStream_t)stream_, mb);
  if (rc == MGS_OK && colsum_out != nullptr) {
    colsum_final_kernel<<<(K + 31) / 32, 1024, 0, (cudaStream_t)stream_>>>(mb.colsum_part, (int)dgrad2_colsum_rows(M), K,
                                                                          colsum_out);
    return check_launch("colsum_final_kernel");
  }
  if (rc >= 0) return rc;
  set_error("mgs_linear_dgrad2: operands do not qualify for the TMA kernel (alignment, tile width or workspace)");
  return MGS_ERR_UNSUPPORTED;
}

extern "C" size_t mgs_linear_wgrad_workspace_bytes(int64_t M, int32_t Nout, int32_t K) {
  if (M <= 0 || Nout <= 0 || K <= 0) return 0;
  const WgradPlan plan = wgrad_plan(M, Nout, K);
  const WgradTmaPlan tp = wgrad_tma_plan(M, Nout, K);
  const size_t old_bytes = plan.splits > 1 ? sizeof(float) * (size_t)plan.splits * Nout * K : 0;
  const size_t tma_bytes = (tp.ok && tp.splits > 1) ? sizeof(float) * (size_t)tp.splits * Nout * ((K + 3) & ~3) : 0;
  const size_t wsum_bytes = Nout == 1 ? sizeof(float) * (size_t)kWsumParts * K : 0;
  return std::max(std::max(old_bytes, tma_bytes), wsum_bytes);
}

extern "C" int mgs_linear_wgrad(const float* g, int64_t ldg, int64_t M, int32_t Nout, const float* a, int64_t lda,
                                int32_t K, float* dw, int64_t lddw, void* workspace, size_t workspace_bytes,
                                mgs_stream_t stream_) {
  MGS_REQUIRE(M >= 0 && M < 0x7fffffff && K > 0 && Nout > 0, "mgs_linear_wgrad: bad sizes");
  MGS_REQUIRE(ldg >= Nout && lda >= K && lddw >= K, "mgs_linear_wgrad: leading dimension too small");
  MGS_REQUIRE(dw, "mgs_linear_wgrad: null output");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (M == 0) {
    MGS_CUDA(cudaMemset2DAsync(dw, sizeof(float) * lddw, 0, sizeof(float) * K, Nout, stream));
    return MGS_OK;
  }
  MGS_REQUIRE(g && a, "mgs_linear_wgrad: null pointer");
  // dw[o][i] = sum_r g[r][o] * a[r][i]  ->  A(m=o,k=r) = g[r*ldg + o], B(k=r,n=i) = a[r*lda + i]
  if (Nout == 1 && workspace && workspace_bytes >= sizeof(float) * (size_t)kWsumParts * K) {   // weighted column sum
    const int parts = (int)(M < kWsumParts ? M : kWsumParts);
    wsum_partial_kernel<<<parts, 256, 0, stream>>>(g, ldg, a, lda, (int)M, K, (float*)workspace);
    if (int rc = check_launch("wsum_partial_kernel")) return rc;
    colsum_final_kernel<<<(K + 31) / 32, 1024, 0, stream>>>((const float*)workspace, parts, K, dw);
    return check_launch("colsum_final_kernel");
  }
  {
    const WgradTmaPlan tp = wgrad_tma_plan(M, Nout, K);               // TMA-fed kernel (tc_wgrad.cuh) where it applies
    const int ldp = (K + 3) & ~3;                                     // partial tiles: rows 16-byte aligned (128-bit stores)
    const size_t need = tp.splits > 1 ? sizeof(float) * (size_t)tp.splits * Nout * ldp : 0;
    if (tp.ok && (need == 0 || (workspace && workspace_bytes >= need && ((uintptr_t)workspace & 15u) == 0))) {
      const int64_t tstride = (int64_t)Nout * ldp;
      float* tdst = tp.splits > 1 ? (float*)workspace : dw;
      const int rc = wgrad_tma(g, ldg, M, Nout, a, lda, K, tdst, tp.splits > 1 ? ldp : lddw, tp, tstride, stream);
      if (rc == MGS_OK) {
        if (tp.splits == 1) return MGS_OK;
        splitk_reduce_kernel<<<grid_for((int64_t)Nout * K, 256, 8), 256, 0, stream>>>(tdst, tp.splits, tstride, Nout, K, dw,
                                                                                      lddw, nullptr, 0, ldp);
        return check_launch("splitk_reduce_kernel");
      }
      if (rc != -1) return rc;
    }
  }
  const WgradPlan plan = wgrad_plan(M, Nout, K);
  const int splits = plan.splits;
  const int64_t stride = (int64_t)Nout * K;
  float* part = (float*)workspace;
  if (splits > 1) {
    const size_t need = sizeof(float) * (size_t)splits * Nout * K;
    if (workspace_bytes < need || !workspace) {
      set_error("mgs_linear_wgrad: workspace too small (%zu < %zu)", workspace_bytes, need);
      return MGS_ERR_WORKSPACE_TOO_SMALL;
    }
  }
  float* dst = splits > 1 ? part : dw;
  const int64_t dst_ld = splits > 1 ? K : lddw;
  if (plan.use_tc) {
    tc::Segment t0{tc_operand(g, ldg, false), tc_operand(a, lda, false), (int)M};
    tc::Segment t1{tc::Operand{nullptr, 0, 1, 1}, tc::Operand{nullptr, 0, 1, 1}, 0};
    if (int rc = tc_launch(t0, t1, nullptr, Nout, K, dst, dst_ld, nullptr, 0, splits, plan.k_per_split, stride, stream)) return rc;
  } else {
    Segment s0{make_operand(g, ldg, false, Nout), make_operand(a, lda, false, K), (int)M};
    Segment s1{Operand{nullptr, 0, 1}, Operand{nullptr, 0, 1}, 0};
    dim3 grid((K + BN - 1) / BN, (Nout + BM - 1) / BM, splits);
    gemm_kernel<false, false><<<grid, kThreads, 0, stream>>>(s0, s1, Nout, K, dst, dst_ld, out_vec(dst, dst_ld), nullptr,
                                                             0, plan.k_per_split, stride);
    if (int rc = check_launch("gemm_kernel<TN>")) return rc;
  }
  if (splits == 1) return MGS_OK;
  splitk_reduce_kernel<<<grid_for(stride, 256, 8), 256, 0, stream>>>(part, splits, stride, Nout, K, dw, lddw);
  return check_launch("splitk_reduce_kernel");
}

extern "C" size_t mgs_colsum_workspace_bytes(int32_t Nout) {
  return Nout > 0 ? sizeof(float) * (size_t)kColParts * Nout : 0;
}

extern "C" int mgs_colsum(const float* g, int64_t ldg, int64_t M, int32_t Nout, float* out, void* workspace,
                          size_t workspace_bytes, mgs_stream_t stream_) {
  MGS_REQUIRE(M >= 0 && M < 0x7fffffff && Nout > 0 && ldg >= Nout, "mgs_colsum: bad sizes");
  MGS_REQUIRE(out, "mgs_colsum: null output");
  if (workspace_bytes < mgs_colsum_workspace_bytes(Nout) || !workspace) {
    set_error("mgs_colsum: workspace too small");
    return MGS_ERR_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  int parts = (int)((M + kColWarps - 1) / kColWarps);
  if (parts > kColParts) parts = kColParts;
  if (parts < 1) parts = 1;
  const int V = vec_width(g, ldg, Nout);
  const int chunks = Nout / V;
  const int iters = iters_for(chunks);
  const size_t smem = sizeof(float) * (size_t)kColWarps * Nout;
  if (iters > 0 && smem <= 48 * 1024) {
#define MGS_L(VV, II) colsum_partial_row_kernel<VV, II><<<parts, kColWarps * 32, smem, stream>>>( \
      g, ldg, (int)M, Nout, chunks, (float*)workspace)
    MGS_DISPATCH_V_ITERS(V, iters, MGS_L);
#undef MGS_L
    if (int rc = check_launch("colsum_partial_row_kernel")) return rc;
  } else {
    colsum_partial_kernel<<<parts, 256, 0, stream>>>(g, ldg, (int)M, Nout, (float*)workspace);
    if (int rc = check_launch("colsum_partial_kernel")) return rc;
  }
  colsum_final_kernel<<<(Nout + 31) / 32, 1024, 0, stream>>>((const float*)workspace, parts, Nout, out);
  return check_launch("colsum_final_kernel");
}
