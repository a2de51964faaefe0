// K4 (small-K variant) -- the GATConv input projection of the reference models (ablation/model1.py:57,
// GATConv(35, 35, heads=10): x[N,35] -> xh[N,350]) fused with the attention scores.
//
//   forward   [xh | a_src | a_dst] = x . [W ; U_src ; U_dst]^T      U_src[h,:] = sum_c att_src[h,c] W[hC+c,:]
//   wgrad     [dW ; dU_src ; dU_dst] = [dxh | da_src | da_dst]^T . x
//
// a_src[n,h] = <xh[n,h,:], att_src[h,:]> = <x[n,:], U_src[h,:]>: computing the scores from the 35 input features
// instead of the 350 projected ones removes a full re-read of xh (gat_scores: 0.10 ms) and, in the backward,
// the att_src / att_dst reduction over xh (gat_bwd_att: 0.085 ms) and the rank-1 updates of dxh; U and its
// gradient are 10 x 35 and stay PyTorch ops in the host layer.
//
// With K = 35 the contraction is 9 float4s: these are HBM-bound (write / read the [N,350] side once), not
// tensor-core work -- the generic 128 x 128 x 16 FFMA tile padded K to 48 and N to 128 and ran at 13 - 24 % of
// HBM peak.  Here the whole weight block (<= 96 KB) lives in shared memory, K is never padded beyond 4, and:
//   forward: CTA = 32 rows x all columns per step, thread = 8 rows x 3 column pairs (conflict-free 64-bit LDS,
//            coalesced 64-bit stores), x tile broadcast from shared memory;
//   wgrad:   thread = 2 output features x all K inputs (<= 72 accumulators), the atoms are split over 2 CTAs
//            per SM, x rows broadcast from shared memory, g read once with coalesced loads (next chunk in
//            flight during the FMAs), deterministic two-stage reduction.
// Both are FP32-issue bound (ncu: 104 M / 86 M warp instructions, half of them FFMA): the accumulators are
// packed pairs and every multiply-add is an FFMA2 (fma.rn.f32x2), which halves the dominant term.
#include <cstdlib>

#include "common.cuh"
#include "stream.cuh"

namespace mgs {
namespace {

constexpr int kFwdThreads = 256;
constexpr int kFwdRows = 32;
constexpr int kFwdColsPerThread = 6;            // columns tx, tx + 64, ...: up to 384 outputs
constexpr int kMaxOut = 64 * kFwdColsPerThread;
constexpr int kMaxK = 64;

struct OutSeg {
  float* p[3];
  int64_t ld[3];
  int n[3];       // columns of each segment; sum <= kMaxOut
};
struct InSeg {
  const float* p[3];
  int64_t ld[3];
  int n[3];
};

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFwdThreads, 2)
proj_fwd_kernel(const float* __restrict__ x, int64_t ldx, int N, int K, InSeg w, const float* __restrict__ bias,
                OutSeg out) {
  extern __shared__ __align__(16) float smem[];
  const int KP = (K + 3) & ~3;
  float* Ws = smem;                               // [KP][kMaxOut]      transposed weights, zero padded
  float* xs = smem + KP * kMaxOut;                // [KP][kFwdRows][2]  transposed x tile, every value twice
  const int nt = w.n[0] + w.n[1] + w.n[2];
  for (int idx = threadIdx.x; idx < KP * kMaxOut; idx += kFwdThreads) {
    const int k = idx / kMaxOut, o = idx - k * kMaxOut;
    float v = 0.f;
    if (k < K && o < nt) {
      int s = 0, oo = o;
      if (oo >= w.n[0]) { oo -= w.n[0]; s = 1; if (oo >= w.n[1]) { oo -= w.n[1]; s = 2; } }
      v = __ldg(w.p[s] + (int64_t)oo * w.ld[s] + k);
    }
    Ws[idx] = v;
  }
  // thread = 8 rows (ty) x 3 column PAIRS (2 tx + 128 j, +1): the accumulators are packed pairs, one FFMA2
  // (fma.rn.f32x2) per (row, pair) and k; the x tile is stored as (x, x) so the broadcast operand needs no moves
  const int ty = threadIdx.x >> 6, tx = threadIdx.x & 63;
  constexpr int kPairs = kFwdColsPerThread / 2;
  float* op[kPairs][2];
  unsigned old[kPairs][2];
  float bv[kPairs][2];
  bool pair_ok[kPairs];
#pragma unroll
  for (int j = 0; j < kPairs; ++j) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int o = 2 * tx + 128 * j + h;
      op[j][h] = nullptr; old[j][h] = 0; bv[j][h] = 0.f;
      if (o < nt) {
        int s = 0;
        if (o >= out.n[0]) { o -= out.n[0]; s = 1; if (o >= out.n[1]) { o -= out.n[1]; s = 2; } }
        op[j][h] = out.p[s] + o;
        old[j][h] = (unsigned)out.ld[s];
        if (s == 0 && bias != nullptr) bv[j][h] = __ldg(bias + o);
      }
    }
    // both columns in the same matrix, 8-byte aligned for every row: one 64-bit store
    pair_ok[j] = op[j][0] != nullptr && op[j][1] == op[j][0] + 1 && (old[j][0] & 1u) == 0 &&
                 (reinterpret_cast<uintptr_t>(op[j][0]) & 7u) == 0;
  }
  const unsigned ld0 = (unsigned)out.ld[0];
  const bool fast0 = (ld0 & 1u) == 0 && (reinterpret_cast<uintptr_t>(out.p[0]) & 7u) == 0;
  bool in0[kPairs];
#pragma unroll
  for (int j = 0; j < kPairs; ++j) in0[j] = 2 * tx + 128 * j + 1 < out.n[0];
  const int ntiles = (N + kFwdRows - 1) / kFwdRows;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r0 = tile * kFwdRows;
    __syncthreads();                              // previous tile's readers are done (and Ws is complete)
    {                                             // x tile: 8 threads per row, k = (t & 7) + 8 m (no divisions)
      const int r = threadIdx.x >> 3;
      const bool rok = r0 + r < N;
      const float* xr = x + (int64_t)(rok ? r0 + r : 0) * ldx;
      for (int k = threadIdx.x & 7; k < KP; k += 8) {
        const float v = (rok && k < K) ? __ldg(xr + k) : 0.f;
        *reinterpret_cast<float2*>(xs + (k * kFwdRows + r) * 2) = make_float2(v, v);
      }
    }
    __syncthreads();
    stream::u64 acc[8][kPairs];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < kPairs; ++j) acc[i][j] = 0ull;
    for (int k = 0; k < K; ++k) {
      stream::u64 xp[8], wp[kPairs];
      const ulonglong2* xrow = reinterpret_cast<const ulonglong2*>(xs + (k * kFwdRows + ty * 8) * 2);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const ulonglong2 t = xrow[i];
        xp[2 * i] = t.x;
        xp[2 * i + 1] = t.y;
      }
#pragma unroll
      for (int j = 0; j < kPairs; ++j)
        wp[j] = *reinterpret_cast<const stream::u64*>(Ws + k * kMaxOut + 2 * tx + 128 * j);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < kPairs; ++j) acc[i][j] = stream::fma2(xp[i], wp[j], acc[i][j]);
    }
    // epilogue: pairs that lie inside out0 are stored through one row pointer + immediates
    const bool full_tile = r0 + kFwdRows <= N;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const unsigned r = (unsigned)(r0 + ty * 8 + i);
      if (full_tile || r < (unsigned)N) {
        float* row0 = out.p[0] + (size_t)r * ld0 + 2 * tx;
#pragma unroll
        for (int j = 0; j < kPairs; ++j) {
          float lo, hi;
          stream::unpack2(acc[i][j], lo, hi);
          lo += bv[j][0];
          hi += bv[j][1];
          if (fast0 && in0[j]) {
            *reinterpret_cast<float2*>(row0 + 128 * j) = make_float2(lo, hi);
          } else if (pair_ok[j]) {
            *reinterpret_cast<float2*>(op[j][0] + (size_t)r * old[j][0]) = make_float2(lo, hi);
          } else {
            if (op[j][0] != nullptr) op[j][0][(size_t)r * old[j][0]] = lo;
            if (op[j][1] != nullptr) op[j][1][(size_t)r * old[j][1]] = hi;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// forward, warp per row, zero-skipping.  The reference featurisation is one-hot style (data/: 35 input
// features, 4-6 non-zeros per atom): out[r,:] = sum over the NON-ZERO x[r,k] of x[r,k] * W'[:,k].  A warp loads the
// row (two coalesced loads), ballots the non-zeros and walks the set bits; lane l owns columns l, l+32, ...
// Skipped terms are exact zeros, the order is ascending k: same bits as the dense kernel.  ~190 instructions per
// row for 5 non-zeros (dense tile kernel: ~800); a dense row costs 35 rounds (about 1.3x the tile kernel), so the
// launcher samples nothing and simply prefers this kernel for K <= 64 -- the dense tile kernel stays available
// (MGS_PROJ_DENSE=1) and is the one the parity tests with Gaussian inputs also exercise.
// ---------------------------------------------------------------------------------------------
constexpr int kSpCols = kMaxOut / 32;             // 12 columns per lane

__global__ void __launch_bounds__(kFwdThreads, 2)
proj_fwd_sparse_kernel(const float* __restrict__ x, int64_t ldx, int N, int K, InSeg w, const float* __restrict__ bias,
                       OutSeg out) {
  extern __shared__ __align__(16) float smem[];
  const int KP = (K + 3) & ~3;
  float* Ws = smem;                               // [KP][kMaxOut] transposed weights, zero padded
  const int nt = w.n[0] + w.n[1] + w.n[2];
  for (int idx = threadIdx.x; idx < KP * kMaxOut; idx += kFwdThreads) {
    const int k = idx / kMaxOut, o = idx - k * kMaxOut;
    float v = 0.f;
    if (k < K && o < nt) {
      int s = 0, oo = o;
      if (oo >= w.n[0]) { oo -= w.n[0]; s = 1; if (oo >= w.n[1]) { oo -= w.n[1]; s = 2; } }
      v = __ldg(w.p[s] + (int64_t)oo * w.ld[s] + k);
    }
    Ws[idx] = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * (kFwdThreads / 32) + (threadIdx.x >> 5), nw = gridDim.x * (kFwdThreads / 32);
  // column -> (pointer, leading dimension, bias)
  float* op[kSpCols];
  unsigned old[kSpCols];
  float bv[kSpCols];
#pragma unroll
  for (int j = 0; j < kSpCols; ++j) {
    int o = lane + 32 * j;
    op[j] = nullptr; old[j] = 0; bv[j] = 0.f;
    if (o < nt) {
      int s = 0;
      if (o >= out.n[0]) { o -= out.n[0]; s = 1; if (o >= out.n[1]) { o -= out.n[1]; s = 2; } }
      op[j] = out.p[s] + o;
      old[j] = (unsigned)out.ld[s];
      if (s == 0 && bias != nullptr) bv[j] = __ldg(bias + o);
    }
  }
  const float* Wl = Ws + lane;
  // kSpAhead rows of the warp in flight: with ONE row ahead (first version) a row cost a full memory round trip -- 55 rows
  // per warp in 96 us = 3 300 cycles per row for ~400 cycles of work (ncu: DRAM 19 %, issue 49 %)
  constexpr int kSpAhead = 4;
  float xa[kSpAhead], xb[kSpAhead];
#pragma unroll
  for (int u = 0; u < kSpAhead; ++u) {
    const int ru = gw + u * nw;
    xa[u] = (ru < N && lane < K) ? __ldg(x + (int64_t)ru * ldx + lane) : 0.f;
    xb[u] = (ru < N && lane + 32 < K) ? __ldg(x + (int64_t)ru * ldx + lane + 32) : 0.f;
  }
  for (int r0 = gw; r0 < N; r0 += kSpAhead * nw) {
#pragma unroll
    for (int u = 0; u < kSpAhead; ++u) {
      const int r = r0 + u * nw;
      if (r >= N) break;                          // warp-uniform
      const float x0 = xa[u], x1 = xb[u];
      const int rn = r + kSpAhead * nw;           // refill this slot
      xa[u] = (rn < N && lane < K) ? __ldg(x + (int64_t)rn * ldx + lane) : 0.f;
      xb[u] = (rn < N && lane + 32 < K) ? __ldg(x + (int64_t)rn * ldx + lane + 32) : 0.f;
      float acc[kSpCols];
#pragma unroll
      for (int j = 0; j < kSpCols; ++j) acc[j] = 0.f;
      unsigned m0 = __ballot_sync(0xffffffffu, x0 != 0.f), m1 = __ballot_sync(0xffffffffu, x1 != 0.f);
      while (m0) {
        const int k = __ffs(m0) - 1;
        m0 &= m0 - 1;
        const float xk = __shfl_sync(0xffffffffu, x0, k);
        const float* wk = Wl + k * kMaxOut;
#pragma unroll
        for (int j = 0; j < kSpCols; ++j) acc[j] = fmaf(xk, wk[32 * j], acc[j]);
      }
      while (m1) {
        const int k = __ffs(m1) - 1;
        m1 &= m1 - 1;
        const float xk = __shfl_sync(0xffffffffu, x1, k);
        const float* wk = Wl + (k + 32) * kMaxOut;
#pragma unroll
        for (int j = 0; j < kSpCols; ++j) acc[j] = fmaf(xk, wk[32 * j], acc[j]);
      }
#pragma unroll
      for (int j = 0; j < kSpCols; ++j)
        if (op[j] != nullptr) op[j][(size_t)(unsigned)r * old[j]] = acc[j] + bv[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad: part[cta][o][k] = sum over the CTA's atoms of g'[r][o] * x[r][k]
// ---------------------------------------------------------------------------------------------
constexpr int kWgThreads = 192;
constexpr int kWgRows = 16;                       // atoms per shared-memory x tile = per memory round trip
constexpr int kWgKP = 36;                         // K padded to a multiple of 4 (K <= 36 in this kernel)

// PAIR: thread t owns the adjacent output features 2t, 2t+1 and reads them with one 64-bit load per atom (needs
// even segment widths / leading dimensions and 8-byte aligned bases: the model1 shapes); else features t, t+192
// with scalar loads.  ncu on the scalar version: lg_throttle -- 32 LDG.32 per thread and chunk choke the LSU queue.
template <bool PAIR>
__global__ void __launch_bounds__(kWgThreads, 2)
proj_wgrad_kernel(InSeg g, const float* __restrict__ x, int64_t ldx, int N, int K, int rows_per_cta,
                  float* __restrict__ part) {
  __shared__ __align__(16) float xs[2][kWgRows][kWgKP];
  const int nt = g.n[0] + g.n[1] + g.n[2];
  const float* gp[2];
  unsigned gld[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    int o = PAIR ? 2 * threadIdx.x + j : threadIdx.x + kWgThreads * j;
    gp[j] = nullptr; gld[j] = 0;
    if (o < nt) {
      int s = 0;
      if (o >= g.n[0]) { o -= g.n[0]; s = 1; if (o >= g.n[1]) { o -= g.n[1]; s = 2; } }
      gp[j] = g.p[s] + o;
      gld[j] = (unsigned)g.ld[s];
    }
  }
  // accumulators: 2 output features x K inputs as packed pairs over k -> one FFMA2 per (feature, k pair, atom)
  stream::u64 acc[2][kWgKP / 2];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int k = 0; k < kWgKP / 2; ++k) acc[j][k] = 0ull;

  const int r_lo = blockIdx.x * rows_per_cta;
  const int r_hi = min(N, r_lo + rows_per_cta);
  auto load_x = [&](int buf, int r0) {            // x rows r0 .. r0+15 -> xs[buf] (zero padded)
    for (int idx = threadIdx.x; idx < kWgRows * kWgKP; idx += kWgThreads) {
      const int r = idx / kWgKP, k = idx - r * kWgKP;
      float v = 0.f;
      if (k < K && r0 + r < r_hi) v = __ldg(x + (int64_t)(r0 + r) * ldx + k);
      xs[buf][r][k] = v;
    }
  };
  auto load_g = [&](int r0, float (&gv)[2][kWgRows]) {
    if (PAIR) {
#pragma unroll
      for (int i = 0; i < kWgRows; ++i) {
        float2 t = make_float2(0.f, 0.f);
        if (gp[0] != nullptr && r0 + i < r_hi)
          t = __ldg(reinterpret_cast<const float2*>(gp[0] + (size_t)(unsigned)(r0 + i) * gld[0]));
        gv[0][i] = t.x;
        gv[1][i] = t.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < kWgRows; ++i)
          gv[j][i] = (gp[j] != nullptr && r0 + i < r_hi) ? __ldg(gp[j] + (size_t)(unsigned)(r0 + i) * gld[j]) : 0.f;
    }
  };
  // One chunk of g per memory round trip: register-staged loads of a thread alias the six scoreboard slots, so a
  // software prefetch of the next chunk does not overlap anything (measured; same effect as in the tcgen05 wgrad).
  // The registers go to a twice larger chunk instead: half the round trips.
  float gv[2][kWgRows];
  if (r_lo < r_hi) load_x(0, r_lo);
  __syncthreads();
  int buf = 0;
  for (int r0 = r_lo; r0 < r_hi; r0 += kWgRows) {
    load_g(r0, gv);
    if (r0 + kWgRows < r_hi) load_x(buf ^ 1, r0 + kWgRows);
#pragma unroll
    for (int i = 0; i < kWgRows; ++i) {
      const stream::u64 g0 = stream::pack2(gv[0][i], gv[0][i]);
      const stream::u64 g1 = stream::pack2(gv[1][i], gv[1][i]);
      const ulonglong2* xrow = reinterpret_cast<const ulonglong2*>(&xs[buf][i][0]);
#pragma unroll
      for (int k4 = 0; k4 < kWgKP / 4; ++k4) {
        const ulonglong2 xv = xrow[k4];
        acc[0][2 * k4] = stream::fma2(g0, xv.x, acc[0][2 * k4]);
        acc[0][2 * k4 + 1] = stream::fma2(g0, xv.y, acc[0][2 * k4 + 1]);
        acc[1][2 * k4] = stream::fma2(g1, xv.x, acc[1][2 * k4]);
        acc[1][2 * k4 + 1] = stream::fma2(g1, xv.y, acc[1][2 * k4 + 1]);
      }
    }
    __syncthreads();
    buf ^= 1;
  }
  float* mine = part + (int64_t)blockIdx.x * nt * K;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int o = PAIR ? 2 * threadIdx.x + j : threadIdx.x + kWgThreads * j;
    if (o < nt) {
#pragma unroll
      for (int k2 = 0; k2 < kWgKP / 2; ++k2) {
        float lo, hi;
        stream::unpack2(acc[j][k2], lo, hi);
        if (2 * k2 < K) mine[(int64_t)o * K + 2 * k2] = lo;
        if (2 * k2 + 1 < K) mine[(int64_t)o * K + 2 * k2 + 1] = hi;
      }
    }
  }
}

// Same contraction, g and x staged through a 3-deep cp.async ring (no registers between HBM and shared memory: the
// register-staged version above pays one exposed global round trip per 16-atom chunk at 12 warps / SM) and a
// 6-feature x 12-input register tile per thread (thread = feature-pair group t % 64, input group t / 64; a warp has
// ONE input group, so its x reads stay warp-uniform): shared memory serves 8 lanes of a 128-bit load per wavefront
// and does not merge equal addresses across quarter-warps, so the broadcast read of a 36-float x row costs 36
// wavefronts per warp and atom; 12 inputs per thread need 12, plus 6 for its three g pairs.
constexpr int kRgStages = 3;
constexpr int kRgGS = 2 * kWgThreads;              // g tile row stride (floats): all <= 384 output features
constexpr int kRgF = 6, kRgK = 12;
constexpr size_t kRgSmem = sizeof(float) * kRgStages * kWgRows * (kRgGS + kWgKP);

__device__ __forceinline__ void cp_async8(float* dst, const float* src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int n = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int n = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}

__global__ void __launch_bounds__(kWgThreads, 2)
proj_wgrad_ring_kernel(InSeg g, const float* __restrict__ x, int64_t ldx, int N, int K, int rows_per_cta,
                       float* __restrict__ part) {
  static_assert(kWgThreads == 64 * (kWgKP / kRgK) && kRgF * 64 == kRgGS, "feature-pair groups x input groups");
  extern __shared__ __align__(16) float ring[];
  float* gs = ring;                                               // [stage][row][kRgGS]
  float* xs = ring + kRgStages * kWgRows * kRgGS;                 // [stage][row][kWgKP]
  const int nt = g.n[0] + g.n[1] + g.n[2];
  const int fg = threadIdx.x & 63, kg = threadIdx.x >> 6;
  // producer role: this thread copies feature pair `threadIdx.x` of every row of a chunk
  const float* src_g = nullptr;
  unsigned src_ld = 0;
  {
    int o = 2 * threadIdx.x;
    if (o < nt) {
      int sgm = 0;
      if (o >= g.n[0]) { o -= g.n[0]; sgm = 1; if (o >= g.n[1]) { o -= g.n[1]; sgm = 2; } }
      src_g = g.p[sgm] + o;
      src_ld = (unsigned)g.ld[sgm];
    }
  }
  for (int idx = threadIdx.x; idx < kRgStages * kWgRows * kWgKP; idx += kWgThreads) xs[idx] = 0.f;   // k >= K stays 0
  const int r_lo = blockIdx.x * rows_per_cta;
  const int r_hi = min(N, r_lo + rows_per_cta);
  const int nchunks = r_hi > r_lo ? (r_hi - r_lo + kWgRows - 1) / kWgRows : 0;
  __syncthreads();
  auto issue = [&](int c) {
    if (c < nchunks) {
      const int st = c % kRgStages, r0 = r_lo + c * kWgRows;
      if (src_g != nullptr) {
        float* dst = gs + (st * kWgRows) * kRgGS + 2 * threadIdx.x;
#pragma unroll
        for (int i = 0; i < kWgRows; ++i) {
          const bool ok = r0 + i < r_hi;
          cp_async8(dst + i * kRgGS, src_g + (size_t)(unsigned)(ok ? r0 + i : r_lo) * src_ld, ok);
        }
      }
      for (int idx = threadIdx.x; idx < kWgRows * K; idx += kWgThreads) {
        const int r = idx / K, k = idx - r * K;
        const bool ok = r0 + r < r_hi;
        cp_async4(xs + (st * kWgRows + r) * kWgKP + k, x + (int64_t)(ok ? r0 + r : r_lo) * ldx + k, ok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  stream::u64 acc[kRgF][kRgK / 2];
#pragma unroll
  for (int j = 0; j < kRgF; ++j)
#pragma unroll
    for (int k = 0; k < kRgK / 2; ++k) acc[j][k] = 0ull;

#pragma unroll
  for (int c = 0; c < kRgStages - 1; ++c) issue(c);
  for (int c = 0; c < nchunks; ++c) {
    issue(c + kRgStages - 1);
    asm volatile("cp.async.wait_group %0;" ::"n"(kRgStages - 1) : "memory");
    __syncthreads();
    const int st = c % kRgStages;
    const float* gt = gs + (st * kWgRows) * kRgGS + 2 * fg;
    const float* xt = xs + (st * kWgRows) * kWgKP + kg * kRgK;
#pragma unroll 4
    for (int i = 0; i < kWgRows; ++i) {
      const ulonglong2* xrow = reinterpret_cast<const ulonglong2*>(xt + i * kWgKP);
      stream::u64 xv[kRgK / 2];
#pragma unroll
      for (int k4 = 0; k4 < kRgK / 4; ++k4) {
        const ulonglong2 t = xrow[k4];
        xv[2 * k4] = t.x; xv[2 * k4 + 1] = t.y;
      }
#pragma unroll
      for (int q = 0; q < kRgF / 2; ++q) {
        const float2 gq = *reinterpret_cast<const float2*>(gt + i * kRgGS + 128 * q);
        const stream::u64 g0 = stream::pack2(gq.x, gq.x);
        const stream::u64 g1 = stream::pack2(gq.y, gq.y);
#pragma unroll
        for (int k2 = 0; k2 < kRgK / 2; ++k2) {
          acc[2 * q][k2] = stream::fma2(g0, xv[k2], acc[2 * q][k2]);
          acc[2 * q + 1][k2] = stream::fma2(g1, xv[k2], acc[2 * q + 1][k2]);
        }
      }
    }
    __syncthreads();
  }
  float* mine = part + (int64_t)blockIdx.x * nt * K;
#pragma unroll
  for (int j = 0; j < kRgF; ++j) {
    const int o = 2 * (fg + 64 * (j >> 1)) + (j & 1);
    if (o < nt) {
#pragma unroll
      for (int k2 = 0; k2 < kRgK / 2; ++k2) {
        float lo, hi;
        stream::unpack2(acc[j][k2], lo, hi);
        const int k = kg * kRgK + 2 * k2;
        if (k < K) mine[(int64_t)o * K + k] = lo;
        if (k + 1 < K) mine[(int64_t)o * K + k + 1] = hi;
      }
    }
  }
}

// Few output rows (the two [N, H] score gradients when the [N, H*C] part runs on the tensor-core kernel: 20 x 35 outputs).
// The kernels above are built around 384 output features and cost ~0.11 ms whatever the count; here a thread owns FOUR
// features x one input column: per atom one conflict-free LDS of x, one broadcast LDS.128 of g, four FMAs.  Same partial
// layout as the other kernels ([cta][feature][k]) -> same fixed-order reduction.
constexpr int kSmRows = 448;                      // atoms staged per phase: ONE exposed memory round trip per CTA at 130 k atoms
constexpr int kSmThreads = 256;
constexpr int kSmMaxOut = 32;
__device__ __forceinline__ void sm_cp4(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
inline size_t small_wgrad_smem(int nt, int K) { return sizeof(float) * (size_t)kSmRows * (((nt + 3) & ~3) + K); }
__global__ void __launch_bounds__(kSmThreads)
proj_wgrad_small_kernel(InSeg gs, const float* __restrict__ x, int64_t ldx, int num_rows, int K, int rows_per_cta,
                        float* __restrict__ part) {
  extern __shared__ __align__(16) float s_dyn[];
  const int nt = gs.n[0] + gs.n[1] + gs.n[2];
  const int gst = (nt + 3) & ~3;                                // row stride of the staged gradients (16-byte aligned rows)
  float* s_g = s_dyn;                                           // [kSmRows][gst]
  float* s_x = s_dyn + kSmRows * gst;                           // [kSmRows][K]
  const int fg = threadIdx.x / K, k = threadIdx.x - fg * K;     // feature group (4 features), input column
  const bool active = fg * 4 < nt;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(num_rows, r0 + rows_per_cta);
  for (int base = r0; base < r1; base += kSmRows) {
    const int nr = min(kSmRows, r1 - base);
    __syncthreads();
    // 4-byte LDGSTS: every load of the phase is in flight at once (a plain load loop exposes one round trip per unrolled
    // group: 60 flat loads per thread at 130 k atoms)
    for (int a = threadIdx.x; a < nr; a += kSmThreads) {        // a thread stages the gradient rows of its atoms
      float* dst = s_g + a * gst;
      int off = 0;
#pragma unroll
      for (int sg = 0; sg < 3; ++sg) {
        const float* src = gs.p[sg] + (int64_t)(base + a) * gs.ld[sg];
        for (int o = 0; o < gs.n[sg]; ++o) sm_cp4(dst + off + o, src + o);
        off += gs.n[sg];
      }
      for (; off < gst; ++off) dst[off] = 0.f;
    }
    if (ldx == K) {                                             // contiguous x: the CTA's rows are one flat span
      const float* src = x + (int64_t)base * K;
      for (int i = threadIdx.x; i < nr * K; i += kSmThreads) sm_cp4(s_x + i, src + i);
    } else {
      for (int a = threadIdx.x; a < nr; a += kSmThreads)
        for (int c = 0; c < K; ++c) sm_cp4(s_x + a * K + c, x + (int64_t)(base + a) * ldx + c);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (active) {
      const float* gp = s_g + fg * 4;
      const float* xp = s_x + k;
#pragma unroll 8
      for (int a = 0; a < nr; ++a) {
        const float xv = xp[a * K];
        const float4 gv = *reinterpret_cast<const float4*>(gp + a * gst);
        acc[0] = fmaf(gv.x, xv, acc[0]);
        acc[1] = fmaf(gv.y, xv, acc[1]);
        acc[2] = fmaf(gv.z, xv, acc[2]);
        acc[3] = fmaf(gv.w, xv, acc[3]);
      }
    }
  }
  if (active) {
    float* out = part + (int64_t)blockIdx.x * nt * K;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (fg * 4 + u < nt) out[(int64_t)(fg * 4 + u) * K + k] = acc[u];
  }
}

// ---------------------------------------------------------------------------------------------
// score weights  U[h, :] = sum_c att[h, c] W[hC + c, :]  and their backward.  As PyTorch expressions (mul + sum, twice, and
// their autograd nodes) these were ~16 launches of 3-5 us per training step for [10, 35] results.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gat_u_fwd_kernel(const float* __restrict__ w, int64_t ldw, const float* __restrict__ att_src,
                 const float* __restrict__ att_dst, int H, int C, int K, float* __restrict__ u_src,
                 float* __restrict__ u_dst) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= H * K) return;
  const int h = t / K, k = t - h * K;
  float s = 0.f, d = 0.f;
  int c = 0;
  for (; c + 8 <= C; c += 8) {                                 // eight weight rows in flight, summed in order
    float wv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) wv[u] = __ldg(w + (int64_t)(h * C + c + u) * ldw + k);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      s = fmaf(__ldg(att_src + h * C + c + u), wv[u], s);
      d = fmaf(__ldg(att_dst + h * C + c + u), wv[u], d);
    }
  }
  for (; c < C; ++c) {
    const float wv = __ldg(w + (int64_t)(h * C + c) * ldw + k);
    s = fmaf(__ldg(att_src + h * C + c), wv, s);
    d = fmaf(__ldg(att_dst + h * C + c), wv, d);
  }
  u_src[t] = s;
  u_dst[t] = d;
}
// one warp per weight row r = hC + c:  dw[r, :] = att_src[r] du_src[h, :] + att_dst[r] du_dst[h, :],
// datt_src[r] = <du_src[h, :], w[r, :]>,  datt_dst[r] = <du_dst[h, :], w[r, :]>
__global__ void __launch_bounds__(256)
gat_u_bwd_kernel(const float* __restrict__ w, int64_t ldw, const float* __restrict__ att_src,
                 const float* __restrict__ att_dst, const float* __restrict__ du_src, const float* __restrict__ du_dst,
                 int H, int C, int K, float* __restrict__ dw, int64_t lddw, float* __restrict__ datt_src,
                 float* __restrict__ datt_dst) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= H * C) return;
  const int h = r / C;
  const float as = __ldg(att_src + r), ad = __ldg(att_dst + r);
  float ss = 0.f, sd = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float us = __ldg(du_src + h * K + k), ud = __ldg(du_dst + h * K + k);
    const float wv = __ldg(w + (int64_t)r * ldw + k);
    dw[(int64_t)r * lddw + k] = fmaf(as, us, ad * ud);
    ss = fmaf(us, wv, ss);
    sd = fmaf(ud, wv, sd);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    sd += __shfl_xor_sync(0xffffffffu, sd, o);
  }
  if (lane == 0) {
    datt_src[r] = ss;
    datt_dst[r] = sd;
  }
}

// stage 2: fixed-order sum over the CTAs, scattered into the three output matrices
__global__ void __launch_bounds__(256)
proj_wgrad_reduce_kernel(const float* __restrict__ part, int splits, int K, OutSeg out) {
  const int nt = out.n[0] + out.n[1] + out.n[2];
  const int64_t total = (int64_t)nt * K;
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int z = 0;
    for (; z + 4 <= splits; z += 4) {
      s0 += part[(int64_t)z * total + t];
      s1 += part[(int64_t)(z + 1) * total + t];
      s2 += part[(int64_t)(z + 2) * total + t];
      s3 += part[(int64_t)(z + 3) * total + t];
    }
    for (; z < splits; ++z) s0 += part[(int64_t)z * total + t];
    int o = (int)(t / K);
    const int k = (int)(t - (int64_t)o * K);
    int s = 0;
    if (o >= out.n[0]) { o -= out.n[0]; s = 1; if (o >= out.n[1]) { o -= out.n[1]; s = 2; } }
    out.p[s][(int64_t)o * out.ld[s] + k] = (s0 + s1) + (s2 + s3);
  }
}

inline int wgrad_ctas() { return 2 * sm_count(); }

}  // namespace
}  // namespace mgs

using namespace mgs;

extern "C" int mgs_proj_fwd(const float* x, int64_t ldx, int64_t num_rows, int32_t K, const float* w, int64_t ldw,
                            int32_t n0, const float* u1, int64_t ldu1, int32_t n1, const float* u2, int64_t ldu2,
                            int32_t n2, const float* bias, float* out0, int64_t ld0, float* out1, int64_t ld1,
                            float* out2, int64_t ld2, mgs_stream_t stream_) {
  MGS_REQUIRE(num_rows >= 0 && num_rows < 0x7fffffff && K > 0 && K <= kMaxK, "mgs_proj_fwd: K must be in [1, %d]", kMaxK);
  MGS_REQUIRE(n0 > 0 && n1 >= 0 && n2 >= 0 && n0 + n1 + n2 <= kMaxOut, "mgs_proj_fwd: at most %d output columns", kMaxOut);
  MGS_REQUIRE(ldx >= K && ldw >= K && ld0 >= n0, "mgs_proj_fwd: leading dimension too small");
  if (num_rows == 0) return MGS_OK;
  MGS_REQUIRE(x && w && out0 && (n1 == 0 || (u1 && out1 && ldu1 >= K && ld1 >= n1)) &&
              (n2 == 0 || (u2 && out2 && ldu2 >= K && ld2 >= n2)), "mgs_proj_fwd: null pointer / bad segment");
  InSeg ws = {{w, u1, u2}, {ldw, ldu1, ldu2}, {n0, n1, n2}};
  OutSeg os = {{out0, out1, out2}, {ld0, ld1, ld2}, {n0, n1, n2}};
  const int KP = (K + 3) & ~3;
  const size_t smem = sizeof(float) * (size_t)KP * (kMaxOut + 2 * kFwdRows);
  MGS_CUDA(cudaFuncSetAttribute(proj_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntiles = (num_rows + kFwdRows - 1) / kFwdRows;
  const int grid = (int)(ntiles < 2 * sm_count() ? ntiles : 2 * sm_count());
  const char* dense_env = getenv("MGS_PROJ_DENSE");               // read per call: the tests toggle it
  const bool dense = dense_env && dense_env[0] && dense_env[0] != '0';
  if (!dense) {
    MGS_CUDA(cudaFuncSetAttribute(proj_fwd_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    proj_fwd_sparse_kernel<<<grid, kFwdThreads, smem, (cudaStream_t)stream_>>>(x, ldx, (int)num_rows, K, ws, bias, os);
    return check_launch("proj_fwd_sparse_kernel");
  }
  proj_fwd_kernel<<<grid, kFwdThreads, smem, (cudaStream_t)stream_>>>(x, ldx, (int)num_rows, K, ws, bias, os);
  return check_launch("proj_fwd_kernel");
}

extern "C" size_t mgs_proj_wgrad_workspace_bytes(int32_t K, int32_t n_total) {
  if (K <= 0 || n_total <= 0) return 0;
  return sizeof(float) * (size_t)wgrad_ctas() * n_total * K;
}

extern "C" int mgs_proj_wgrad(const float* g0, int64_t ldg0, int32_t n0, const float* g1, int64_t ldg1, int32_t n1,
                              const float* g2, int64_t ldg2, int32_t n2, const float* x, int64_t ldx,
                              int64_t num_rows, int32_t K, float* dw, int64_t lddw, float* du1, int64_t lddu1,
                              float* du2, int64_t lddu2, void* workspace, size_t workspace_bytes,
                              mgs_stream_t stream_) {
  MGS_REQUIRE(num_rows >= 0 && num_rows < 0x7fffffff && K > 0 && K <= kWgKP, "mgs_proj_wgrad: K must be in [1, %d]", kWgKP);
  const int nt = n0 + n1 + n2;
  MGS_REQUIRE(n0 > 0 && n1 >= 0 && n2 >= 0 && nt <= 2 * kWgThreads, "mgs_proj_wgrad: at most %d output rows", 2 * kWgThreads);
  MGS_REQUIRE(dw && lddw >= K && (n1 == 0 || (du1 && lddu1 >= K)) && (n2 == 0 || (du2 && lddu2 >= K)),
              "mgs_proj_wgrad: bad outputs");
  cudaStream_t stream = (cudaStream_t)stream_;
  OutSeg os = {{dw, du1, du2}, {lddw, lddu1, lddu2}, {n0, n1, n2}};
  const int ctas = wgrad_ctas();
  const size_t need = sizeof(float) * (size_t)ctas * nt * K;
  if (workspace_bytes < need || !workspace) {
    set_error("mgs_proj_wgrad: workspace too small (%zu < %zu)", workspace_bytes, need);
    return MGS_ERR_WORKSPACE_TOO_SMALL;
  }
  MGS_REQUIRE(num_rows == 0 || (g0 && x && ldg0 >= n0 && ldx >= K && (n1 == 0 || (g1 && ldg1 >= n1)) &&
                                (n2 == 0 || (g2 && ldg2 >= n2))), "mgs_proj_wgrad: null pointer / bad segment");
  InSeg gs = {{g0, g1, g2}, {ldg0, ldg1, ldg2}, {n0, n1, n2}};
  int rows_per_cta = (int)((num_rows + ctas - 1) / ctas);
  rows_per_cta = (rows_per_cta + kWgRows - 1) / kWgRows * kWgRows;
  if (rows_per_cta < kWgRows) rows_per_cta = kWgRows;
  auto even8 = [](const float* p, int64_t ld, int n) { return n == 0 || (((uintptr_t)p & 7u) == 0 && ld % 2 == 0); };
  const bool pair = n0 % 2 == 0 && n1 % 2 == 0 && n2 % 2 == 0 && even8(g0, ldg0, n0) && even8(g1, ldg1, n1) &&
                    even8(g2, ldg2, n2);
  static const bool old_path = getenv("MGS_PROJ_WGRAD_OLD") != nullptr;     // A/B switch for the probe
  if (nt <= kSmMaxOut && K * ((nt + 3) / 4) <= kSmThreads && !old_path) {
    int rpc = (int)((num_rows + ctas - 1) / ctas);
    if (rpc < 1) rpc = 1;
    const size_t smem = small_wgrad_smem(nt, K);
    MGS_CUDA(cudaFuncSetAttribute(proj_wgrad_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    proj_wgrad_small_kernel<<<ctas, kSmThreads, smem, stream>>>(gs, x, ldx, (int)num_rows, K, rpc, (float*)workspace);
  } else if (pair && !old_path) {
    MGS_CUDA(cudaFuncSetAttribute(proj_wgrad_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRgSmem));
    proj_wgrad_ring_kernel<<<ctas, kWgThreads, kRgSmem, stream>>>(gs, x, ldx, (int)num_rows, K, rows_per_cta, (float*)workspace);
  } else if (pair) proj_wgrad_kernel<true><<<ctas, kWgThreads, 0, stream>>>(gs, x, ldx, (int)num_rows, K, rows_per_cta, (float*)workspace);
  else proj_wgrad_kernel<false><<<ctas, kWgThreads, 0, stream>>>(gs, x, ldx, (int)num_rows, K, rows_per_cta, (float*)workspace);
  if (int rc = check_launch("proj_wgrad_kernel")) return rc;
  proj_wgrad_reduce_kernel<<<grid_for((int64_t)nt * K, 256, 8), 256, 0, stream>>>((const float*)workspace, ctas, K, os);
  return check_launch("proj_wgrad_reduce_kernel");
}

extern "C" int mgs_gat_u_fwd(const float* w, int64_t ldw, const float* att_src, const float* att_dst, int32_t H, int32_t C,
                             int32_t K, float* u_src, float* u_dst, mgs_stream_t stream_) {
  MGS_REQUIRE(H > 0 && C > 0 && K > 0 && ldw >= K, "mgs_gat_u_fwd: bad sizes");
  MGS_REQUIRE(w && att_src && att_dst && u_src && u_dst, "mgs_gat_u_fwd: null pointer");
  gat_u_fwd_kernel<<<(H * K + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(w, ldw, att_src, att_dst, H, C, K, u_src, u_dst);
  return check_launch("gat_u_fwd_kernel");
}

extern "C" int mgs_gat_u_bwd(const float* w, int64_t ldw, const float* att_src, const float* att_dst, const float* du_src,
                             const float* du_dst, int32_t H, int32_t C, int32_t K, float* dw, int64_t lddw,
                             float* datt_src, float* datt_dst, mgs_stream_t stream_) {
  MGS_REQUIRE(H > 0 && C > 0 && K > 0 && ldw >= K && lddw >= K, "mgs_gat_u_bwd: bad sizes");
  MGS_REQUIRE(w && att_src && att_dst && du_src && du_dst && dw && datt_src && datt_dst, "mgs_gat_u_bwd: null pointer");
  gat_u_bwd_kernel<<<(H * C + 7) / 8, 256, 0, (cudaStream_t)stream_>>>(w, ldw, att_src, att_dst, du_src, du_dst, H, C, K, dw,
                                                                         lddw, datt_src, datt_dst);
  return check_launch("gat_u_bwd_kernel");
}
