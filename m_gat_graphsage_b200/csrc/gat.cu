// K2 -- GATConv message passing, forward and backward.
//
// Forward  = scores (a_src, a_dst) -> alpha (edge softmax per destination, thread per (atom, head))
//            -> aggregate (thread per (atom, feature chunk), coalesced 64/128-bit loads of xh rows).
// Backward = bwd_edge (warp per destination: g_i and the gathered xh_j rows are staged in shared memory
//            with coalesced loads, then every (slot, head) dot product is owned by one lane group;
//            softmax Jacobian + leaky_relu' in registers) -> bwd_node (per source, same access pattern
//            as the forward aggregate over the by-source structure) -> bwd_att (two-stage column sums).
// No [E', H, C] message tensor is ever materialised (the reference's PyG path writes and re-reads ~3.1x
// the size of xh); the only per-edge arrays are alpha / dr of shape [(E+N), H].
// All of this is HBM/L2-bound gather work: no tensor cores.  Algorithmic bytes / atom, (H,C)=(10,35):
//   aggregate fwd: 4HC (xh) + 4HC (out) + 4H*E'/N (alpha) + 4 + 4E/N  ~ 2.93 KB
//   bwd_edge:      4HC (g) + 4HC (xh) + 12H*E'/N (alpha, dr)          ~ 3.2 KB
//   bwd_node:      4HC (g) + 4HC (dxh) + 8H*E'/N                      ~ 3.05 KB
#include "common.cuh"
#include "stream.cuh"
#include "edge.cuh"
#include "edge_mma.cuh"
#include "edge_fma.cuh"

namespace mgs {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kSlotBatch = 4;  // xh rows staged per warp at a time in bwd_edge

__device__ __forceinline__ float leaky(float v, float slope) { return v > 0.f ? v : v * slope; }

// lanes-per-head split for segmented dot products: S = power of two, H * S <= 32 when possible,
// slices of at least 4 elements
inline int pick_split(int H, int C) {
  int S = 1;
  while (S * 2 * H <= 32 && (C + S * 2 - 1) / (S * 2) >= 4) S *= 2;
  return S;
}

// ---------------------------------------------------------------------------------------------
// scores: a_src[n,h] = <xh[n,h,:], att_src[h,:]>, a_dst likewise.  Warp per row.
// dynamic smem: att_src[HC] | att_dst[HC] | per-warp row buffer [kWarps][HC]
// ---------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(kThreads)
gat_scores_kernel(const float* __restrict__ xh, int64_t ld, int N, int H, int C, int S,
                  const float* __restrict__ att_src, const float* __restrict__ att_dst,
                  float* __restrict__ a_src, float* __restrict__ a_dst) {
  extern __shared__ __align__(16) float smem[];
  const int HC = H * C;
  const int HCp = (HC + 3) & ~3;
  float* s_as = smem;
  float* s_ad = smem + HCp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* xs = smem + 2 * HCp + warp * HCp;
  for (int f = threadIdx.x; f < HC; f += kThreads) {
    s_as[f] = att_src[f];
    s_ad[f] = att_dst[f];
  }
  __syncthreads();
  const int L = (C + S - 1) / S;  // slice length
  const int units = H * S;
  for (int n = blockIdx.x * kWarps + warp; n < N; n += gridDim.x * kWarps) {
    const float* rowp = xh + (int64_t)n * ld;
    for (int f = lane * V; f < HC; f += 32 * V) {
      Vec<V> v = Vec<V>::load(rowp + f);
#pragma unroll
      for (int u = 0; u < V; ++u) xs[f + u] = v.v[u];
    }
    __syncwarp();
    for (int ub = 0; ub < units; ub += 32) {
      const int u = ub + lane;
      float ps = 0.f, pd = 0.f;
      int h = 0, s = 0;
      if (u < units) {
        h = u / S;
        s = u - h * S;
        const int c0 = s * L;
        const int len = min(L, C - c0);
        if (len > 0) {
          int c = u % len;  // rotated start: spreads lanes over banks
          const int base = h * C + c0;
          for (int t = 0; t < len; ++t) {
            const float xv = xs[base + c];
            ps = fmaf(xv, s_as[base + c], ps);
            pd = fmaf(xv, s_ad[base + c], pd);
            if (++c == len) c = 0;
          }
        }
      }
      for (int o = S >> 1; o > 0; o >>= 1) {
        ps += __shfl_xor_sync(0xffffffffu, ps, o);
        pd += __shfl_xor_sync(0xffffffffu, pd, o);
      }
      if (u < units && s == 0) {
        a_src[(int64_t)n * H + h] = ps;
        a_dst[(int64_t)n * H + h] = pd;
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// alpha: thread per (destination i, head h); slots of i: rowptr[i]+i+k (k-th in-edge), self at rowptr[i+1]+i
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
gat_alpha_kernel(const float* __restrict__ a_src, const float* __restrict__ a_dst, int N, int H,
                 const int* __restrict__ rowptr, const int* __restrict__ col, float slope,
                 float* __restrict__ alpha) {
  const int64_t total = (int64_t)N * H;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (int64_t)gridDim.x * kThreads) {
    const int i = (int)(t / H);
    const int h = (int)(t - (int64_t)i * H);
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
    const float ad = __ldg(a_dst + t);
    const float e_self = leaky(__ldg(a_src + t) + ad, slope);
    if (end - beg <= 8) {
      // molecules (degree <= 6): neighbour ids, then scores, as two batches of independent loads; logits stay in registers,
      // alpha is written once.  Same operations in the same order as the general path below: same bits.
      const int deg = end - beg;
      int j[8];
      float e[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) j[k] = k < deg ? __ldg(col + beg + k) : i;
#pragma unroll
      for (int k = 0; k < 8; ++k) e[k] = j[k] != i ? leaky(__ldg(a_src + (int64_t)j[k] * H + h) + ad, slope) : -INFINITY;
      float m = e_self;
#pragma unroll
      for (int k = 0; k < 8; ++k) m = fmaxf(m, e[k]);
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (j[k] != i) {
          e[k] = expf(e[k] - m);
          sum = __fadd_rn(sum, e[k]);
        } else {
          e[k] = 0.f;
        }
      }
      const float p_self = expf(e_self - m);
      sum = __fadd_rn(sum, p_self);
      sum = __fadd_rn(sum, 1e-16f);
      float* arow = alpha + ((int64_t)beg + i) * H + h;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < deg) arow[(int64_t)k * H] = __fdiv_rn(e[k], sum);
      arow[(int64_t)deg * H] = __fdiv_rn(p_self, sum);
      continue;
    }
    float m = e_self;
    for (int p = beg; p < end; ++p) {
      const int j = __ldg(col + p);
      if (j == i) continue;
      m = fmaxf(m, leaky(__ldg(a_src + (int64_t)j * H + h) + ad, slope));
    }
    float* arow = alpha + ((int64_t)beg + i) * H + h;
    float sum = 0.f;
    for (int p = beg; p < end; ++p) {
      const int j = __ldg(col + p);
      float pe = 0.f;
      if (j != i) {
        pe = expf(leaky(__ldg(a_src + (int64_t)j * H + h) + ad, slope) - m);
        sum = __fadd_rn(sum, pe);
      }
      arow[(int64_t)(p - beg) * H] = pe;
    }
    const float p_self = expf(e_self - m);
    sum = __fadd_rn(sum, p_self);
    sum = __fadd_rn(sum, 1e-16f);
    for (int p = beg; p < end; ++p) {
      float* a = arow + (int64_t)(p - beg) * H;
      *a = __fdiv_rn(*a, sum);
    }
    arow[(int64_t)(end - beg) * H] = __fdiv_rn(p_self, sum);
  }
}

// ---------------------------------------------------------------------------------------------
// aggregate: thread per (destination i, V-wide feature chunk)
// ---------------------------------------------------------------------------------------------
template <int V, bool WEIGHTED>
__global__ void __launch_bounds__(kThreads)
gat_aggr_fwd_kernel(const float* __restrict__ xh, int64_t ld, int N, int H, int C, int chunks,
                    const float* __restrict__ alpha, const int* __restrict__ rowptr,
                    const int* __restrict__ col, const int* __restrict__ perm,
                    const float* __restrict__ ew, const float* __restrict__ bias,
                    float* __restrict__ out, int64_t ldo) {
  const int64_t total = (int64_t)N * chunks;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (int64_t)gridDim.x * kThreads) {
    const int i = (int)(t / chunks);
    const int f = (int)(t - (int64_t)i * chunks) * V;
    int hh[V];
#pragma unroll
    for (int u = 0; u < V; ++u) hh[u] = (f + u) / C;
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
    const float* arow = alpha + ((int64_t)beg + i) * H;
    Vec<V> acc = vzero<V>();
    for (int p = beg; p <= end; p += 4) {
      int j[4];
      Vec<V> v[4];
      float a[4][V];
      float w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int pk = p + k;
        j[k] = pk < end ? __ldg(col + pk) : (pk == end ? i : -1);
        if (pk < end && j[k] == i) j[k] = -1;  // pre-existing self loop: removed by GATConv
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (j[k] >= 0) {
          v[k] = Vec<V>::load(xh + (int64_t)j[k] * ld + f);
          const float* ak = arow + (int64_t)(p + k - beg) * H;
          w[k] = 1.f;
          if (WEIGHTED && p + k < end) w[k] = __ldg(ew + __ldg(perm + p + k));
#pragma unroll
          for (int u = 0; u < V; ++u) {
            a[k][u] = (u > 0 && hh[u] == hh[u - 1]) ? a[k][u - 1] : __ldg(ak + hh[u]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (j[k] >= 0) {
#pragma unroll
          for (int u = 0; u < V; ++u) {
            float m = __fmul_rn(a[k][u], v[k].v[u]);       // msg = alpha * x_j
            if (WEIGHTED) m = __fmul_rn(m, w[k]);          // msg * sigmoid(edge_mask)   (A.4)
            acc.v[u] = __fadd_rn(acc.v[u], m);
          }
        }
      }
    }
    if (bias != nullptr) {
#pragma unroll
      for (int u = 0; u < V; ++u) acc.v[u] = __fadd_rn(acc.v[u], __ldg(bias + f + u));
    }
    acc.store(out + (int64_t)i * ldo + f);
  }
}

// ---------------------------------------------------------------------------------------------
// bwd_edge: warp per destination row.
// per-warp dynamic smem: gs[HCp] | xs[kSlotBatch][HCp] | dal[kSlotBatch*H]
// ---------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(kThreads)
gat_bwd_edge_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ xh, int64_t ld,
                    int N, int H, int C, int S, int warps_per_block, int per_warp_floats,
                    const float* __restrict__ alpha, const float* __restrict__ amask,
                    const float* __restrict__ a_src, const float* __restrict__ a_dst, float slope,
                    const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ perm,
                    const float* __restrict__ ew, float* __restrict__ dr, float* __restrict__ da_dst,
                    float* __restrict__ dew) {
  extern __shared__ __align__(16) float smem[];
  const int HC = H * C;
  const int HCp = (HC + 3) & ~3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= warps_per_block) return;
  float* gs = smem + (size_t)warp * per_warp_floats;
  float* xs = gs + HCp;
  float* dal = xs + kSlotBatch * HCp;
  const int L = (C + S - 1) / S;
  const int units_per_slot = H * S;

  for (int i = blockIdx.x * warps_per_block + warp; i < N; i += gridDim.x * warps_per_block) {
    const int beg = rowptr[i], end = rowptr[i + 1];
    const int nslots = end - beg + 1;
    const int64_t slot0 = (int64_t)beg + i;
    {
      const float* gp = g + (int64_t)i * ldg;
      for (int f = lane * V; f < HC; f += 32 * V) {
        Vec<V> v = Vec<V>::load(gp + f);
#pragma unroll
        for (int u = 0; u < V; ++u) gs[f + u] = v.v[u];
      }
    }
    // sum_h[lane-th head (+32r)] = sum_k alpha * dalpha ; heads beyond 32 use the strided loop below
    float sum_reg[4] = {0.f, 0.f, 0.f, 0.f};  // supports H <= 128

    for (int kb = 0; kb < nslots; kb += kSlotBatch) {
      const int nk = min(kSlotBatch, nslots - kb);
      // stage the source rows of this batch of slots
      for (int k = 0; k < nk; ++k) {
        const int sl = kb + k;
        int j = (sl == nslots - 1) ? i : col[beg + sl];
        const float* xp = xh + (int64_t)j * ld;
        float* xd = xs + k * HCp;
        for (int f = lane * V; f < HC; f += 32 * V) {
          Vec<V> v = Vec<V>::load(xp + f);
#pragma unroll
          for (int u = 0; u < V; ++u) xd[f + u] = v.v[u];
        }
      }
      __syncwarp();
      const int units = nk * units_per_slot;
      for (int ub = 0; ub < units; ub += 32) {
        const int u = ub + lane;
        float dot = 0.f;
        int k = 0, h = 0, s = 0;
        if (u < units) {
          k = u / units_per_slot;
          const int r = u - k * units_per_slot;
          h = r / S;
          s = r - h * S;
          const int c0 = s * L;
          const int len = min(L, C - c0);
          if (len > 0) {
            int c = u % len;
            const float* xa = xs + k * HCp + h * C + c0;
            const float* ga = gs + h * C + c0;
            for (int t = 0; t < len; ++t) {
              dot = fmaf(ga[c], xa[c], dot);
              if (++c == len) c = 0;
            }
          }
        }
        for (int o = S >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if (u < units && s == 0) dal[k * H + h] = dot;
      }
      __syncwarp();
      // per head: mask / edge weight, accumulate sum_k alpha * dalpha, park dalpha in dr
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int h = r * 32 + lane;
        if (h < H) {
          for (int k = 0; k < nk; ++k) {
            const int sl = kb + k;
            const int64_t slot = slot0 + sl;
            const bool is_self = (sl == nslots - 1);
            float da = dal[k * H + h];
            const float a = alpha[slot * H + h];
            float a_used = a;
            if (amask != nullptr) a_used = a * amask[slot * H + h];
            if (!is_self && col[beg + sl] == i) da = 0.f;  // removed self loop
            if (dew != nullptr) dal[k * H + h] = a_used * da;  // contribution to d w_e (before w_e scaling)
            if (ew != nullptr && !is_self) da *= ew[perm[beg + sl]];
            if (amask != nullptr) da *= amask[slot * H + h];
            sum_reg[r] = fmaf(a, da, sum_reg[r]);
            dr[slot * H + h] = da;
          }
        }
      }
      if (dew != nullptr) {
        __syncwarp();
        if (lane == 0) {
          for (int k = 0; k < nk; ++k) {
            const int sl = kb + k;
            if (sl == nslots - 1) continue;
            float sacc = 0.f;
            for (int h = 0; h < H; ++h) sacc += dal[k * H + h];
            dew[perm[beg + sl]] = sacc;
          }
        }
      }
      __syncwarp();
    }
    // softmax Jacobian + leaky_relu' ; da_dst[i,h] = sum over slots (ascending edge id, self last)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int h = r * 32 + lane;
      if (h < H) {
        const float adst = a_dst[(int64_t)i * H + h];
        float acc = 0.f;
        for (int sl = 0; sl < nslots; ++sl) {
          const int64_t slot = slot0 + sl;
          const bool is_self = (sl == nslots - 1);
          const int j = is_self ? i : col[beg + sl];
          float d = 0.f;
          if (is_self || j != i) {
            const float a = alpha[slot * H + h];
            const float da = dr[slot * H + h];
            const float de = a * (da - sum_reg[r]);
            const float raw = a_src[(int64_t)j * H + h] + adst;
            d = raw > 0.f ? de : de * slope;
            acc = __fadd_rn(acc, d);
          }
          dr[slot * H + h] = d;
        }
        da_dst[(int64_t)i * H + h] = acc;
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// bwd_edge, block-streamed with HEAD-ALIGNED lanes (fast path; the staging kernel above is the fallback
// for H > 32 or very wide heads).
// P = 32 / H lanes share one head, each owning Q = ceil(C / P) consecutive channels, so the dot product
// <g_i[h,:], xh_j[h,:]> is Q FMAs per lane + a P-lane reduction and no shared-memory transposition (the
// staging kernel spent most of its issue slots there: 0.58 ms, 11 % of HBM peak).  Structure as in
// stream.cuh: a warp owns 32 consecutive destination rows, the slots (in-edges + self loop) of the block form
// one stream consumed G at a time with all gathers in flight; the lane-group leader keeps (alpha, d alpha,
// leaky_relu') of up to 8 slots of the current row in a shared-memory ring and finishes the softmax Jacobian
// when the stream passes the row end (longer rows spill to the dr buffer).
// ---------------------------------------------------------------------------------------------
namespace bes {
constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
constexpr int kList = 128;
constexpr int kRing = 8;

template <int QT, bool VEC4>
__global__ void __launch_bounds__(kThreads)
gat_bwd_edge_stream_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ xh, int64_t ld,
                           int N, int H, int C, int P, int Q, const float* __restrict__ alpha,
                           const float* __restrict__ amask, const float* __restrict__ a_src,
                           const float* __restrict__ a_dst, float slope, const int* __restrict__ rowptr,
                           const int* __restrict__ col, const int* __restrict__ perm,
                           const float* __restrict__ ew, float* __restrict__ dr, float* __restrict__ da_dst,
                           float* __restrict__ dew) {
  constexpr int G = QT <= 12 ? 4 : (QT <= 24 ? 2 : 1);
  __shared__ int s_src[kWarps][kList];
  __shared__ float s_w[kWarps][kList];
  __shared__ int s_eid[kWarps][kList];
  __shared__ float s_ring[kWarps][kRing][3][32];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hl = lane / P, pl = lane - hl * P;
  const bool active = hl < H;
  const bool leader = active && pl == 0;
  const int f0 = hl * C + pl * Q;
  const int nq = active ? max(0, min(Q, C - pl * Q)) : 0;
  const bool pow2 = (P & (P - 1)) == 0;
  const int nblocks = (N + 31) / 32;
  const int nwarps = gridDim.x * kWarps;

  auto load_slice = [&](const float* rowp, float (&dst)[QT]) {
    if (VEC4) {
#pragma unroll
      for (int q = 0; q < QT; q += 4) {
        if (q < nq) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(rowp + f0 + q));
          dst[q] = v.x; dst[q + 1] = v.y; dst[q + 2] = v.z; dst[q + 3] = v.w;
        }
      }
    } else {
#pragma unroll
      for (int q = 0; q < QT; ++q)
        if (q < nq) dst[q] = __ldg(rowp + f0 + q);
    }
  };
  auto group_sum = [&](float v) {   // sum over the P lanes of a head; valid in the leader
    if (pow2) {
      for (int o = P >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    } else {
      float t = v;
      for (int o = 1; o < P; ++o) t += __shfl_down_sync(0xffffffffu, v, o);
      v = t;
    }
    return v;
  };

  for (int bb = blockIdx.x * kWarps + warp; bb < nblocks; bb += nwarps) {
    const int i0 = (nblocks - 1 - bb) * 32;
    const int nrows = min(32, N - i0);
    const int my = i0 + lane;
    int beg = 0, len = 0;
    if (lane < nrows) {
      beg = __ldg(rowptr + my);
      len = __ldg(rowptr + my + 1) - beg + 1;                 // in-edges + self loop
    }
    int ve = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, ve, o);
      if (lane >= o) ve += t;
    }
    const int vs = ve - len;
    const int total = __shfl_sync(0xffffffffu, ve, 31);

    // per-row state (uniform across the warp unless noted)
    int r = 0;
    int vend_r = __shfl_sync(0xffffffffu, ve, 0);
    int vbeg_r = 0;
    int beg_r = __shfl_sync(0xffffffffu, beg, 0);
    float gq[QT];
    load_slice(g + (int64_t)i0 * ldg, gq);
    float adst = leader ? __ldg(a_dst + (int64_t)i0 * H + hl) : 0.f;
    float sum = 0.f;                                          // leader: sum_k alpha_k * dalpha_k of the current row

    auto finalize_row = [&]() {
      const int i = i0 + r;
      const int nslots = vend_r - vbeg_r;
      if (leader) {
        float acc = 0.f;
        for (int k = 0; k < nslots; ++k) {
          const int64_t slot = (int64_t)beg_r + i + k;
          float a, da, fac;
          if (k < kRing) {
            a = s_ring[warp][k][0][lane]; da = s_ring[warp][k][1][lane]; fac = s_ring[warp][k][2][lane];
          } else {                                            // long row: recompute from the spilled d alpha
            const int j = (k == nslots - 1) ? i : __ldg(col + beg_r + k);
            a = __ldg(alpha + slot * H + hl);
            da = dr[slot * H + hl];
            const float raw = __ldg(a_src + (int64_t)j * H + hl) + adst;
            fac = (k != nslots - 1 && j == i) ? 0.f : (raw > 0.f ? 1.f : slope);
          }
          const float d = a * (da - sum) * fac;
          dr[slot * H + hl] = d;
          acc = __fadd_rn(acc, d);
        }
        da_dst[(int64_t)i * H + hl] = acc;
      }
      __syncwarp();
    };

    for (int w0 = 0; w0 < total; w0 += kList) {
      const int w1 = min(total, w0 + kList);
      __syncwarp();
      for (int t = max(vs, w0); t < min(ve, w1); ++t) {       // every lane lists the slots of its own row
        const int k = t - vs;
        int j, e = -1;
        float w = 1.f;
        if (k == len - 1) {
          j = my;
        } else {
          j = __ldg(col + beg + k);
          if (ew != nullptr || dew != nullptr) e = __ldg(perm + beg + k);
          if (ew != nullptr) w = __ldg(ew + e);
          if (j == my) j = -1;                                // pre-existing self loop: removed by GATConv
        }
        s_src[warp][t - w0] = j;
        s_w[warp][t - w0] = w;
        s_eid[warp][t - w0] = e;
      }
      __syncwarp();
      const int wl = w1 - w0;
      for (int t = 0; t < wl; t += G) {
        int j[G];
        float xq[G][QT], al[G], as[G], mk[G];
#pragma unroll
        for (int k = 0; k < G; ++k) j[k] = (t + k < wl) ? s_src[warp][t + k] : -2;
#pragma unroll
        for (int k = 0; k < G; ++k) {
          al[k] = 0.f; as[k] = 0.f; mk[k] = 1.f;
          if (j[k] >= 0) {
            load_slice(xh + (int64_t)j[k] * ld, xq[k]);
            if (leader) as[k] = __ldg(a_src + (int64_t)j[k] * H + hl);
          }
        }
#pragma unroll
        for (int k = 0; k < G; ++k) {
          if (t + k >= wl) break;
          const int vt = w0 + t + k;
          while (vt >= vend_r) {                              // the stream passed the end of row r
            finalize_row();
            ++r;
            vbeg_r = vend_r;
            vend_r = __shfl_sync(0xffffffffu, ve, min(r, 31));
            beg_r = __shfl_sync(0xffffffffu, beg, min(r, 31));
            load_slice(g + (int64_t)(i0 + r) * ldg, gq);
            adst = leader ? __ldg(a_dst + (int64_t)(i0 + r) * H + hl) : 0.f;
            sum = 0.f;
          }
          const int i = i0 + r;
          const int kin = vt - vbeg_r;                        // slot index inside the row
          const int64_t slot = (int64_t)beg_r + i + kin;
          float dot = 0.f;
          if (j[k] >= 0) {
#pragma unroll
            for (int q = 0; q < QT; ++q)
              if (q < nq) dot = fmaf(gq[q], xq[k][q], dot);
          }
          dot = group_sum(dot);
          float a = 0.f, m = 1.f;
          if (leader) {
            a = __ldg(alpha + slot * H + hl);
            if (amask != nullptr) m = __ldg(amask + slot * H + hl);
          }
          const float w = s_w[warp][t + k];
          if (dew != nullptr) {                               // d w_e = sum_h alpha_used * <g_i, xh_j>
            float c = leader && j[k] >= 0 ? a * m * dot : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            const int e = s_eid[warp][t + k];
            if (lane == 0 && e >= 0) dew[e] = c;
          }
          if (leader) {
            const float da = (j[k] >= 0) ? dot * w * m : 0.f;
            const float raw = as[k] + adst;
            const float fac = (j[k] >= 0) ? (raw > 0.f ? 1.f : slope) : 0.f;
            sum = fmaf(a, da, sum);
            if (kin < kRing) {
              s_ring[warp][kin][0][lane] = a; s_ring[warp][kin][1][lane] = da; s_ring[warp][kin][2][lane] = fac;
            } else {
              dr[slot * H + hl] = da;
            }
          }
        }
      }
    }
    while (r < nrows) {                                       // last row(s) of the block
      finalize_row();
      ++r;
      if (r < nrows) {
        vbeg_r = vend_r;
        vend_r = __shfl_sync(0xffffffffu, ve, r);
        beg_r = __shfl_sync(0xffffffffu, beg, r);
        sum = 0.f;
      }
    }
  }
}
}  // namespace bes

// ---------------------------------------------------------------------------------------------
// bwd_node part 1: da_src[j,h] = sum over out-slots of dr (ascending edge id, self last)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
gat_bwd_dasrc_kernel(const float* __restrict__ dr, int N, int H, const int* __restrict__ rowptr,
                     const int* __restrict__ colptr, const int* __restrict__ row,
                     const int* __restrict__ csc_pos, float* __restrict__ da_src) {
  const int64_t total = (int64_t)N * H;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (int64_t)gridDim.x * kThreads) {
    const int j = (int)(t / H);
    const int h = (int)(t - (int64_t)j * H);
    const int beg = __ldg(colptr + j), end = __ldg(colptr + j + 1);
    float acc = 0.f;
    for (int q = beg; q < end; ++q) {
      const int i = __ldg(row + q);
      if (i == j) continue;
      acc = __fadd_rn(acc, __ldg(dr + ((int64_t)__ldg(csc_pos + q) + i) * H + h));
    }
    acc = __fadd_rn(acc, __ldg(dr + ((int64_t)__ldg(rowptr + j + 1) + j) * H + h));
    da_src[t] = acc;
  }
}

// bwd_node part 2: dxh[j, f] (thread per (source j, V-wide chunk))
template <int V, bool WEIGHTED>
__global__ void __launch_bounds__(kThreads)
gat_bwd_node_kernel(const float* __restrict__ g, int64_t ldg, int N, int H, int C, int chunks,
                    const float* __restrict__ alpha_used, const float* __restrict__ da_src,
                    const float* __restrict__ da_dst, const float* __restrict__ att_src,
                    const float* __restrict__ att_dst, const int* __restrict__ rowptr,
                    const int* __restrict__ colptr, const int* __restrict__ row,
                    const int* __restrict__ csc_pos, const int* __restrict__ permt,
                    const float* __restrict__ ew, float* __restrict__ dxh, int64_t lddxh) {
  const int64_t total = (int64_t)N * chunks;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (int64_t)gridDim.x * kThreads) {
    const int j = (int)(t / chunks);
    const int f = (int)(t - (int64_t)j * chunks) * V;
    int hh[V];
#pragma unroll
    for (int u = 0; u < V; ++u) hh[u] = (f + u) / C;
    const int beg = __ldg(colptr + j), end = __ldg(colptr + j + 1);
    Vec<V> acc = vzero<V>();
    for (int q = beg; q <= end; q += 4) {
      int i[4];
      int64_t slot[4];
      Vec<V> v[4];
      float a[4][V];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int qk = q + k;
        if (qk < end) {
          i[k] = __ldg(row + qk);
          slot[k] = (int64_t)__ldg(csc_pos + qk) + i[k];
          if (i[k] == j) i[k] = -1;
        } else if (qk == end) {
          i[k] = j;
          slot[k] = (int64_t)__ldg(rowptr + j + 1) + j;
        } else {
          i[k] = -1;
          slot[k] = 0;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (i[k] >= 0) {
          v[k] = Vec<V>::load(g + (int64_t)i[k] * ldg + f);
          const float* ak = alpha_used + slot[k] * H;
#pragma unroll
          for (int u = 0; u < V; ++u)
            a[k][u] = (u > 0 && hh[u] == hh[u - 1]) ? a[k][u - 1] : __ldg(ak + hh[u]);
          if (WEIGHTED && q + k < end) {
            const float w = __ldg(ew + __ldg(permt + q + k));
#pragma unroll
            for (int u = 0; u < V; ++u) a[k][u] = __fmul_rn(a[k][u], w);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (i[k] >= 0) {
#pragma unroll
          for (int u = 0; u < V; ++u) acc.v[u] = __fadd_rn(acc.v[u], __fmul_rn(a[k][u], v[k].v[u]));
        }
      }
    }
    if (att_src != nullptr) {
#pragma unroll
      for (int u = 0; u < V; ++u) {
        const float ds = __ldg(da_src + (int64_t)j * H + hh[u]);
        const float dd = __ldg(da_dst + (int64_t)j * H + hh[u]);
        acc.v[u] = fmaf(ds, __ldg(att_src + f + u), acc.v[u]);
        acc.v[u] = fmaf(dd, __ldg(att_dst + f + u), acc.v[u]);
      }
    }
    acc.store(dxh + (int64_t)j * lddxh + f);
  }
}

// ---------------------------------------------------------------------------------------------
// bwd_att: datt_src[f] = sum_n da_src[n, f / C] * xh[n, f]  (stage 1: per-CTA partials; stage 2: fixed-order sum)
// ---------------------------------------------------------------------------------------------
constexpr int kAttParts = 592;  // 4 CTAs per SM on 148 SMs

// stage 1, warp-per-row mapping: each warp streams whole xh rows (coalesced, 2 rows in flight) and keeps
// per-lane partial sums of da_src[n,h] * xh[n,f] and da_dst[n,h] * xh[n,f]; warps are combined through
// shared memory in fixed order (deterministic).  Reads xh exactly once.
template <int V, int ITERS>
__global__ void __launch_bounds__(kThreads)
gat_bwd_att_partial_row_kernel(const float* __restrict__ xh, int64_t ld, int N, int H, int C, int chunks,
                               const float* __restrict__ da_src, const float* __restrict__ da_dst,
                               float* __restrict__ part /* [2][gridDim.x][HC] */) {
  extern __shared__ float sm[];  // [2][kWarps][HC]
  const int HC = H * C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int hh[ITERS][V];
#pragma unroll
  for (int t = 0; t < ITERS; ++t)
#pragma unroll
    for (int u = 0; u < V; ++u) hh[t][u] = min(((lane + 32 * t) * V + u) / C, H - 1);
  Vec<V> as[ITERS], ad[ITERS];
#pragma unroll
  for (int t = 0; t < ITERS; ++t) { as[t] = vzero<V>(); ad[t] = vzero<V>(); }
  const int stride = gridDim.x * kWarps;
  for (int n0 = blockIdx.x * kWarps + warp; n0 < N; n0 += 2 * stride) {
    Vec<V> v[2][ITERS];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int n = n0 + k * stride;
      if (n < N) {
#pragma unroll
        for (int t = 0; t < ITERS; ++t)
          if (lane + 32 * t < chunks) v[k][t] = Vec<V>::load(xh + (int64_t)n * ld + (lane + 32 * t) * V);
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int n = n0 + k * stride;
      if (n < N) {
        const float* ds = da_src + (int64_t)n * H;
        const float* dd = da_dst + (int64_t)n * H;
#pragma unroll
        for (int t = 0; t < ITERS; ++t)
          if (lane + 32 * t < chunks) {
#pragma unroll
            for (int u = 0; u < V; ++u) {
              as[t].v[u] = fmaf(__ldg(ds + hh[t][u]), v[k][t].v[u], as[t].v[u]);
              ad[t].v[u] = fmaf(__ldg(dd + hh[t][u]), v[k][t].v[u], ad[t].v[u]);
            }
          }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < ITERS; ++t)
    if (lane + 32 * t < chunks)
#pragma unroll
      for (int u = 0; u < V; ++u) {
        sm[warp * HC + (lane + 32 * t) * V + u] = as[t].v[u];
        sm[(kWarps + warp) * HC + (lane + 32 * t) * V + u] = ad[t].v[u];
      }
  __syncthreads();
  for (int f = threadIdx.x; f < HC; f += kThreads) {
    float s = 0.f, d = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { s += sm[w * HC + f]; d += sm[(kWarps + w) * HC + f]; }
    part[(int64_t)blockIdx.x * HC + f] = s;
    part[((int64_t)gridDim.x + blockIdx.x) * HC + f] = d;
  }
}

// fallback (very wide rows): thread per feature
__global__ void __launch_bounds__(kThreads)
gat_bwd_att_partial_kernel(const float* __restrict__ xh, int64_t ld, int N, int H, int C,
                           const float* __restrict__ da_src, const float* __restrict__ da_dst,
                           float* __restrict__ part /* [2][gridDim.x][HC] */) {
  const int HC = H * C;
  for (int f0 = 0; f0 < HC; f0 += kThreads) {
    const int f = f0 + threadIdx.x;
    if (f >= HC) break;
    const int h = f / C;
    float ps = 0.f, pd = 0.f;
    for (int n = blockIdx.x; n < N; n += gridDim.x) {
      const float xv = __ldg(xh + (int64_t)n * ld + f);
      ps = fmaf(__ldg(da_src + (int64_t)n * H + h), xv, ps);
      pd = fmaf(__ldg(da_dst + (int64_t)n * H + h), xv, pd);
    }
    part[(int64_t)blockIdx.x * HC + f] = ps;
    part[((int64_t)gridDim.x + blockIdx.x) * HC + f] = pd;
  }
}

// stage 2: 32 features x 8 partial-slices per CTA, fixed summation order (deterministic)
__global__ void __launch_bounds__(kThreads)
gat_bwd_att_final_kernel(const float* __restrict__ part, int parts, int HC,
                         float* __restrict__ datt_src, float* __restrict__ datt_dst) {
  __shared__ float sm[2][8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int f = blockIdx.x * 32 + tx;
  float s = 0.f, d = 0.f;
  if (f < HC)
    for (int p = ty; p < parts; p += 8) {
      s += part[(int64_t)p * HC + f];
      d += part[((int64_t)parts + p) * HC + f];
    }
  sm[0][ty][tx] = s;
  sm[1][ty][tx] = d;
  __syncthreads();
  if (ty == 0 && f < HC) {
    float ts = 0.f, td = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { ts += sm[0][k][tx]; td += sm[1][k][tx]; }
    datt_src[f] = ts;
    datt_dst[f] = td;
  }
}

}  // namespace
}  // namespace mgs

using namespace mgs;

#define MGS_GAT_COMMON_CHECKS(NAME)                                                                     \
  MGS_REQUIRE(num_nodes >= 0 && num_nodes < 0x7fffffff, NAME ": bad num_nodes");                        \
  MGS_REQUIRE(heads > 0 && channels > 0 && (int64_t)heads * channels < (1 << 20), NAME ": bad heads/channels"); \
  MGS_REQUIRE(heads <= 128, NAME ": heads > 128 unsupported")

extern "C" int mgs_gat_scores_fwd(const float* xh, int64_t ld, int64_t num_nodes, int32_t heads, int32_t channels,
                                  const float* att_src, const float* att_dst, float* a_src, float* a_dst,
                                  mgs_stream_t stream_) {
  MGS_GAT_COMMON_CHECKS("mgs_gat_scores_fwd");
  const int HC = heads * channels;
  MGS_REQUIRE(ld >= HC, "mgs_gat_scores_fwd: leading dimension < heads*channels");
  if (num_nodes == 0) return MGS_OK;
  MGS_REQUIRE(xh && att_src && att_dst && a_src && a_dst, "mgs_gat_scores_fwd: null pointer");
  const int HCp = (HC + 3) & ~3;
  const size_t smem = sizeof(float) * (size_t)HCp * (2 + kWarps);
  if (smem > 200 * 1024) {
    set_error("mgs_gat_scores_fwd: heads*channels=%d too large for shared memory staging", HC);
    return MGS_ERR_UNSUPPORTED;
  }
  const int V = vec_width(xh, ld, HC);
  const int S = pick_split(heads, channels);
  const int grid = grid_for(num_nodes * 32, kThreads, 4);
  cudaStream_t stream = (cudaStream_t)stream_;
#define MGS_SCORES(VV)                                                                                     \
  do {                                                                                                     \
    if (smem > 48 * 1024)                                                                                  \
      MGS_CUDA(cudaFuncSetAttribute(gat_scores_kernel<VV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    gat_scores_kernel<VV><<<grid, kThreads, smem, stream>>>(xh, ld, (int)num_nodes, heads, channels, S,    \
                                                            att_src, att_dst, a_src, a_dst);               \
  } while (0)
  if (V == 4) MGS_SCORES(4); else if (V == 2) MGS_SCORES(2); else MGS_SCORES(1);
#undef MGS_SCORES
  return check_launch("gat_scores_kernel");
}

extern "C" int mgs_gat_alpha_fwd(const float* a_src, const float* a_dst, int64_t num_nodes, int32_t heads,
                                 const int32_t* rowptr, const int32_t* col, float negative_slope,
                                 float* alpha, mgs_stream_t stream_) {
  MGS_REQUIRE(num_nodes >= 0 && num_nodes < 0x7fffffff && heads > 0, "mgs_gat_alpha_fwd: bad sizes");
  if (num_nodes == 0) return MGS_OK;
  MGS_REQUIRE(a_src && a_dst && rowptr && alpha, "mgs_gat_alpha_fwd: null pointer");
  const int grid = grid_for(num_nodes * heads, kThreads, 8);
  gat_alpha_kernel<<<grid, kThreads, 0, (cudaStream_t)stream_>>>(a_src, a_dst, (int)num_nodes, heads, rowptr, col,
                                                                  negative_slope, alpha);
  return check_launch("gat_alpha_kernel");
}

extern "C" int mgs_gat_aggr_fwd(const float* xh, int64_t ld, int64_t num_nodes, int32_t heads, int32_t channels,
                                const float* alpha_used, const int32_t* rowptr, const int32_t* col,
                                const int32_t* perm, const float* edge_weight, const float* bias, float* out,
                                int64_t ldo, int32_t activation, uint32_t* relu_bits, int32_t bits_words,
                                mgs_stream_t stream_) {
  MGS_GAT_COMMON_CHECKS("mgs_gat_aggr_fwd");
  MGS_REQUIRE(activation >= 0 && activation <= 2, "mgs_gat_aggr_fwd: activation must be 0 (none), 1 (ReLU) or 2 (ELU)");
  const int HC = heads * channels;
  MGS_REQUIRE(ld >= HC && ldo >= HC, "mgs_gat_aggr_fwd: leading dimension < heads*channels");
  if (num_nodes == 0) return MGS_OK;
  MGS_REQUIRE(xh && alpha_used && rowptr && out, "mgs_gat_aggr_fwd: null pointer");
  MGS_REQUIRE(!edge_weight || perm, "mgs_gat_aggr_fwd: edge_weight needs perm");
  int V = min_int(vec_width(xh, ld, HC), vec_width(out, ldo, HC));
  if (bias) V = min_int(V, vec_width(bias, HC, HC));
  const int chunks = HC / V;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int iters = iters_for(chunks);
  if (iters > 0 && heads <= 32) {   // block-streamed fast path (stream.cuh): alpha travels by warp shuffle
    stream::Args sa = {};
    sa.src = xh; sa.lds = ld; sa.dst = out; sa.ldd = ldo;
    sa.N = (int)num_nodes; sa.chunks = chunks; sa.H = heads; sa.C = channels;
    sa.ptr = rowptr; sa.idx = col; sa.eid = perm; sa.ew = edge_weight; sa.rowptr = rowptr;
    sa.alpha = alpha_used; sa.bias = bias; sa.activation = activation;
    if (relu_bits != nullptr) {
      MGS_REQUIRE(activation == 1 && bits_words == V * iters,
                  "mgs_gat_aggr_fwd: relu_bits needs the ReLU epilogue and bits_words == V * iterations (%d here)", V * iters);
      sa.bits_out = relu_bits;
    }
    return stream::launch<stream::GAT_FWD>(sa, V, iters, stream, "gat_aggr_fwd(stream)");
  }
  MGS_REQUIRE(activation == 0 && relu_bits == nullptr, "mgs_gat_aggr_fwd: fused activation needs heads <= 32 and rows of <= 256 vector chunks");
  const int grid = grid_for(num_nodes * chunks, kThreads, 8);
#define MGS_AGGR(VV, WW)                                                                                  \
  gat_aggr_fwd_kernel<VV, WW><<<grid, kThreads, 0, stream>>>(xh, ld, (int)num_nodes, heads, channels, chunks, \
                                                             alpha_used, rowptr, col, perm, edge_weight, bias, out, ldo)
  if (edge_weight) { if (V == 4) MGS_AGGR(4, true); else if (V == 2) MGS_AGGR(2, true); else MGS_AGGR(1, true); }
  else { if (V == 4) MGS_AGGR(4, false); else if (V == 2) MGS_AGGR(2, false); else MGS_AGGR(1, false); }
#undef MGS_AGGR
  return check_launch("gat_aggr_fwd_kernel");
}

extern "C" int mgs_gat_bwd_edge(const float* g, int64_t ldg, const float* xh, int64_t ld, int64_t num_nodes,
                                int32_t heads, int32_t channels, const float* alpha, const float* alpha_mask,
                                const float* a_src, const float* a_dst, float negative_slope,
                                const int32_t* rowptr, const int32_t* col, const int32_t* perm,
                                const float* edge_weight, float* dr, float* da_dst, float* d_edge_weight,
                                mgs_stream_t stream_) {
  MGS_GAT_COMMON_CHECKS("mgs_gat_bwd_edge");
  const int HC = heads * channels;
  MGS_REQUIRE(ld >= HC && ldg >= HC, "mgs_gat_bwd_edge: leading dimension < heads*channels");
  if (num_nodes == 0) return MGS_OK;
  MGS_REQUIRE(g && xh && alpha && a_src && a_dst && rowptr && dr && da_dst, "mgs_gat_bwd_edge: null pointer");
  MGS_REQUIRE((!edge_weight && !d_edge_weight) || perm, "mgs_gat_bwd_edge: edge weights need perm");
  {   // tensor-core path (edge_mma.cuh): per-head sums as mma.sync against the head-indicator matrix
    const char* e = std::getenv("MGS_EDGE_MMA");                    // read per call: tests / probes toggle it
    const bool plain = edge_weight == nullptr && d_edge_weight == nullptr && alpha_mask == nullptr;
    if (!(e && e[0] == '0') && plain && heads <= 16 && HC % 2 == 0 && HC <= 16384 && (uintptr_t)g % 8 == 0 && (uintptr_t)xh % 8 == 0 &&
        ldg % 2 == 0 && ld % 2 == 0 && (int64_t)num_nodes * 64 * heads < (1ll << 31)) {
      cudaStream_t st = (cudaStream_t)stream_;
      const int grid = sm_count() * 4;                              // 32 warps per SM, each walking tiles of 16 slots
      const int64_t pad4 = (HC + 3) & ~3;
      const bool wide = (uintptr_t)g % 16 == 0 && (uintptr_t)xh % 16 == 0 && ldg % 4 == 0 && ld % 4 == 0 && ldg >= pad4 &&
                        ld >= pad4;                               // 128-bit loads: the last one may cover padding floats
      const size_t hsmem = (size_t)((HC + 15) / 16 * 16 + 16);     // head-of-column table, one byte per (padded) column
      // compile-time shapes (the reference's GATConv(35, 35, heads=10) and the stress trunk's (35, 32, heads=8)): plain
      // FMAs with constant head boundaries, a quarter of the instructions (edge_fma.cuh)
      const char* ef = std::getenv("MGS_EDGE_FMA");                 // 0: off; 2 / 3 / 4: loads in flight (probes)
      const int fma_d = ef ? std::atoi(ef) : 3;
      bool fma_done = false;
      if (wide && fma_d > 0) {
        const int fgrid = sm_count() * 2;
#define MGS_EFMA(HV, CV, DV) efma::gat_bwd_edge_fma_kernel<HV, CV, DV><<<fgrid, efma::kThreads, 0, st>>>(               \
      g, ldg, xh, ld, (int)num_nodes, rowptr, col, dr)
#define MGS_EFMA_D(HV, CV) do { if (fma_d == 2) MGS_EFMA(HV, CV, 2); else if (fma_d == 4) MGS_EFMA(HV, CV, 4);         \
                                else MGS_EFMA(HV, CV, 3); fma_done = true; } while (0)
        if (heads == 10 && channels == 35) MGS_EFMA_D(10, 35);
        else if (heads == 8 && channels == 32) MGS_EFMA_D(8, 32);
#undef MGS_EFMA_D
#undef MGS_EFMA
        if (fma_done) { if (int rc = check_launch("gat_bwd_edge_fma_kernel")) return rc; }
      }
      if (!fma_done) {
#define MGS_EMMA(NTV, WV) emma::gat_bwd_edge_mma_kernel<NTV, WV><<<grid, emma::kThreads, hsmem, st>>>(              \
      g, ldg, xh, ld, (int)num_nodes, heads, channels, rowptr, col, dr)
        if (heads <= 8) { if (wide) MGS_EMMA(1, 4); else MGS_EMMA(1, 2); }
        else { if (wide) MGS_EMMA(2, 4); else MGS_EMMA(2, 2); }
#undef MGS_EMMA
        if (int rc = check_launch("gat_bwd_edge_mma_kernel")) return rc;
      }
      emma::gat_bwd_edge_softmax_kernel<<<grid_for((int64_t)num_nodes * heads, 256, 8), 256, 0, st>>>(
          alpha, a_src, a_dst, negative_slope, rowptr, col, (int)num_nodes, heads, dr, da_dst);
      return check_launch("gat_bwd_edge_softmax_kernel");
    }
  }
  {   // staged fast path (edge.cuh): coalesced gathers, head-aligned read-back from shared memory
    const int P = heads <= 32 ? 32 / heads : 0;
    const int Q = P > 0 ? (channels + P - 1) / P : 0;
    const int V = min_int(vec_width(xh, ld, HC), vec_width(g, ldg, HC));
    const int chunks = HC / (V > 0 ? V : 1);
    const int iters = iters_for(chunks);
    const bool small = (int64_t)num_nodes * (ld > ldg ? ld : ldg) < (1ll << 31) && ((int64_t)num_nodes * 64 + 1) * heads < (1ll << 31);
    // 32-bit indices inside: rows * ld and (E + N) * H must fit (E <= 63 N assumed; the per-slot arrays of larger
    // graphs would not fit 32 bits anyway) -- else the older kernels below take over
    if (P > 0 && Q <= 12 && V >= 2 && iters > 0 && small) {
      edge::Args ea = {};
      ea.g = g; ea.ldg = ldg; ea.xh = xh; ea.ld = ld;
      ea.N = (int)num_nodes; ea.H = heads; ea.C = channels; ea.P = P; ea.Q = Q; ea.chunks = chunks;
      edge::pick_order(heads, channels, P, Q, &ea.interleaved, &ea.rot_a, &ea.rot_b);
      ea.alpha = alpha; ea.amask = alpha_mask; ea.a_src = a_src; ea.a_dst = a_dst; ea.slope = negative_slope;
      ea.rowptr = rowptr; ea.col = col; ea.perm = perm; ea.ew = edge_weight;
      ea.dr = dr; ea.da_dst = da_dst; ea.dew = d_edge_weight;
      const int egrid = grid_for_rows(num_nodes, edge::kWarps, 2);
      cudaStream_t st = (cudaStream_t)stream_;
#define MGS_EDGE3(VV, II, QQ, XX)                                                                           \
  do {                                                                                                      \
    MGS_CUDA(cudaFuncSetAttribute(edge::gat_bwd_edge_kernel<VV, II, QQ, XX>,                                \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
    edge::gat_bwd_edge_kernel<VV, II, QQ, XX><<<egrid, edge::kThreads, smem, st>>>(ea);                     \
  } while (0)
#define MGS_EDGE(VV, II)                                                                                    \
  do {                                                                                                      \
    const size_t smem = edge::smem_bytes<VV, II>();                                                         \
    if (Q <= 4) { if (extra) MGS_EDGE3(VV, II, 4, true); else MGS_EDGE3(VV, II, 4, false); }                \
    else { if (extra) MGS_EDGE3(VV, II, 12, true); else MGS_EDGE3(VV, II, 12, false); }                     \
  } while (0)
      const bool extra = edge_weight != nullptr || d_edge_weight != nullptr || alpha_mask != nullptr;
      if (V == 4) {
        if (iters == 1) MGS_EDGE(4, 1); else if (iters == 2) MGS_EDGE(4, 2); else if (iters == 4) MGS_EDGE(4, 4);
        else if (iters == 6) MGS_EDGE(4, 6); else MGS_EDGE(4, 8);
      } else {
        if (iters == 1) MGS_EDGE(2, 1); else if (iters == 2) MGS_EDGE(2, 2); else if (iters == 4) MGS_EDGE(2, 4);
        else if (iters == 6) MGS_EDGE(2, 6); else MGS_EDGE(2, 8);
      }
#undef MGS_EDGE3
#undef MGS_EDGE
      return check_launch("gat_bwd_edge_kernel(staged)");
    }
  }
  if (heads <= 32) {   // head-aligned block-streamed path with direct (strided) loads
    const int P = 32 / heads;
    const int Q = (channels + P - 1) / P;
    if (Q <= 32) {
      const bool v4 = channels % 4 == 0 && Q % 4 == 0 && vec_width(xh, ld, HC) == 4 && vec_width(g, ldg, HC) == 4;
      const int nblocks = (int)((num_nodes + 31) / 32);
      const int sgrid = grid_for((int64_t)nblocks * 32, bes::kThreads, 8);
      cudaStream_t st = (cudaStream_t)stream_;
#define MGS_BES(QT, V4)                                                                                    \
  bes::gat_bwd_edge_stream_kernel<QT, V4><<<sgrid, bes::kThreads, 0, st>>>(                                \
      g, ldg, xh, ld, (int)num_nodes, heads, channels, P, Q, alpha, alpha_mask, a_src, a_dst, negative_slope, \
      rowptr, col, perm, edge_weight, dr, da_dst, d_edge_weight)
      if (v4) {
        if (Q <= 4) MGS_BES(4, true); else if (Q <= 8) MGS_BES(8, true); else if (Q <= 16) MGS_BES(16, true);
        else MGS_BES(32, true);
      } else {
        if (Q <= 4) MGS_BES(4, false); else if (Q <= 8) MGS_BES(8, false); else if (Q <= 12) MGS_BES(12, false);
        else if (Q <= 16) MGS_BES(16, false); else MGS_BES(32, false);
      }
#undef MGS_BES
      return check_launch("gat_bwd_edge_stream_kernel");
    }
  }
  const int HCp = (HC + 3) & ~3;
  const int per_warp = HCp * (1 + kSlotBatch) + ((kSlotBatch * heads + 3) & ~3);
  int warps = kWarps;
  while (warps > 1 && sizeof(float) * (size_t)per_warp * warps > 100 * 1024) warps >>= 1;
  const size_t smem = sizeof(float) * (size_t)per_warp * warps;
  if (smem > 200 * 1024) {
    set_error("mgs_gat_bwd_edge: heads*channels=%d too large for shared memory staging", HC);
    return MGS_ERR_UNSUPPORTED;
  }
  const int V = min_int(vec_width(xh, ld, HC), vec_width(g, ldg, HC));
  const int S = pick_split(heads, channels);
  const int grid = grid_for((num_nodes + warps - 1) / warps * kThreads, kThreads, 4);
  cudaStream_t stream = (cudaStream_t)stream_;
#define MGS_BWD_EDGE(VV)                                                                                    \
  do {                                                                                                      \
    if (smem > 48 * 1024)                                                                                   \
      MGS_CUDA(cudaFuncSetAttribute(gat_bwd_edge_kernel<VV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    gat_bwd_edge_kernel<VV><<<grid, kThreads, smem, stream>>>(g, ldg, xh, ld, (int)num_nodes, heads, channels, S, \
        warps, per_warp, alpha, alpha_mask, a_src, a_dst, negative_slope, rowptr, col, perm, edge_weight, dr,  \
        da_dst, d_edge_weight);                                                                             \
  } while (0)
  if (V == 4) MGS_BWD_EDGE(4); else if (V == 2) MGS_BWD_EDGE(2); else MGS_BWD_EDGE(1);
#undef MGS_BWD_EDGE
  return check_launch("gat_bwd_edge_kernel");
}

extern "C" int mgs_gat_bwd_node(const float* g, int64_t ldg, int64_t num_nodes, int32_t heads, int32_t channels,
                                const float* alpha_used, const float* dr, const float* da_dst,
                                const float* att_src, const float* att_dst, const int32_t* rowptr,
                                const int32_t* colptr, const int32_t* row, const int32_t* csc_pos,
                                const int32_t* permt, const float* edge_weight, float* dxh, int64_t lddxh,
                                float* da_src, mgs_stream_t stream_) {
  MGS_GAT_COMMON_CHECKS("mgs_gat_bwd_node");
  const int HC = heads * channels;
  MGS_REQUIRE(ldg >= HC && lddxh >= HC, "mgs_gat_bwd_node: leading dimension < heads*channels");
  if (num_nodes == 0) return MGS_OK;
  MGS_REQUIRE(g && alpha_used && dr && da_dst && rowptr && colptr && dxh && da_src, "mgs_gat_bwd_node: null pointer");
  MGS_REQUIRE((att_src == nullptr) == (att_dst == nullptr), "mgs_gat_bwd_node: att_src / att_dst go together");
  MGS_REQUIRE(!edge_weight || permt, "mgs_gat_bwd_node: edge_weight needs permt");
  cudaStream_t stream = (cudaStream_t)stream_;
  gat_bwd_dasrc_kernel<<<grid_for(num_nodes * heads, kThreads, 8), kThreads, 0, stream>>>(
      dr, (int)num_nodes, heads, rowptr, colptr, row, csc_pos, da_src);
  if (int rc = check_launch("gat_bwd_dasrc_kernel")) return rc;
  int V = min_int(vec_width(g, ldg, HC), vec_width(dxh, lddxh, HC));
  if (att_src) V = min_int(V, min_int(vec_width(att_src, HC, HC), vec_width(att_dst, HC, HC)));
  const int chunks = HC / V;
  const int iters = iters_for(chunks);
  if (iters > 0 && heads <= 32) {
    stream::Args sa = {};
    sa.src = g; sa.lds = ldg; sa.dst = dxh; sa.ldd = lddxh;
    sa.N = (int)num_nodes; sa.chunks = chunks; sa.H = heads; sa.C = channels;
    sa.ptr = colptr; sa.idx = row; sa.eid = permt; sa.ew = edge_weight; sa.rowptr = rowptr; sa.csc_pos = csc_pos;
    sa.alpha = alpha_used; sa.da_src = da_src; sa.da_dst = da_dst; sa.att_src = att_src; sa.att_dst = att_dst;
    return stream::launch<stream::GAT_BWD_NODE>(sa, V, iters, stream, "gat_bwd_node(stream)");
  }
  const int grid = grid_for(num_nodes * chunks, kThreads, 8);
#define MGS_NODE(VV, WW)                                                                                   \
  gat_bwd_node_kernel<VV, WW><<<grid, kThreads, 0, stream>>>(g, ldg, (int)num_nodes, heads, channels, chunks, \
      alpha_used, da_src, da_dst, att_src, att_dst, rowptr, colptr, row, csc_pos, permt, edge_weight, dxh, lddxh)
  if (edge_weight) { if (V == 4) MGS_NODE(4, true); else if (V == 2) MGS_NODE(2, true); else MGS_NODE(1, true); }
  else { if (V == 4) MGS_NODE(4, false); else if (V == 2) MGS_NODE(2, false); else MGS_NODE(1, false); }
#undef MGS_NODE
  return check_launch("gat_bwd_node_kernel");
}

extern "C" size_t mgs_gat_bwd_att_workspace_bytes(int32_t heads, int32_t channels) {
  if (heads <= 0 || channels <= 0) return 0;
  return sizeof(float) * 2 * (size_t)kAttParts * heads * channels;
}

extern "C" int mgs_gat_bwd_att(const float* xh, int64_t ld, int64_t num_nodes, int32_t heads, int32_t channels,
                               const float* da_src, const float* da_dst, float* datt_src, float* datt_dst,
                               void* workspace, size_t workspace_bytes, mgs_stream_t stream_) {
  MGS_GAT_COMMON_CHECKS("mgs_gat_bwd_att");
  const int HC = heads * channels;
  MGS_REQUIRE(ld >= HC, "mgs_gat_bwd_att: leading dimension < heads*channels");
  MGS_REQUIRE(datt_src && datt_dst, "mgs_gat_bwd_att: null output");
  if (workspace_bytes < mgs_gat_bwd_att_workspace_bytes(heads, channels) || !workspace) {
    set_error("mgs_gat_bwd_att: workspace too small");
    return MGS_ERR_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  int parts = (int)((num_nodes + kWarps - 1) / kWarps);
  if (parts > kAttParts) parts = kAttParts;
  if (parts < 1) parts = 1;
  const int V = vec_width(xh, ld, HC);
  const int chunks = HC / V;
  const int iters = iters_for(chunks);
  const size_t smem = sizeof(float) * 2 * (size_t)kWarps * HC;
  if (iters > 0 && smem <= 48 * 1024) {
#define MGS_L(VV, II) gat_bwd_att_partial_row_kernel<VV, II><<<parts, kThreads, smem, stream>>>( \
      xh, ld, (int)num_nodes, heads, channels, chunks, da_src, da_dst, (float*)workspace)
    MGS_DISPATCH_V_ITERS(V, iters, MGS_L);
#undef MGS_L
    if (int rc = check_launch("gat_bwd_att_partial_row_kernel")) return rc;
  } else {
    gat_bwd_att_partial_kernel<<<parts, kThreads, 0, stream>>>(xh, ld, (int)num_nodes, heads, channels, da_src,
                                                               da_dst, (float*)workspace);
    if (int rc = check_launch("gat_bwd_att_partial_kernel")) return rc;
  }
  gat_bwd_att_final_kernel<<<(HC + 31) / 32, kThreads, 0, stream>>>((const float*)workspace, parts,
                                                                                   HC, datt_src, datt_dst);
  return check_launch("gat_bwd_att_final_kernel");
}
