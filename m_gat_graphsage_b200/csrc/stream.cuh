// Block-streamed gather/accumulate skeleton shared by the aggregation kernels
//   SAGE_FWD      out[i]  = (sum_{e=(j->i)} w_e x[j]) / max(indeg(i), 1)            (K1 forward)
//   SAGE_BWD      gx[j]   =  sum_{e=(j->i)} w_e g[i] / max(indeg(i), 1)             (K1 backward)
//   GAT_FWD       out[i]  =  sum_{slots of i} alpha[slot,h] w_e xh[j] + bias        (K2 aggregate)
//   GAT_BWD_NODE  dxh[j]  =  sum_{out-slots of j} alpha[slot,h] w_e g[i] + da_src[j,h] att_src + da_dst[j,h] att_dst
//   SUM           dst[i] (+)= sum_{entries of i} w_e src[j]                          (fused SAGEConv backward: the
//                 1 / indeg scaling already applied by the producing GEMM's epilogue)
//
// Shape: a warp owns a contiguous range of output rows and walks it in blocks of 32:
//   1. lane l reads the pointer pair of row l (one coalesced load), a warp scan turns the row lengths into
//      positions in one flat entry stream for the block (rows without entries get one dummy entry, so every
//      row ends with an entry carrying the "last" flag);
//   2. every lane writes its row's entries (source row id | last flag, edge weight, slot / count) to a
//      per-warp shared-memory list (<= 256 entries per window; longer blocks are processed in windows, so any
//      degree works), padded with no-op entries to a multiple of G;
//   3. the warp streams the list G = 4 entries at a time: all G x ITERS vector loads of a group are issued
//      before the first accumulate (lane l owns chunks l, l+32, ... of a row: every load is one coalesced
//      256/512-byte access); an entry with the "last" flag finalises and stores its row.
// Arithmetic and summation order of the SAGE modes are exactly the oracle's (left fold in ascending edge id,
// separate multiply / add, true division), so those results stay bit-exact.
//
// B200 history (model1 batch, F = 350, 130 k rows): thread-per-chunk 20 % of HBM peak, warp-per-row 25 %,
// first block-stream version 28 % on real activations.  ncu on that version: 41 M (SAGE) to 111 M (GAT) warp
// instructions per launch with 16 warps / SM -> issue-bound at 44 % issue utilisation, DRAM 47 % busy; 155 to
// 245 instructions per gathered row where ~20 are loads and adds.  What this version removes:
//   * IEEE division (`__fdiv_rn`): ~10 slots per element and a ~100-instruction slow path for zero dividends
//     -- half of a post-ReLU activation matrix; 2x (fwd) / 2.7x (bwd) slower on real data than on random
//     rows.  Now: one reciprocal per row / entry + 3 FMAs per element, range check per row (common.cuh);
//   * per-element predicates, branches and 64-bit address arithmetic (5 instructions per load): addresses are
//     (row pointer + lane offset) + immediates, only the last two iterations are predicated; row ends are a
//     flag in the entry, not a scan of the row table; alpha travels by shuffle only (H <= 32; wider layers
//     use the row kernels);
//   * `* w_e` when there is no explainer edge mask (separate instantiation; x * 1.0f == x, bit-identical);
//   * scalar FADD/FFMA: packed `add.rn.f32x2` / `fma.rn.f32x2` (FADD2 / FFMA2), IEEE per component;
//   * round-robin blocks (4083 blocks over 2368 warps idled 14 % of the machine): even contiguous split.
#pragma once

#include "common.cuh"

namespace mgs {
namespace stream {

constexpr int kRows = 32;
constexpr int kListMax = 256;
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

enum Mode { SAGE_FWD = 0, SAGE_BWD = 1, GAT_FWD = 2, GAT_BWD_NODE = 3, SUM = 4 };

struct Args {
  const float* src;      // gathered matrix: x (SAGE_FWD), g (SAGE_BWD, GAT_BWD_NODE, SUM), xh (GAT_FWD)
  int64_t lds;
  float* dst;
  int64_t ldd;
  int N, chunks, H, C;
  const int* ptr;        // rowptr (by destination) for *_FWD, colptr (by source) for *_BWD* / SUM
  const int* idx;        // col for *_FWD, row for *_BWD* / SUM
  const int* eid;        // perm / permt: original edge id of an entry (edge weights)
  const float* ew;       // optional edge weights by original edge id
  const int* rowptr;     // by-destination pointers (SAGE_BWD in-degree, GAT_BWD_NODE self slot)
  const int* csc_pos;    // GAT_BWD_NODE: position of an entry in the by-destination order
  const float* alpha;    // GAT: [(E+N), H] in slot order
  const float* bias;     // GAT_FWD (optional)
  const float* da_src;   // GAT_BWD_NODE [N, H]
  const float* da_dst;
  const float* att_src;  // GAT_BWD_NODE [H*C]
  const float* att_dst;
  int accumulate;        // SUM / SAGE_BWD: dst = base + sum (base == dst: in place) instead of dst = sum
  const float* base;
  int64_t ldb;
  int activation;        // GAT_FWD epilogue: 0 none, 1 ReLU, 2 ELU(alpha = 1)  (the reference applies it right after the
                         // layer: ablation/model1.py:68-69, gnn/gat.py:63)
  const float* mask;     // SAGE_BWD / SUM epilogue: dst = mask <= 0 ? 0 : dst -- the backward of the ReLU that PRODUCED this
  int64_t ldm;           // layer's input (mask = that ReLU's output), fused into the gradient's producer
  // The same mask as ONE BIT per element (V * ITERS words per row; bit l of word t * V + u = column (l + 32 t) V + u, i.e.
  // the lane mapping of these kernels): written by the GAT_FWD ReLU epilogue, read by the SAGE_BWD epilogue -- 48 bytes
  // per row instead of a second pass over the 1400-byte activation row (the float mask cost 0.087 ms per step).
  unsigned* bits_out;
  const unsigned* bits_in;
};

// ---- V floats of one lane: packed pairs so that adds / FMAs are FADD2 / FFMA2 -------------------------
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 p, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

template <int V> struct Row;
template <> struct Row<1> {
  float x;
  __device__ __forceinline__ static Row load(const float* p) { Row r; r.x = __ldg(p); return r; }
  __device__ __forceinline__ static Row load_rw(const float* p) { Row r; r.x = *p; return r; }
  __device__ __forceinline__ static Row zero() { Row r; r.x = 0.f; return r; }
  __device__ __forceinline__ void add(const Row& o) { x = __fadd_rn(x, o.x); }
  __device__ __forceinline__ void fma(const Row& a, const Row& v) { x = fmaf(a.x, v.x, x); }
  __device__ __forceinline__ float get(int) const { return x; }
  __device__ __forceinline__ void set(int, float f) { x = f; }
  __device__ __forceinline__ void store(float* p) const { *p = x; }
};
template <> struct Row<2> {
  u64 p;
  __device__ __forceinline__ static Row load(const float* q) {
    Row r; r.p = __ldg(reinterpret_cast<const u64*>(q)); return r;
  }
  __device__ __forceinline__ static Row load_rw(const float* q) {
    Row r; r.p = *reinterpret_cast<const u64*>(q); return r;
  }
  __device__ __forceinline__ static Row zero() { Row r; r.p = 0ull; return r; }
  __device__ __forceinline__ void add(const Row& o) { p = add2(p, o.p); }
  __device__ __forceinline__ void fma(const Row& a, const Row& v) { p = fma2(a.p, v.p, p); }
  __device__ __forceinline__ float get(int u) const { float lo, hi; unpack2(p, lo, hi); return u ? hi : lo; }
  __device__ __forceinline__ void set(int u, float f) { float lo, hi; unpack2(p, lo, hi); p = u ? pack2(lo, f) : pack2(f, hi); }
  __device__ __forceinline__ void store(float* q) const { *reinterpret_cast<u64*>(q) = p; }
};
template <> struct Row<4> {
  u64 p[2];
  __device__ __forceinline__ static Row load(const float* q) {
    const ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2*>(q));
    Row r; r.p[0] = t.x; r.p[1] = t.y; return r;
  }
  __device__ __forceinline__ static Row load_rw(const float* q) {
    const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(q);
    Row r; r.p[0] = t.x; r.p[1] = t.y; return r;
  }
  __device__ __forceinline__ static Row zero() { Row r; r.p[0] = 0ull; r.p[1] = 0ull; return r; }
  __device__ __forceinline__ void add(const Row& o) { p[0] = add2(p[0], o.p[0]); p[1] = add2(p[1], o.p[1]); }
  __device__ __forceinline__ void fma(const Row& a, const Row& v) {
    p[0] = fma2(a.p[0], v.p[0], p[0]); p[1] = fma2(a.p[1], v.p[1], p[1]);
  }
  __device__ __forceinline__ float get(int u) const { float lo, hi; unpack2(p[u >> 1], lo, hi); return (u & 1) ? hi : lo; }
  __device__ __forceinline__ void set(int u, float f) {
    float lo, hi; unpack2(p[u >> 1], lo, hi);
    p[u >> 1] = (u & 1) ? pack2(lo, f) : pack2(f, hi);
  }
  __device__ __forceinline__ void store(float* q) const {
    *reinterpret_cast<ulonglong2*>(q) = make_ulonglong2(p[0], p[1]);
  }
};

// Divide the V floats of each row in `r` by the count `cnt` (rc = __frcp_rn(cnt)) with the bits of IEEE
// division.  Branch-free 3-FMA quotients + sign-of-zero fix; one range test for the whole vector decides
// whether the (rare) lanes holding infinities or dividends below 2^-100 redo theirs with __fdiv_rn
// (see div_by_count in common.cuh for why this is exact).
template <int V, int NR>
__device__ __forceinline__ void divide_rows(Row<V> (&r)[NR], float cnt, float rc) {
  float amax = 0.f;
  unsigned kmin = 0xffffffffu;
  float q1[NR][V];
#pragma unroll
  for (int t = 0; t < NR; ++t)
#pragma unroll
    for (int u = 0; u < V; ++u) {
      const float a = r[t].get(u);
      const float q = __fmul_rn(a, rc);
      const float res = __fmaf_rn(-cnt, q, a);
      const float q2 = __fmaf_rn(res, rc, q);
      q1[t][u] = __uint_as_float(__float_as_uint(q2) | (__float_as_uint(a) & 0x80000000u));
      amax = fmaxf(amax, fabsf(a));
      kmin = min(kmin, (__float_as_uint(a) & 0x7fffffffu) - 1u);   // zero -> 0xffffffff: never the minimum
    }
  if (amax > 0x1p100f || kmin < 0x0d800000u - 1u) {                // inf / huge, or 0 < |a| < 2^-100
#pragma unroll
    for (int t = 0; t < NR; ++t)
#pragma unroll
      for (int u = 0; u < V; ++u) q1[t][u] = __fdiv_rn(r[t].get(u), cnt);
  }
#pragma unroll
  for (int t = 0; t < NR; ++t)
#pragma unroll
    for (int u = 0; u < V; ++u) r[t].set(u, q1[t][u]);
}

template <int V, int ITERS> struct GroupOf {
  static constexpr int value = (48 / (V * ITERS)) >= 4 ? 4 : ((48 / (V * ITERS)) >= 2 ? 2 : 1);
};

template <int MODE, int V, int ITERS, bool WEIGHTED>
__global__ void __launch_bounds__(kThreads, 2) stream_kernel(const Args a) {
  constexpr bool kGat = MODE == GAT_FWD || MODE == GAT_BWD_NODE;
  constexpr int G = GroupOf<V, ITERS>::value;
  static_assert(kListMax % G == 0, "windows are padded to a multiple of G");
  __shared__ int s_code[kWarps][kListMax];  // (source row + 1) | last-entry-of-its-row << 31; low bits 0 = no gather
  __shared__ float s_w[kWarps][kListMax];   // edge weight (WEIGHTED only)
  __shared__ int s_aux[kWarps][kListMax];   // GAT: slot
  __shared__ float s_cnt[kWarps][kListMax]; // SAGE_BWD: in-degree of the entry's target

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = gridDim.x * kWarps;
  const int rows_per_warp = (a.N + nwarps - 1) / nwarps;
  const int row_lo = (blockIdx.x * kWarps + warp) * rows_per_warp;
  const int row_hi = min(a.N, row_lo + rows_per_warp);

  // Per-lane constants.  Lane l owns chunks l, l + 32, ...: addresses are (row pointer + l * V) + a compile-time
  // immediate per iteration.  ITERS comes from iters_for(): only the last two iterations can be partial, and
  // those are predicated by two per-thread flags (no per-load compare, nothing read past a row end).
  const int lane_off = lane * V;
  const bool tail_a = lane + 32 * (ITERS - 2) < a.chunks;
  const bool tail_b = lane + 32 * (ITERS - 1) < a.chunks;
  auto act = [&](int t) { return t < ITERS - 2 ? true : (t == ITERS - 2 ? tail_a : tail_b); };
  int hh[ITERS][V];
  if (kGat) {
#pragma unroll
    for (int t = 0; t < ITERS; ++t)
#pragma unroll
      for (int u = 0; u < V; ++u) hh[t][u] = min(((lane + 32 * t) * V + u) / a.C, a.H - 1);
  }

  for (int i0 = row_lo; i0 < row_hi; i0 += kRows) {
    const int nrows = min(kRows, row_hi - i0);
    const int my = i0 + lane;
    int beg = 0, len = 0, elen = 0;
    if (lane < nrows) {
      beg = __ldg(a.ptr + my);
      elen = __ldg(a.ptr + my + 1) - beg;
      len = kGat ? elen + 1 : max(elen, 1);
    }
    int ve = len;   // inclusive scan of the row lengths: virtual end of my row in the block's entry stream
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, ve, o);
      if (lane >= o) ve += t;
    }
    const int vs = ve - len;
    const int total = __shfl_sync(0xffffffffu, ve, 31);

    Row<V> acc[ITERS];
#pragma unroll
    for (int t = 0; t < ITERS; ++t) acc[t] = Row<V>::zero();
    int r = 0;                                              // row currently being accumulated

    auto finalize = [&](int row) {
      const int i = i0 + row;
      float* dst = a.dst + (int64_t)i * a.ldd + lane_off;
      if (MODE == SAGE_FWD) {
        const float cnt = (float)max(__shfl_sync(0xffffffffu, elen, row), 1);
        divide_rows<V, ITERS>(acc, cnt, __frcp_rn(cnt));
      }
      if (MODE == GAT_FWD && a.bias != nullptr) {
#pragma unroll
        for (int t = 0; t < ITERS; ++t)
          if (act(t)) acc[t].add(Row<V>::load(a.bias + lane_off + 32 * V * t));
      }
      if (MODE == GAT_BWD_NODE && a.att_src != nullptr) {
        float das = 0.f, dad = 0.f;
        if (lane < a.H) {                                   // one coalesced load, then shuffles
          das = __ldg(a.da_src + (int64_t)i * a.H + lane);
          dad = __ldg(a.da_dst + (int64_t)i * a.H + lane);
        }
#pragma unroll
        for (int t = 0; t < ITERS; ++t) {
          Row<V> s, d;
#pragma unroll
          for (int u = 0; u < V; ++u) {
            s.set(u, __shfl_sync(0xffffffffu, das, hh[t][u]));
            d.set(u, __shfl_sync(0xffffffffu, dad, hh[t][u]));
          }
          if (act(t)) {
            acc[t].fma(s, Row<V>::load(a.att_src + lane_off + 32 * V * t));
            acc[t].fma(d, Row<V>::load(a.att_dst + lane_off + 32 * V * t));
          }
        }
      }
      if ((MODE == SUM || MODE == SAGE_BWD) && a.accumulate) {
        const float* bp = a.base + (int64_t)i * a.ldb + lane_off;
#pragma unroll
        for (int t = 0; t < ITERS; ++t)
          if (act(t)) acc[t].add(Row<V>::load_rw(bp + 32 * V * t));
      }
      if (MODE == GAT_FWD && a.activation != 0) {             // same comparisons as ATen's threshold / elu kernels
#pragma unroll
        for (int t = 0; t < ITERS; ++t)
#pragma unroll
          for (int u = 0; u < V; ++u) {
            const float v = acc[t].get(u);
            acc[t].set(u, v <= 0.f ? (a.activation == 1 ? 0.f : expm1f(v)) : v);
          }
      }
      if (MODE == GAT_FWD && a.bits_out != nullptr) {
        unsigned mine = 0u;
#pragma unroll
        for (int t = 0; t < ITERS; ++t)
#pragma unroll
          for (int u = 0; u < V; ++u) {
            const unsigned b = __ballot_sync(0xffffffffu, act(t) && acc[t].get(u) > 0.f);
            if (lane == t * V + u) mine = b;
          }
        if (lane < V * ITERS) a.bits_out[(int64_t)i * (V * ITERS) + lane] = mine;
      }
      if ((MODE == SUM || MODE == SAGE_BWD) && a.bits_in != nullptr) {
        static_assert(V * ITERS <= 32, "one mask word per lane");
        const unsigned mine = lane < V * ITERS ? __ldg(a.bits_in + (int64_t)i * (V * ITERS) + lane) : 0u;
#pragma unroll
        for (int t = 0; t < ITERS; ++t)
#pragma unroll
          for (int u = 0; u < V; ++u) {
            const unsigned b = __shfl_sync(0xffffffffu, mine, t * V + u);
            if (!((b >> lane) & 1u)) acc[t].set(u, 0.f);
          }
      }
      if ((MODE == SUM || MODE == SAGE_BWD) && a.mask != nullptr) {
        const float* mp = a.mask + (int64_t)i * a.ldm + lane_off;
#pragma unroll
        for (int t = 0; t < ITERS; ++t)
          if (act(t)) {
            const Row<V> m = Row<V>::load(mp + 32 * V * t);
#pragma unroll
            for (int u = 0; u < V; ++u)
              if (m.get(u) <= 0.f) acc[t].set(u, 0.f);
          }
      }
#pragma unroll
      for (int t = 0; t < ITERS; ++t) {
        if (act(t)) acc[t].store(dst + 32 * V * t);
        acc[t] = Row<V>::zero();
      }
    };

    for (int w0 = 0; w0 < total; w0 += kListMax) {
      const int w1 = min(total, w0 + kListMax);
      const int wl = w1 - w0;
      const int wlp = (wl + G - 1) / G * G;
      __syncwarp();
      // ---- build this window of the entry list: every lane writes the entries of its own row ----
      for (int t = max(vs, w0); t < min(ve, w1); ++t) {
        const int k = t - vs;
        int j = -1, aux = 0;
        float w = 1.f;
        if (kGat && k == len - 1) {                         // the self loop PyG appends last
          j = my;
          aux = (MODE == GAT_FWD) ? beg + elen + my : __ldg(a.rowptr + my + 1) + my;
        } else if (k < elen) {                              // (else: the dummy entry of an empty row)
          const int p = beg + k;
          j = __ldg(a.idx + p);
          if (WEIGHTED) w = __ldg(a.ew + __ldg(a.eid + p));
          if (MODE == SAGE_BWD)
            s_cnt[warp][t - w0] = (float)max(__ldg(a.rowptr + j + 1) - __ldg(a.rowptr + j), 1);
          if (MODE == GAT_FWD) aux = p + my;
          if (MODE == GAT_BWD_NODE) aux = __ldg(a.csc_pos + p) + j;
          if (kGat && j == my) j = -1;                      // pre-existing self loop: removed by GATConv
        }
        s_code[warp][t - w0] = (j + 1) | (k == len - 1 ? (int)0x80000000u : 0);
        if (WEIGHTED) s_w[warp][t - w0] = w;
        if (kGat) s_aux[warp][t - w0] = aux;
      }
      if (lane < wlp - wl) s_code[warp][wl + lane] = 0;     // padding: no gather, not a row end
      __syncwarp();
      // ---- stream the window ----
      for (int t = 0; t < wlp; t += G) {
        int code[G];
        bool has[G];
        float w[G], al[G], cn[G];
        Row<V> v[G][ITERS];
#pragma unroll
        for (int k = 0; k < G; ++k) {
          code[k] = s_code[warp][t + k];
          w[k] = 1.f; al[k] = 0.f; cn[k] = 1.f;
          const int j = (code[k] & 0x7fffffff) - 1;
          has[k] = j >= 0;
          if (has[k]) {
            const float* srow = a.src + (int64_t)j * a.lds + lane_off;
#pragma unroll
            for (int it = 0; it < ITERS; ++it)
              if (act(it)) v[k][it] = Row<V>::load(srow + 32 * V * it);
            if (kGat && lane < a.H) al[k] = __ldg(a.alpha + (int64_t)s_aux[warp][t + k] * a.H + lane);
            if (WEIGHTED) w[k] = s_w[warp][t + k];
            if (MODE == SAGE_BWD) cn[k] = s_cnt[warp][t + k];
          }
        }
#pragma unroll
        for (int k = 0; k < G; ++k) {
          if (has[k]) {
            if (MODE == SAGE_BWD) {
#pragma unroll
              for (int it = ITERS - 2 < 0 ? 0 : ITERS - 2; it < ITERS; ++it)
                if (!act(it)) v[k][it] = Row<V>::zero();     // the divider looks at every component
              divide_rows<V, ITERS>(v[k], cn[k], __frcp_rn(cn[k]));
            }
            if (kGat && WEIGHTED) al[k] = __fmul_rn(al[k], w[k]);
          }
#pragma unroll
          for (int it = 0; it < ITERS; ++it) {
            if (kGat) {
              Row<V> av;                                    // every lane takes part in the shuffles
#pragma unroll
              for (int u = 0; u < V; ++u) av.set(u, __shfl_sync(0xffffffffu, al[k], hh[it][u]));
              if (has[k] && act(it)) acc[it].fma(av, v[k][it]);
            } else if (has[k] && act(it)) {
              if (WEIGHTED) {
#pragma unroll
                for (int u = 0; u < V; ++u) v[k][it].set(u, __fmul_rn(v[k][it].get(u), w[k]));
              }
              acc[it].add(v[k][it]);
            }
          }
          if (code[k] < 0) {                                // last entry of row r
            finalize(r);
            ++r;
          }
        }
      }
    }
  }
}

template <int MODE>
inline int launch(const Args& a, int V, int iters, cudaStream_t stream, const char* what) {
  // one warp per <= kRows rows when the batch is small, else two resident CTAs per SM and an even split
  const int grid = grid_for_rows(a.N, kWarps, 2);
  if (a.ew != nullptr) {
#define MGS_L(VV, II) stream_kernel<MODE, VV, II, true><<<grid, kThreads, 0, stream>>>(a)
    MGS_DISPATCH_V_ITERS(V, iters, MGS_L);
#undef MGS_L
  } else {
#define MGS_L(VV, II) stream_kernel<MODE, VV, II, false><<<grid, kThreads, 0, stream>>>(a)
    MGS_DISPATCH_V_ITERS(V, iters, MGS_L);
#undef MGS_L
  }
  return check_launch(what);
}

}  // namespace stream
}  // namespace mgs
