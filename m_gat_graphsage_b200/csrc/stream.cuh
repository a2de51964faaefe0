// Block-streamed gather/accumulate skeleton shared by the four aggregation kernels
//   SAGE_FWD      out[i]  = (sum_{e=(j->i)} w_e x[j]) / max(indeg(i), 1)            (K1 forward)
//   SAGE_BWD      gx[j]   =  sum_{e=(j->i)} w_e g[i] / max(indeg(i), 1)             (K1 backward)
//   GAT_FWD       out[i]  =  sum_{slots of i} alpha[slot,h] w_e xh[j] + bias        (K2 aggregate)
//   GAT_BWD_NODE  dxh[j]  =  sum_{out-slots of j} alpha[slot,h] w_e g[i] + da_src[j,h] att_src + da_dst[j,h] att_dst
//
// Why this shape (B200 measurements that led here): a thread-per-chunk kernel ran at 20 % of HBM peak and a
// warp-per-row kernel at 25 %: both spend most of their time in the dependent chain row pointer -> neighbour
// id -> feature row, three global round trips during which almost nothing is in flight.  Here a warp owns a
// BLOCK of 32 consecutive output rows:
//   1. lane l reads the pointer pair of row l (one coalesced load), a warp scan turns the row lengths into
//      positions in one flat entry stream for the block;
//   2. every lane writes its row's entries (source row id, edge weight, slot / count) to a per-warp shared
//      memory list (<= 256 entries per window; longer blocks are processed in windows, so any degree works);
//   3. the warp then streams the list G = 4 entries at a time: all G x ITERS vector loads of a group are
//      issued before the first accumulate (lane l owns chunks l, l+32, ... of a row: every load is one
//      coalesced 256/512-byte access), rows are finalised and stored in order as the stream passes their end.
// Metadata latency is paid once per 32 rows instead of once per row, and G rows x 1.4 KB are in flight per
// warp at all times.  Arithmetic and summation order are exactly those of the row kernels (left fold in
// ascending edge id, separate multiply / add, true division), so SAGE results stay bit-exact vs the oracle.
#pragma once

#include "common.cuh"

namespace mgs {
namespace stream {

constexpr int kRows = 32;
constexpr int kListMax = 256;
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

enum Mode { SAGE_FWD = 0, SAGE_BWD = 1, GAT_FWD = 2, GAT_BWD_NODE = 3 };

struct Args {
  const float* src;      // gathered matrix: x (SAGE_FWD), g (SAGE_BWD, GAT_BWD_NODE), xh (GAT_FWD)
  int64_t lds;
  float* dst;
  int64_t ldd;
  int N, chunks, H, C;
  const int* ptr;        // rowptr (by destination) for *_FWD, colptr (by source) for *_BWD*
  const int* idx;        // col for *_FWD, row for *_BWD*
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
};

template <int V, int ITERS> struct GroupOf {
  static constexpr int value = (48 / (V * ITERS)) >= 4 ? 4 : ((48 / (V * ITERS)) >= 2 ? 2 : 1);
};

template <int MODE, int V, int ITERS>
__global__ void __launch_bounds__(kThreads, 2) stream_kernel(const Args a) {
  constexpr bool kGat = MODE == GAT_FWD || MODE == GAT_BWD_NODE;
  constexpr int G = GroupOf<V, ITERS>::value;
  __shared__ int s_src[kWarps][kListMax];
  __shared__ float s_w[kWarps][kListMax];
  __shared__ int s_aux[kWarps][kListMax];   // SAGE_BWD: in-degree of the target; GAT: slot

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = gridDim.x * kWarps;
  const int nblocks = (a.N + kRows - 1) / kRows;
  const bool alpha_by_shuffle = a.H <= 32;

  int hh[ITERS][V];   // head of every element this lane owns (row independent)
  if (kGat) {
#pragma unroll
    for (int t = 0; t < ITERS; ++t)
#pragma unroll
      for (int u = 0; u < V; ++u) hh[t][u] = min(((lane + 32 * t) * V + u) / a.C, a.H - 1);
  }

  // Row blocks are visited in DESCENDING order.  The tensors of a 4096-molecule batch (183 MB) exceed the
  // 126 MB L2, and the neighbouring kernels (PyTorch's element-wise ReLU, the projection GEMM) sweep rows in
  // ascending order: reading the producer's most recently written rows first finds them still in L2, and
  // writing the low rows last leaves them in L2 for the consumer.  Measured on B200 (SAGE aggregate, F = 350):
  // 0.10 ms with a cold L2, 0.20 ms right after an ascending ReLU when sweeping ascending as well.
  for (int bb = blockIdx.x * kWarps + warp; bb < nblocks; bb += nwarps) {
    const int blk = nblocks - 1 - bb;
    const int i0 = blk * kRows;
    const int nrows = min(kRows, a.N - i0);
    const int my = i0 + lane;
    int beg = 0, len = 0, elen = 0;
    if (lane < nrows) {
      beg = __ldg(a.ptr + my);
      elen = __ldg(a.ptr + my + 1) - beg;
      len = elen + (kGat ? 1 : 0);
    }
    int ve = len;   // inclusive scan of the row lengths: virtual end of my row in the block's entry stream
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, ve, o);
      if (lane >= o) ve += t;
    }
    const int vs = ve - len;
    const int total = __shfl_sync(0xffffffffu, ve, 31);

    Vec<V> acc[ITERS];
#pragma unroll
    for (int t = 0; t < ITERS; ++t) acc[t] = vzero<V>();
    int r = 0;                                              // row currently being accumulated
    int vend_r = __shfl_sync(0xffffffffu, ve, 0);

    auto finalize = [&](int row) {
      const int i = i0 + row;
      float* dst = a.dst + (int64_t)i * a.ldd;
      float cnt = 1.f;
      if (MODE == SAGE_FWD) cnt = (float)max(__shfl_sync(0xffffffffu, elen, row), 1);
#pragma unroll
      for (int t = 0; t < ITERS; ++t) {
        const int c = lane + 32 * t;
        if (c < a.chunks) {
#pragma unroll
          for (int u = 0; u < V; ++u) {
            float val = acc[t].v[u];
            if (MODE == SAGE_FWD) val = __fdiv_rn(val, cnt);
            if (MODE == GAT_FWD && a.bias != nullptr) val = __fadd_rn(val, __ldg(a.bias + c * V + u));
            if (MODE == GAT_BWD_NODE) {
              const int f = c * V + u;
              val = fmaf(__ldg(a.da_src + (int64_t)i * a.H + hh[t][u]), __ldg(a.att_src + f), val);
              val = fmaf(__ldg(a.da_dst + (int64_t)i * a.H + hh[t][u]), __ldg(a.att_dst + f), val);
            }
            acc[t].v[u] = val;
          }
          acc[t].store(dst + c * V);
          acc[t] = vzero<V>();
        }
      }
    };

    for (int w0 = 0; w0 < total; w0 += kListMax) {
      const int w1 = min(total, w0 + kListMax);
      __syncwarp();
      // ---- build this window of the entry list: every lane writes the entries of its own row ----
      for (int t = max(vs, w0); t < min(ve, w1); ++t) {
        const int k = t - vs;
        int j, aux = 0;
        float w = 1.f;
        if (kGat && k == len - 1) {                         // the self loop PyG appends last
          j = my;
          aux = (MODE == GAT_FWD) ? beg + elen + my : __ldg(a.rowptr + my + 1) + my;
        } else {
          const int p = beg + k;
          j = __ldg(a.idx + p);
          if (a.ew != nullptr) w = __ldg(a.ew + __ldg(a.eid + p));
          if (MODE == SAGE_BWD) aux = max(__ldg(a.rowptr + j + 1) - __ldg(a.rowptr + j), 1);
          if (MODE == GAT_FWD) aux = p + my;
          if (MODE == GAT_BWD_NODE) aux = __ldg(a.csc_pos + p) + j;
          if (kGat && j == my) j = -1;                      // pre-existing self loop: removed by GATConv
        }
        s_src[warp][t - w0] = j;
        s_w[warp][t - w0] = w;
        s_aux[warp][t - w0] = aux;
      }
      __syncwarp();
      // ---- stream the window ----
      const int wl = w1 - w0;
      for (int t = 0; t < wl; t += G) {
        int j[G], aux[G];
        float w[G], al[G];
        Vec<V> v[G][ITERS];
#pragma unroll
        for (int k = 0; k < G; ++k) {
          j[k] = -1;
          if (t + k < wl) {
            j[k] = s_src[warp][t + k];
            w[k] = s_w[warp][t + k];
            aux[k] = s_aux[warp][t + k];
          }
        }
#pragma unroll
        for (int k = 0; k < G; ++k) {
          if (j[k] >= 0) {
            const float* srow = a.src + (int64_t)j[k] * a.lds;
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
              const int c = lane + 32 * it;
              if (c < a.chunks) v[k][it] = Vec<V>::load(srow + c * V);
            }
            if (kGat && alpha_by_shuffle) al[k] = lane < a.H ? __ldg(a.alpha + (int64_t)aux[k] * a.H + lane) : 0.f;
          }
        }
#pragma unroll
        for (int k = 0; k < G; ++k) {
          if (t + k < wl) {
            const int vt = w0 + t + k;
            while (vt >= vend_r) {                          // the stream passed the end of row r
              finalize(r);
              ++r;
              vend_r = __shfl_sync(0xffffffffu, ve, min(r, 31));
            }
            if (j[k] >= 0) {
              float cntk = 1.f;
              if (MODE == SAGE_BWD) cntk = (float)aux[k];
#pragma unroll
              for (int it = 0; it < ITERS; ++it) {
                const bool active = lane + 32 * it < a.chunks;
#pragma unroll
                for (int u = 0; u < V; ++u) {
                  float av = 0.f;
                  if (kGat) {
                    // every lane takes part in the shuffle (lanes past the row end hold a clamped head id)
                    if (alpha_by_shuffle) av = __shfl_sync(0xffffffffu, al[k], hh[it][u]);
                    else if (active) av = __ldg(a.alpha + (int64_t)aux[k] * a.H + hh[it][u]);
                  }
                  if (active) {
                    float m;
                    if (MODE == SAGE_FWD) m = __fmul_rn(v[k][it].v[u], w[k]);
                    else if (MODE == SAGE_BWD) m = __fmul_rn(__fdiv_rn(v[k][it].v[u], cntk), w[k]);
                    else if (MODE == GAT_FWD) m = __fmul_rn(__fmul_rn(av, v[k][it].v[u]), w[k]);
                    else m = __fmul_rn(__fmul_rn(av, w[k]), v[k][it].v[u]);
                    acc[it].v[u] = __fadd_rn(acc[it].v[u], m);
                  }
                }
              }
            }
          }
        }
      }
    }
    while (r < nrows) {                                     // rows after the last entry (e.g. isolated atoms)
      finalize(r);
      ++r;
    }
  }
}

template <int MODE>
inline int launch(const Args& a, int V, int iters, cudaStream_t stream, const char* what) {
  const int nblocks = (a.N + kRows - 1) / kRows;
  const int grid = grid_for((int64_t)nblocks * 32, kThreads, 2);
#define MGS_L(VV, II) stream_kernel<MODE, VV, II><<<grid, kThreads, 0, stream>>>(a)
  MGS_DISPATCH_V_ITERS(V, iters, MGS_L);
#undef MGS_L
  return check_launch(what);
}

}  // namespace stream
}  // namespace mgs
