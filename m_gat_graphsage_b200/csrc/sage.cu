// K1 -- SAGEConv mean aggregation, forward and backward.
//
// HBM-bound gather.  One thread owns one VEC-wide feature chunk of one destination row, so a warp
// reads/writes consecutive addresses (coalesced 64/128-bit accesses); the <= 6 neighbour rows of
// an atom are gathered with all loads of a 4-edge group issued before the first add (MLP), and
// are L2 hits after first touch (x of a 4096-molecule batch is 18-180 MB, molecules are local).
// Summation is the oracle's order: left fold from 0.0 over in-edges in ascending edge id, separate
// multiply / add (no FMA contraction), true fp32 division by the in-degree -- bit-exact vs ATen
// scatter_add_ + divide on the CPU.
// Algorithmic bytes / atom: 4F (x) + 4F (out) + 4 (rowptr) + 4*E/N (col)  [F=350: ~2.81 KB].
#include "common.cuh"

namespace mgs {
namespace {

constexpr int kThreads = 256;
constexpr int kGroup = 4;

template <int V, bool WEIGHTED>
__global__ void __launch_bounds__(kThreads)
sage_aggr_fwd_kernel(const float* __restrict__ x, int64_t ldx, int N, int chunks,
                     const int* __restrict__ rowptr, const int* __restrict__ col,
                     const int* __restrict__ perm, const float* __restrict__ ew,
                     float* __restrict__ out, int64_t ldo) {
  const int64_t total = (int64_t)N * chunks;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (int64_t)gridDim.x * kThreads) {
    const int i = (int)(t / chunks);
    const int c = (int)(t - (int64_t)i * chunks) * V;
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
    Vec<V> acc = vzero<V>();
    for (int p = beg; p < end; p += kGroup) {
      int j[kGroup];
      float w[kGroup];
      Vec<V> v[kGroup];
#pragma unroll
      for (int k = 0; k < kGroup; ++k) j[k] = (p + k < end) ? __ldg(col + p + k) : -1;
#pragma unroll
      for (int k = 0; k < kGroup; ++k) {
        if (j[k] >= 0) {
          v[k] = Vec<V>::load(x + (int64_t)j[k] * ldx + c);
          if (WEIGHTED) w[k] = __ldg(ew + __ldg(perm + p + k));
        }
      }
#pragma unroll
      for (int k = 0; k < kGroup; ++k) {
        if (j[k] >= 0) {
#pragma unroll
          for (int u = 0; u < V; ++u) {
            float m = WEIGHTED ? __fmul_rn(v[k].v[u], w[k]) : v[k].v[u];
            acc.v[u] = __fadd_rn(acc.v[u], m);
          }
        }
      }
    }
    const float cnt = (float)max(end - beg, 1);
#pragma unroll
    for (int u = 0; u < V; ++u) acc.v[u] = __fdiv_rn(acc.v[u], cnt);
    acc.store(out + (int64_t)i * ldo + c);
  }
}

template <int V, bool WEIGHTED>
__global__ void __launch_bounds__(kThreads)
sage_aggr_bwd_kernel(const float* __restrict__ g, int64_t ldg, int N, int chunks,
                     const int* __restrict__ rowptr, const int* __restrict__ colptr,
                     const int* __restrict__ row, const int* __restrict__ permt,
                     const float* __restrict__ ew, float* __restrict__ gx, int64_t ldgx) {
  const int64_t total = (int64_t)N * chunks;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (int64_t)gridDim.x * kThreads) {
    const int j = (int)(t / chunks);
    const int c = (int)(t - (int64_t)j * chunks) * V;
    const int beg = __ldg(colptr + j), end = __ldg(colptr + j + 1);
    Vec<V> acc = vzero<V>();
    for (int q = beg; q < end; q += kGroup) {
      int i[kGroup];
      float cnt[kGroup], w[kGroup];
      Vec<V> v[kGroup];
#pragma unroll
      for (int k = 0; k < kGroup; ++k) i[k] = (q + k < end) ? __ldg(row + q + k) : -1;
#pragma unroll
      for (int k = 0; k < kGroup; ++k) {
        if (i[k] >= 0) {
          v[k] = Vec<V>::load(g + (int64_t)i[k] * ldg + c);
          cnt[k] = (float)max(__ldg(rowptr + i[k] + 1) - __ldg(rowptr + i[k]), 1);
          if (WEIGHTED) w[k] = __ldg(ew + __ldg(permt + q + k));
        }
      }
#pragma unroll
      for (int k = 0; k < kGroup; ++k) {
        if (i[k] >= 0) {
#pragma unroll
          for (int u = 0; u < V; ++u) {
            float m = __fdiv_rn(v[k].v[u], cnt[k]);          // d(sum/count) first, as autograd does
            if (WEIGHTED) m = __fmul_rn(m, w[k]);
            acc.v[u] = __fadd_rn(acc.v[u], m);
          }
        }
      }
    }
    acc.store(gx + (int64_t)j * ldgx + c);
  }
}

// ---------------------------------------------------------------------------------------------
// Warp-per-row variants (the fast path; see common.cuh "warp-per-row mapping").  Same arithmetic and
// the same left-fold order as the flat kernels above, which remain the fallback for rows wider than
// 8 x 32 chunks.  Multiplying by an edge weight of exactly 1.0f is exact, so one code path serves the
// weighted (explainer) and unweighted cases bit-identically.
// ---------------------------------------------------------------------------------------------
template <int V, int ITERS> struct GroupOf { static constexpr int value = (48 / (V * ITERS)) >= 4 ? 4 : ((48 / (V * ITERS)) >= 2 ? 2 : 1); };

template <int V, int ITERS>
__global__ void __launch_bounds__(kThreads)
sage_aggr_fwd_row_kernel(const float* __restrict__ x, int64_t ldx, int N, int chunks,
                         const int* __restrict__ rowptr, const int* __restrict__ col,
                         const int* __restrict__ perm, const float* __restrict__ ew,
                         float* __restrict__ out, int64_t ldo) {
  constexpr int G = GroupOf<V, ITERS>::value;
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (kThreads / 32);
  for (int i = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); i < N; i += nwarps) {
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
    Vec<V> acc[ITERS];
#pragma unroll
    for (int t = 0; t < ITERS; ++t) acc[t] = vzero<V>();
    for (int p = beg; p < end; p += G) {
      int jl = -1;
      float wl = 1.f;
      if (lane < G && p + lane < end) {
        jl = __ldg(col + p + lane);
        if (ew != nullptr) wl = __ldg(ew + __ldg(perm + p + lane));
      }
      int j[G];
      float w[G];
      Vec<V> v[G][ITERS];
#pragma unroll
      for (int k = 0; k < G; ++k) {
        j[k] = __shfl_sync(0xffffffffu, jl, k);
        w[k] = __shfl_sync(0xffffffffu, wl, k);
      }
#pragma unroll
      for (int k = 0; k < G; ++k) {
        if (j[k] >= 0) {
          const float* src = x + (int64_t)j[k] * ldx;
#pragma unroll
          for (int t = 0; t < ITERS; ++t) {
            const int c = lane + 32 * t;
            if (c < chunks) v[k][t] = Vec<V>::load(src + c * V);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < G; ++k) {
        if (j[k] >= 0) {
#pragma unroll
          for (int t = 0; t < ITERS; ++t) {
            if (lane + 32 * t < chunks) {
#pragma unroll
              for (int u = 0; u < V; ++u) acc[t].v[u] = __fadd_rn(acc[t].v[u], __fmul_rn(v[k][t].v[u], w[k]));
            }
          }
        }
      }
    }
    const float cnt = (float)max(end - beg, 1);
    float* dst = out + (int64_t)i * ldo;
#pragma unroll
    for (int t = 0; t < ITERS; ++t) {
      const int c = lane + 32 * t;
      if (c < chunks) {
#pragma unroll
        for (int u = 0; u < V; ++u) acc[t].v[u] = __fdiv_rn(acc[t].v[u], cnt);
        acc[t].store(dst + c * V);
      }
    }
  }
}

template <int V, int ITERS>
__global__ void __launch_bounds__(kThreads)
sage_aggr_bwd_row_kernel(const float* __restrict__ g, int64_t ldg, int N, int chunks,
                         const int* __restrict__ rowptr, const int* __restrict__ colptr,
                         const int* __restrict__ row, const int* __restrict__ permt,
                         const float* __restrict__ ew, float* __restrict__ gx, int64_t ldgx) {
  constexpr int G = GroupOf<V, ITERS>::value;
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (kThreads / 32);
  for (int j = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); j < N; j += nwarps) {
    const int beg = __ldg(colptr + j), end = __ldg(colptr + j + 1);
    Vec<V> acc[ITERS];
#pragma unroll
    for (int t = 0; t < ITERS; ++t) acc[t] = vzero<V>();
    for (int q = beg; q < end; q += G) {
      int il = -1;
      float wl = 1.f, cl = 1.f;
      if (lane < G && q + lane < end) {
        il = __ldg(row + q + lane);
        cl = (float)max(__ldg(rowptr + il + 1) - __ldg(rowptr + il), 1);
        if (ew != nullptr) wl = __ldg(ew + __ldg(permt + q + lane));
      }
      int i[G];
      float w[G], cnt[G];
      Vec<V> v[G][ITERS];
#pragma unroll
      for (int k = 0; k < G; ++k) {
        i[k] = __shfl_sync(0xffffffffu, il, k);
        w[k] = __shfl_sync(0xffffffffu, wl, k);
        cnt[k] = __shfl_sync(0xffffffffu, cl, k);
      }
#pragma unroll
      for (int k = 0; k < G; ++k) {
        if (i[k] >= 0) {
          const float* src = g + (int64_t)i[k] * ldg;
#pragma unroll
          for (int t = 0; t < ITERS; ++t) {
            const int c = lane + 32 * t;
            if (c < chunks) v[k][t] = Vec<V>::load(src + c * V);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < G; ++k) {
        if (i[k] >= 0) {
#pragma unroll
          for (int t = 0; t < ITERS; ++t) {
            if (lane + 32 * t < chunks) {
#pragma unroll
              for (int u = 0; u < V; ++u)
                acc[t].v[u] = __fadd_rn(acc[t].v[u], __fmul_rn(__fdiv_rn(v[k][t].v[u], cnt[k]), w[k]));
            }
          }
        }
      }
    }
    float* dst = gx + (int64_t)j * ldgx;
#pragma unroll
    for (int t = 0; t < ITERS; ++t) {
      const int c = lane + 32 * t;
      if (c < chunks) acc[t].store(dst + c * V);
    }
  }
}

// d_edge_weight[e] = < g[i,:] / cnt_i , x[j,:] >; one warp per destination row, lanes over features
__global__ void __launch_bounds__(kThreads)
sage_edge_weight_bwd_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ x, int64_t ldx,
                            int N, int F, const int* __restrict__ rowptr, const int* __restrict__ col,
                            const int* __restrict__ perm, float* __restrict__ dew) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = kThreads / 32;
  for (int i = blockIdx.x * warps_per_block + (threadIdx.x >> 5); i < N; i += gridDim.x * warps_per_block) {
    const int beg = rowptr[i], end = rowptr[i + 1];
    const float cnt = (float)max(end - beg, 1);
    for (int p = beg; p < end; ++p) {
      const int j = col[p];
      float s = 0.f;
      for (int f = lane; f < F; f += 32)
        s += __fdiv_rn(__ldg(g + (int64_t)i * ldg + f), cnt) * __ldg(x + (int64_t)j * ldx + f);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) dew[perm[p]] = s;
    }
  }
}

template <bool BWD, typename... Args>
int dispatch(int V, bool weighted, int grid, cudaStream_t stream, Args... args) {
#define MGS_LAUNCH(VV, WW)                                                                   \
  do {                                                                                       \
    if constexpr (BWD) sage_aggr_bwd_kernel<VV, WW><<<grid, kThreads, 0, stream>>>(args...); \
    else sage_aggr_fwd_kernel<VV, WW><<<grid, kThreads, 0, stream>>>(args...);               \
  } while (0)
  if (V == 4) { if (weighted) MGS_LAUNCH(4, true); else MGS_LAUNCH(4, false); }
  else if (V == 2) { if (weighted) MGS_LAUNCH(2, true); else MGS_LAUNCH(2, false); }
  else { if (weighted) MGS_LAUNCH(1, true); else MGS_LAUNCH(1, false); }
#undef MGS_LAUNCH
  return check_launch(BWD ? "sage_aggr_bwd_kernel" : "sage_aggr_fwd_kernel");
}

}  // namespace
}  // namespace mgs

using namespace mgs;

extern "C" int mgs_sage_aggr_fwd(const float* x, int64_t ldx, int64_t num_nodes, int32_t num_feat,
                                 const int32_t* rowptr, const int32_t* col, const int32_t* perm,
                                 const float* edge_weight, float* out, int64_t ldo, mgs_stream_t stream_) {
  MGS_REQUIRE(num_nodes >= 0 && num_nodes < 0x7fffffff && num_feat > 0, "mgs_sage_aggr_fwd: bad sizes");
  MGS_REQUIRE(ldx >= num_feat && ldo >= num_feat, "mgs_sage_aggr_fwd: leading dimension < num_feat");
  if (num_nodes == 0) return MGS_OK;
  MGS_REQUIRE(x && out && rowptr, "mgs_sage_aggr_fwd: null pointer");  // col may be null when E == 0
  MGS_REQUIRE(!edge_weight || perm, "mgs_sage_aggr_fwd: edge_weight needs perm");
  const int V = min_int(vec_width(x, ldx, num_feat), vec_width(out, ldo, num_feat));
  const int chunks = num_feat / V;
  const int iters = iters_for(chunks);
  if (iters > 0) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const int grid = grid_for(num_nodes * 32, kThreads, 8);
#define MGS_L(VV, II) sage_aggr_fwd_row_kernel<VV, II><<<grid, kThreads, 0, stream>>>( \
      x, ldx, (int)num_nodes, chunks, rowptr, col, perm, edge_weight, out, ldo)
    MGS_DISPATCH_V_ITERS(V, iters, MGS_L);
#undef MGS_L
    return check_launch("sage_aggr_fwd_row_kernel");
  }
  const int grid = grid_for(num_nodes * chunks, kThreads, 8);
  return dispatch<false>(V, edge_weight != nullptr, grid, (cudaStream_t)stream_, x, ldx, (int)num_nodes, chunks,
                         rowptr, col, perm, edge_weight, out, ldo);
}

extern "C" int mgs_sage_aggr_bwd(const float* g, int64_t ldg, int64_t num_nodes, int32_t num_feat,
                                 const int32_t* rowptr, const int32_t* colptr, const int32_t* row,
                                 const int32_t* permt, const float* edge_weight, float* gx, int64_t ldgx,
                                 mgs_stream_t stream_) {
  MGS_REQUIRE(num_nodes >= 0 && num_nodes < 0x7fffffff && num_feat > 0, "mgs_sage_aggr_bwd: bad sizes");
  MGS_REQUIRE(ldg >= num_feat && ldgx >= num_feat, "mgs_sage_aggr_bwd: leading dimension < num_feat");
  if (num_nodes == 0) return MGS_OK;
  MGS_REQUIRE(g && gx && rowptr && colptr, "mgs_sage_aggr_bwd: null pointer");
  MGS_REQUIRE(!edge_weight || permt, "mgs_sage_aggr_bwd: edge_weight needs permt");
  const int V = min_int(vec_width(g, ldg, num_feat), vec_width(gx, ldgx, num_feat));
  const int chunks = num_feat / V;
  const int iters = iters_for(chunks);
  if (iters > 0) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const int grid = grid_for(num_nodes * 32, kThreads, 8);
#define MGS_L(VV, II) sage_aggr_bwd_row_kernel<VV, II><<<grid, kThreads, 0, stream>>>( \
      g, ldg, (int)num_nodes, chunks, rowptr, colptr, row, permt, edge_weight, gx, ldgx)
    MGS_DISPATCH_V_ITERS(V, iters, MGS_L);
#undef MGS_L
    return check_launch("sage_aggr_bwd_row_kernel");
  }
  const int grid = grid_for(num_nodes * chunks, kThreads, 8);
  return dispatch<true>(V, edge_weight != nullptr, grid, (cudaStream_t)stream_, g, ldg, (int)num_nodes, chunks,
                        rowptr, colptr, row, permt, edge_weight, gx, ldgx);
}

extern "C" int mgs_sage_aggr_bwd_edge_weight(const float* g, int64_t ldg, const float* x, int64_t ldx,
                                             int64_t num_nodes, int32_t num_feat, const int32_t* rowptr,
                                             const int32_t* col, const int32_t* perm, float* d_edge_weight,
                                             mgs_stream_t stream_) {
  MGS_REQUIRE(num_nodes >= 0 && num_nodes < 0x7fffffff && num_feat > 0, "mgs_sage_aggr_bwd_edge_weight: bad sizes");
  if (num_nodes == 0) return MGS_OK;
  MGS_REQUIRE(g && x && rowptr, "mgs_sage_aggr_bwd_edge_weight: null pointer");
  const int grid = grid_for(num_nodes * 32, kThreads, 8);
  sage_edge_weight_bwd_kernel<<<grid, kThreads, 0, (cudaStream_t)stream_>>>(
      g, ldg, x, ldx, (int)num_nodes, num_feat, rowptr, col, perm, d_edge_weight);
  return check_launch("sage_edge_weight_bwd_kernel");
}
