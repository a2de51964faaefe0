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
// The production path is the block-streamed kernel of stream.cuh (same arithmetic); the flat
// thread-per-chunk kernels in this file are the fallback for rows wider than 8 x 32 vector chunks.
#include "common.cuh"
#include "stream.cuh"

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

__global__ void __launch_bounds__(kThreads)
selftest_div_kernel(int count_lo, int count_hi, unsigned long long stride, unsigned long long* mismatches) {
  unsigned long long bad = 0;
  for (int c = count_lo; c <= count_hi; ++c) {
    const float b = (float)c;
    const float rc = __frcp_rn(b);
    for (unsigned long long bits = ((unsigned long long)blockIdx.x * kThreads + threadIdx.x) * stride;
         bits < 0x100000000ull; bits += (unsigned long long)gridDim.x * kThreads * stride) {
      const float a = __uint_as_float((unsigned)bits);
      const float want = __fdiv_rn(a, b);
      const float got = div_by_count(a, b, rc);
      if (__float_as_uint(want) != __float_as_uint(got) && !(want != want && got != got)) ++bad;
    }
  }
  if (bad) atomicAdd(mismatches, bad);
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

extern "C" int mgs_selftest_div(int32_t count_lo, int32_t count_hi, uint64_t stride, unsigned long long* mismatches,
                                mgs_stream_t stream_) {
  MGS_REQUIRE(count_lo >= 1 && count_hi >= count_lo && stride >= 1 && mismatches, "mgs_selftest_div: bad arguments");
  MGS_CUDA(cudaMemsetAsync(mismatches, 0, sizeof(unsigned long long), (cudaStream_t)stream_));
  selftest_div_kernel<<<sm_count() * 8, kThreads, 0, (cudaStream_t)stream_>>>(count_lo, count_hi, stride, mismatches);
  return check_launch("selftest_div_kernel");
}

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
  if (iters > 0) {   // block-streamed fast path (stream.cuh); the flat kernel below handles very wide rows
    stream::Args sa = {};
    sa.src = x; sa.lds = ldx; sa.dst = out; sa.ldd = ldo;
    sa.N = (int)num_nodes; sa.chunks = chunks; sa.H = 1; sa.C = num_feat;
    sa.ptr = rowptr; sa.idx = col; sa.eid = perm; sa.ew = edge_weight;
    return stream::launch<stream::SAGE_FWD>(sa, V, iters, (cudaStream_t)stream_, "sage_aggr_fwd(stream)");
  }
  const int grid = grid_for(num_nodes * chunks, kThreads, 8);
  return dispatch<false>(V, edge_weight != nullptr, grid, (cudaStream_t)stream_, x, ldx, (int)num_nodes, chunks,
                         rowptr, col, perm, edge_weight, out, ldo);
}

static int sage_aggr_bwd_impl(const uint32_t* bits, int bits_words, const float* mask, int64_t ldmask, const float* base, int64_t ldbase, const float* g, int64_t ldg, int64_t num_nodes, int32_t num_feat,
                                 const int32_t* rowptr, const int32_t* colptr, const int32_t* row,
                                 const int32_t* permt, const float* edge_weight, float* gx, int64_t ldgx,
                                 mgs_stream_t stream_) {
  MGS_REQUIRE(num_nodes >= 0 && num_nodes < 0x7fffffff && num_feat > 0, "mgs_sage_aggr_bwd: bad sizes");
  MGS_REQUIRE(ldg >= num_feat && ldgx >= num_feat, "mgs_sage_aggr_bwd: leading dimension < num_feat");
  if (num_nodes == 0) return MGS_OK;
  MGS_REQUIRE(g && gx && rowptr && colptr, "mgs_sage_aggr_bwd: null pointer");
  MGS_REQUIRE(!edge_weight || permt, "mgs_sage_aggr_bwd: edge_weight needs permt");
  int V = min_int(vec_width(g, ldg, num_feat), vec_width(gx, ldgx, num_feat));
  if (base) V = min_int(V, vec_width(base, ldbase, num_feat));
  if (mask) V = min_int(V, vec_width(mask, ldmask, num_feat));
  const int chunks = num_feat / V;
  const int iters = iters_for(chunks);
  if (iters > 0) {
    stream::Args sa = {};
    sa.src = g; sa.lds = ldg; sa.dst = gx; sa.ldd = ldgx;
    sa.N = (int)num_nodes; sa.chunks = chunks; sa.H = 1; sa.C = num_feat;
    sa.ptr = colptr; sa.idx = row; sa.eid = permt; sa.ew = edge_weight; sa.rowptr = rowptr; sa.accumulate = base != nullptr; sa.base = base; sa.ldb = ldbase;
    sa.mask = mask; sa.ldm = ldmask;
    if (bits != nullptr) {
      MGS_REQUIRE(bits_words == V * iters, "mgs_sage_aggr_bwd_accumulate: relu_bits were written with another row layout "
                  "(%d words per row, this launch needs %d)", bits_words, V * iters);
      sa.bits_in = bits;
    }
    return stream::launch<stream::SAGE_BWD>(sa, V, iters, (cudaStream_t)stream_, "sage_aggr_bwd(stream)");
  }
  MGS_REQUIRE(base == nullptr && mask == nullptr && bits == nullptr, "mgs_sage_aggr_bwd_accumulate: rows wider than %d floats are not supported", 8 * 32 * 4);
  const int grid = grid_for(num_nodes * chunks, kThreads, 8);
  return dispatch<true>(V, edge_weight != nullptr, grid, (cudaStream_t)stream_, g, ldg, (int)num_nodes, chunks,
                        rowptr, colptr, row, permt, edge_weight, gx, ldgx);
}

extern "C" int mgs_sage_aggr_bwd(const float* g, int64_t ldg, int64_t num_nodes, int32_t num_feat,
                                 const int32_t* rowptr, const int32_t* colptr, const int32_t* row,
                                 const int32_t* permt, const float* edge_weight, float* gx, int64_t ldgx,
                                 mgs_stream_t stream_) {
  return sage_aggr_bwd_impl(nullptr, 0, nullptr, 0, nullptr, 0, g, ldg, num_nodes, num_feat, rowptr, colptr, row, permt, edge_weight, gx, ldgx, stream_);
}

extern "C" int mgs_sage_aggr_bwd_accumulate(const float* g, int64_t ldg, int64_t num_nodes, int32_t num_feat,
                                            const int32_t* rowptr, const int32_t* colptr, const int32_t* row,
                                            const int32_t* permt, const float* edge_weight, const float* base,
                                            int64_t ldbase, const float* relu_mask, int64_t ldmask,
                                            const uint32_t* relu_bits, int32_t bits_words, float* gx,
                                            int64_t ldgx, mgs_stream_t stream_) {
  MGS_REQUIRE(base != nullptr && ldbase >= num_feat, "mgs_sage_aggr_bwd_accumulate: base matrix missing");
  MGS_REQUIRE(relu_mask == nullptr || ldmask >= num_feat, "mgs_sage_aggr_bwd_accumulate: mask leading dimension < num_feat");
  MGS_REQUIRE(relu_mask == nullptr || relu_bits == nullptr, "mgs_sage_aggr_bwd_accumulate: one mask form at a time");
  return sage_aggr_bwd_impl(relu_bits, bits_words, relu_mask, ldmask, base, ldbase, g, ldg, num_nodes, num_feat, rowptr, colptr, row, permt, edge_weight, gx,
                            ldgx, stream_);
}

// GCNConv / GINConv neighbourhood sum (gnn/gcn.py:46-48, gnn/gat-gcn.py:58, gnn/gin.py:64-77): the same block-streamed
// gather without the division by the in-degree; with the CSC arrays it is its own backward.
extern "C" int mgs_sum_aggr(const float* src, int64_t lds, int64_t num_nodes, int32_t num_feat, const int32_t* ptr,
                            const int32_t* idx, const int32_t* eid, const float* edge_weight, const float* base,
                            int64_t ldbase, float* dst, int64_t ldd, mgs_stream_t stream_) {
  MGS_REQUIRE(num_nodes >= 0 && num_nodes < 0x7fffffff && num_feat > 0, "mgs_sum_aggr: bad sizes");
  MGS_REQUIRE(lds >= num_feat && ldd >= num_feat && (!base || ldbase >= num_feat), "mgs_sum_aggr: leading dimension < num_feat");
  if (num_nodes == 0) return MGS_OK;
  MGS_REQUIRE(src && dst && ptr, "mgs_sum_aggr: null pointer");
  MGS_REQUIRE(!edge_weight || eid, "mgs_sum_aggr: edge_weight needs the edge id array");
  int V = min_int(vec_width(src, lds, num_feat), vec_width(dst, ldd, num_feat));
  if (base) V = min_int(V, vec_width(base, ldbase, num_feat));
  const int chunks = num_feat / V;
  const int iters = iters_for(chunks);
  MGS_REQUIRE(iters > 0, "mgs_sum_aggr: rows wider than %d floats are not supported", 8 * 32 * 4);
  stream::Args sa = {};
  sa.src = src; sa.lds = lds; sa.dst = dst; sa.ldd = ldd;
  sa.N = (int)num_nodes; sa.chunks = chunks; sa.H = 1; sa.C = num_feat;
  sa.ptr = ptr; sa.idx = idx; sa.eid = eid; sa.ew = edge_weight;
  sa.accumulate = base != nullptr; sa.base = base; sa.ldb = ldbase;
  return stream::launch<stream::SUM>(sa, V, iters, (cudaStream_t)stream_, "sum_aggr(stream)");
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
