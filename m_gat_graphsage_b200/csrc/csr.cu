// K0 -- sorted-CSR / segment-pointer builder (integer, bit-exact).
//
// Result is identical to  perm = argsort(dst, stable); rowptr = [0, cumsum(bincount(dst))];
// col = src[perm]  (and the same by source), for ANY edge order and multiplicity:
//   1. count:  in-/out-degree histograms with integer atomics; the value returned by the atomic is
//              kept as a provisional (unordered) rank of the edge inside its row;
//   2. scan:   three-phase block scan of both histograms -> rowptr / colptr;
//   3. fill:   perm[rowptr[dst] + rank] = edge id  (row content complete, order arbitrary);
//   4. sort:   every row is sorted by edge id (rows are <= 6 long for molecules; insertion sort,
//              heap sort beyond 32 entries) -> the unique stable order, independent of atomic order;
//              col/row gathered, inverse permutation recorded;
//   5. csc_pos[q] = inv[permt[q]].
// HBM-bound integer work: 16E + 8N bytes read, 4(N+1)*2 + 20E written; no tensor cores involved.
#include "common.cuh"

namespace mgs {
namespace {

constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanChunk = kScanThreads * kScanItems;

__global__ void csr_count_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                 int E, int N, int* __restrict__ rowptr, int* __restrict__ colptr,
                                 int* __restrict__ rank_in, int* __restrict__ rank_out,
                                 int* __restrict__ status) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    int64_t s = src[e], d = dst[e];
    if (s < 0 || s >= N || d < 0 || d >= N) {
      atomicOr(status, MGS_STATUS_EDGE_OUT_OF_RANGE);
      s = s < 0 ? 0 : (s >= N ? N - 1 : s);   // stay memory-safe; result is flagged invalid
      d = d < 0 ? 0 : (d >= N ? N - 1 : d);
    }
    rank_in[e] = atomicAdd(&rowptr[(int)d + 1], 1);
    rank_out[e] = atomicAdd(&colptr[(int)s + 1], 1);
  }
}

// ---- three-phase inclusive scan over data[1..n] for two arrays (blockIdx.y selects) ----------
__device__ __forceinline__ int block_inclusive_scan(int v, int* smem_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) smem_warp[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int w = smem_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    smem_warp[lane] = w;
  }
  __syncthreads();
  if (warp > 0) v += smem_warp[warp - 1];
  return v;
}

__global__ void scan_reduce_kernel(const int* __restrict__ a0, const int* __restrict__ a1, int n,
                                   int* __restrict__ blocksum, int nblocks) {
  __shared__ int sw[32];
  const int* a = blockIdx.y == 0 ? a0 : a1;
  int base = blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (base + k < n) s += a[1 + base + k];
  int incl = block_inclusive_scan(s, sw);
  if (threadIdx.x == kScanThreads - 1) blocksum[blockIdx.y * nblocks + blockIdx.x] = incl;
}

__global__ void scan_blocksums_kernel(int* __restrict__ blocksum, int nblocks) {
  __shared__ int sw[32];
  __shared__ int carry_s;
  int* b = blocksum + blockIdx.y * nblocks;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += kScanThreads) {
    int i = base + threadIdx.x;
    int v = i < nblocks ? b[i] : 0;
    int incl = block_inclusive_scan(v, sw);
    int carry = carry_s;
    if (i < nblocks) b[i] = carry + incl - v;  // exclusive
    __syncthreads();
    if (threadIdx.x == kScanThreads - 1) carry_s = carry + incl;
    __syncthreads();
  }
}

__global__ void scan_apply_kernel(int* __restrict__ a0, int* __restrict__ a1, int n,
                                  const int* __restrict__ blocksum, int nblocks) {
  __shared__ int sw[32];
  int* a = blockIdx.y == 0 ? a0 : a1;
  int base = blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (base + k < n) ? a[1 + base + k] : 0;
    s += v[k];
  }
  int incl = block_inclusive_scan(s, sw);
  int run = incl - s + (blocksum ? blocksum[blockIdx.y * nblocks + blockIdx.x] : 0);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    run += v[k];
    if (base + k < n) a[1 + base + k] = run;
  }
}

__global__ void csr_fill_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                int E, int N, const int* __restrict__ rowptr, const int* __restrict__ colptr,
                                const int* __restrict__ rank_in, const int* __restrict__ rank_out,
                                int* __restrict__ perm, int* __restrict__ permt) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    int64_t s = src[e], d = dst[e];
    s = s < 0 ? 0 : (s >= N ? N - 1 : s);
    d = d < 0 ? 0 : (d >= N ? N - 1 : d);
    perm[rowptr[(int)d] + rank_in[e]] = e;
    permt[colptr[(int)s] + rank_out[e]] = e;
  }
}

__device__ void sort_segment(int* a, int n) {
  if (n <= 32) {
    for (int i = 1; i < n; ++i) {
      int key = a[i], j = i - 1;
      while (j >= 0 && a[j] > key) { a[j + 1] = a[j]; --j; }
      a[j + 1] = key;
    }
    return;
  }
  // heap sort for the (non-molecular) high-degree case
  for (int start = n / 2 - 1; start >= 0; --start) {
    int root = start;
    for (;;) {
      int child = 2 * root + 1;
      if (child >= n) break;
      if (child + 1 < n && a[child] < a[child + 1]) ++child;
      if (a[root] >= a[child]) break;
      int t = a[root]; a[root] = a[child]; a[child] = t;
      root = child;
    }
  }
  for (int end = n - 1; end > 0; --end) {
    int t = a[0]; a[0] = a[end]; a[end] = t;
    int root = 0;
    for (;;) {
      int child = 2 * root + 1;
      if (child >= end) break;
      if (child + 1 < end && a[child] < a[child + 1]) ++child;
      if (a[root] >= a[child]) break;
      int u = a[root]; a[root] = a[child]; a[child] = u;
      root = child;
    }
  }
}

// thread t < N sorts row t of the by-destination structure, thread N + t row t of the by-source one
__global__ void csr_sort_rows_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int N,
                                     const int* __restrict__ rowptr, const int* __restrict__ colptr,
                                     int* __restrict__ perm, int* __restrict__ permt,
                                     int* __restrict__ col, int* __restrict__ row, int* __restrict__ inv) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < 2 * N; t += gridDim.x * blockDim.x) {
    const bool by_dst = t < N;
    const int i = by_dst ? t : t - N;
    const int* ptr = by_dst ? rowptr : colptr;
    int* p = by_dst ? perm : permt;
    const int beg = ptr[i], end = ptr[i + 1];
    sort_segment(p + beg, end - beg);
    for (int q = beg; q < end; ++q) {
      int e = p[q];
      if (by_dst) {
        int64_t s = src[e];
        col[q] = (int)(s < 0 ? 0 : (s >= N ? N - 1 : s));
        inv[e] = q;
      } else {
        int64_t d = dst[e];
        row[q] = (int)(d < 0 ? 0 : (d >= N ? N - 1 : d));
      }
    }
  }
}

__global__ void csr_csc_pos_kernel(const int* __restrict__ permt, const int* __restrict__ inv, int E,
                                   int* __restrict__ csc_pos) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < E; q += gridDim.x * blockDim.x)
    csc_pos[q] = inv[permt[q]];
}

__global__ void graph_ptr_kernel(const int64_t* __restrict__ batch, int N, int B, int* __restrict__ gptr,
                                 int* __restrict__ status) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= N; i += gridDim.x * blockDim.x) {
    int64_t prev = (i == 0) ? -1 : batch[i - 1];
    int64_t cur = (i == N) ? B : batch[i];
    if (i < N && (cur < 0 || cur >= B)) atomicOr(status, MGS_STATUS_BATCH_OUT_OF_RANGE);
    if (cur < prev) atomicOr(status, MGS_STATUS_BATCH_NOT_SORTED);
    prev = prev < -1 ? -1 : (prev >= B ? B - 1 : prev);
    cur = cur < 0 ? 0 : (cur > B ? B : cur);
    for (int64_t g = prev + 1; g <= cur; ++g)
      if (g <= B) gptr[g] = i;
  }
}

struct Workspace {
  int* rank_in;
  int* rank_out;
  int* inv;
  int* blocksum;
  int nblocks;
  size_t bytes;
};

Workspace carve(void* base, int64_t N, int64_t E) {
  Workspace w;
  w.nblocks = (int)((N + kScanChunk - 1) / kScanChunk);
  if (w.nblocks < 1) w.nblocks = 1;
  size_t e_al = ((size_t)E + 3) & ~(size_t)3;
  char* p = (char*)base;
  w.rank_in = (int*)p;  p += e_al * sizeof(int);
  w.rank_out = (int*)p; p += e_al * sizeof(int);
  w.inv = (int*)p;      p += e_al * sizeof(int);
  w.blocksum = (int*)p; p += (size_t)2 * w.nblocks * sizeof(int);
  w.bytes = (size_t)(p - (char*)base);
  return w;
}

}  // namespace
}  // namespace mgs

using namespace mgs;

extern "C" size_t mgs_csr_workspace_bytes(int64_t num_nodes, int64_t num_edges) {
  if (num_nodes < 0 || num_edges < 0) return 0;
  return carve(nullptr, num_nodes, num_edges).bytes + 16;
}

extern "C" int mgs_csr_build(const int64_t* edge_index, int64_t edge_row_stride, int64_t num_edges,
                             int64_t num_nodes, int32_t* rowptr, int32_t* col, int32_t* perm,
                             int32_t* colptr, int32_t* row, int32_t* permt, int32_t* csc_pos,
                             int32_t* status, void* workspace, size_t workspace_bytes,
                             mgs_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MGS_REQUIRE(num_nodes >= 0 && num_edges >= 0, "mgs_csr_build: negative size");
  MGS_REQUIRE(num_nodes + num_edges < (int64_t)0x7fffffff, "mgs_csr_build: N + E must fit in int32");
  MGS_REQUIRE(rowptr && colptr && status, "mgs_csr_build: null output pointer");
  MGS_REQUIRE(num_edges == 0 || (edge_index && col && perm && row && permt && csc_pos),
              "mgs_csr_build: null edge pointer");
  MGS_REQUIRE(num_edges == 0 || num_nodes > 0, "mgs_csr_build: edges without nodes");
  if (workspace_bytes < mgs_csr_workspace_bytes(num_nodes, num_edges) || (num_edges > 0 && !workspace)) {
    set_error("mgs_csr_build: workspace too small (%zu < %zu)", workspace_bytes,
              mgs_csr_workspace_bytes(num_nodes, num_edges));
    return MGS_ERR_WORKSPACE_TOO_SMALL;
  }
  const int N = (int)num_nodes, E = (int)num_edges;
  Workspace w = carve(workspace, N, E);
  const int64_t* src = edge_index;
  const int64_t* dst = edge_index + edge_row_stride;

  MGS_CUDA(cudaMemsetAsync(rowptr, 0, sizeof(int) * ((size_t)N + 1), stream));
  MGS_CUDA(cudaMemsetAsync(colptr, 0, sizeof(int) * ((size_t)N + 1), stream));
  if (E == 0) return MGS_OK;

  const int T = 256;
  csr_count_kernel<<<grid_for(E, T, 8), T, 0, stream>>>(src, dst, E, N, rowptr, colptr, w.rank_in,
                                                          w.rank_out, status);
  if (int rc = check_launch("csr_count_kernel")) return rc;

  dim3 sgrid(w.nblocks, 2);
  if (w.nblocks > 1) {
    scan_reduce_kernel<<<sgrid, kScanThreads, 0, stream>>>(rowptr, colptr, N, w.blocksum, w.nblocks);
    if (int rc = check_launch("scan_reduce_kernel")) return rc;
    scan_blocksums_kernel<<<dim3(1, 2), kScanThreads, 0, stream>>>(w.blocksum, w.nblocks);
    if (int rc = check_launch("scan_blocksums_kernel")) return rc;
  }
  scan_apply_kernel<<<sgrid, kScanThreads, 0, stream>>>(rowptr, colptr, N,
                                                         w.nblocks > 1 ? w.blocksum : nullptr, w.nblocks);
  if (int rc = check_launch("scan_apply_kernel")) return rc;

  csr_fill_kernel<<<grid_for(E, T, 8), T, 0, stream>>>(src, dst, E, N, rowptr, colptr, w.rank_in,
                                                         w.rank_out, perm, permt);
  if (int rc = check_launch("csr_fill_kernel")) return rc;
  csr_sort_rows_kernel<<<grid_for(2 * (int64_t)N, T, 8), T, 0, stream>>>(src, dst, N, rowptr, colptr, perm,
                                                                          permt, col, row, w.inv);
  if (int rc = check_launch("csr_sort_rows_kernel")) return rc;
  csr_csc_pos_kernel<<<grid_for(E, T, 8), T, 0, stream>>>(permt, w.inv, E, csc_pos);
  return check_launch("csr_csc_pos_kernel");
}

extern "C" int mgs_graph_ptr(const int64_t* batch, int64_t num_nodes, int64_t num_graphs, int32_t* gptr,
                             int32_t* status, mgs_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MGS_REQUIRE(num_nodes >= 0 && num_graphs >= 0, "mgs_graph_ptr: negative size");
  MGS_REQUIRE(num_nodes < 0x7fffffff && num_graphs < 0x7fffffff, "mgs_graph_ptr: sizes must fit in int32");
  MGS_REQUIRE(gptr && status && (num_nodes == 0 || batch), "mgs_graph_ptr: null pointer");
  const int T = 256;
  graph_ptr_kernel<<<grid_for(num_nodes + 1, T, 8), T, 0, stream>>>(batch, (int)num_nodes, (int)num_graphs,
                                                                     gptr, status);
  return check_launch("graph_ptr_kernel");
}


// ------------------------------------------------------------------------------------------------------------------
// Wire format of a molecule batch (data.WireBatch): the reference's atom features are one-hot groups, exactly 0.0 / 1.0
// (train.py:33-43), so a row of F <= 64 features crosses PCIe as F bits (8 bytes per atom instead of 140), edge_index as
// int32 and the batch vector as B + 1 segment pointers.  One launch expands all three on the device.
// ------------------------------------------------------------------------------------------------------------------
namespace mgs {
namespace {
__global__ void __launch_bounds__(256)
wire_expand_kernel(const unsigned long long* __restrict__ bits, int N, int F, float* __restrict__ x, int64_t ldx,
                   const int* __restrict__ ei32, int64_t E, long long* __restrict__ ei64,
                   const int* __restrict__ gptr, int B, long long* __restrict__ batch) {
  const int64_t tid = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t t = tid; t < (int64_t)N * F; t += stride) {          // x[n, f] = bit f of word n
    const int n = (int)(t / F), f = (int)(t - (int64_t)n * F);
    x[(int64_t)n * ldx + f] = (float)((__ldg(bits + n) >> f) & 1ull);
  }
  for (int64_t t = tid; t < 2 * E; t += stride) ei64[t] = (long long)__ldg(ei32 + t);
  for (int64_t t = tid; t < N; t += stride) {                        // molecule of atom t: last b with gptr[b] <= t
    int lo = 0, hi = B;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(gptr + mid) <= (int)t) lo = mid; else hi = mid;
    }
    batch[t] = lo;
  }
}
}  // namespace
}  // namespace mgs

extern "C" int mgs_wire_expand(const uint64_t* bits, int64_t num_nodes, int32_t num_feat, float* x, int64_t ldx,
                               const int32_t* edge_index32, int64_t num_edges, int64_t* edge_index64,
                               const int32_t* gptr, int64_t num_graphs, int64_t* batch, mgs_stream_t stream_) {
  using namespace mgs;
  MGS_REQUIRE(num_nodes >= 0 && num_nodes < 0x7fffffff && num_edges >= 0 && num_graphs >= 0 && num_graphs < 0x7fffffff,
              "mgs_wire_expand: bad sizes");
  MGS_REQUIRE(num_feat > 0 && num_feat <= 64 && ldx >= num_feat, "mgs_wire_expand: 1 <= num_feat <= 64 bits per atom");
  if (num_nodes == 0 && num_edges == 0) return MGS_OK;
  MGS_REQUIRE((num_nodes == 0 || (bits && x && batch && gptr && num_graphs > 0)) && (num_edges == 0 || (edge_index32 && edge_index64)),
              "mgs_wire_expand: null pointer");
  const int64_t work = num_nodes * num_feat > 2 * num_edges ? num_nodes * num_feat : 2 * num_edges;
  wire_expand_kernel<<<grid_for(work, 256, 8), 256, 0, (cudaStream_t)stream_>>>(
      (const unsigned long long*)bits, (int)num_nodes, num_feat, x, ldx, edge_index32, num_edges, (long long*)edge_index64,
      gptr, (int)num_graphs, (long long*)batch);
  return check_launch("wire_expand_kernel");
}
