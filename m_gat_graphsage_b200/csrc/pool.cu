// K3 -- segmented global max / mean / add pooling, forward and backward.
//
// One thread owns one VEC-wide feature chunk of one molecule and walks the molecule's 11-94 atom rows:
// a warp reads consecutive addresses of one row at a time (coalesced), every x element is read exactly
// once.  Rows are folded in ascending atom order from 0.0 (mean/add) like ATen's CPU scatter_add_, the
// mean uses a true division by max(count, 1); max of an empty molecule is 0 (include_self=False on a
// zero tensor).  Backward of max reproduces ATen's scatter_reduce('amax') gradient: split evenly over
// exact ties, the zero-initialised destination counting as one extra tie when the maximum is exactly 0.
// Algorithmic bytes / atom: 4F read (+ 4F*B/N written); backward 4F (x re-read) + 4F (gx) .
#include "common.cuh"

#include <cfloat>

namespace mgs {
namespace {

constexpr int kThreads = 256;

template <int V, int MODE>
__global__ void __launch_bounds__(kThreads)
pool_fwd_kernel(const float* __restrict__ x, int64_t ldx, const int* __restrict__ gptr, int B, int chunks,
                float* __restrict__ out, int64_t ldo) {
  const int64_t total = (int64_t)B * chunks;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (int64_t)gridDim.x * kThreads) {
    const int gidx = (int)(t / chunks);
    const int c = (int)(t - (int64_t)gidx * chunks) * V;
    const int beg = __ldg(gptr + gidx), end = __ldg(gptr + gidx + 1);
    Vec<V> acc;
#pragma unroll
    for (int u = 0; u < V; ++u) acc.v[u] = (MODE == MGS_POOL_MAX) ? -INFINITY : 0.f;
    int r = beg;
    for (; r + 4 <= end; r += 4) {
      Vec<V> v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = Vec<V>::load(x + (int64_t)(r + k) * ldx + c);
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int u = 0; u < V; ++u)
          acc.v[u] = (MODE == MGS_POOL_MAX) ? fmaxf(acc.v[u], v[k].v[u]) : __fadd_rn(acc.v[u], v[k].v[u]);
    }
    for (; r < end; ++r) {
      Vec<V> v = Vec<V>::load(x + (int64_t)r * ldx + c);
#pragma unroll
      for (int u = 0; u < V; ++u)
        acc.v[u] = (MODE == MGS_POOL_MAX) ? fmaxf(acc.v[u], v.v[u]) : __fadd_rn(acc.v[u], v.v[u]);
    }
    if (MODE == MGS_POOL_MAX) {
      if (end == beg) acc = vzero<V>();
    } else if (MODE == MGS_POOL_MEAN) {
      const float cnt = (float)max(end - beg, 1);
#pragma unroll
      for (int u = 0; u < V; ++u) acc.v[u] = __fdiv_rn(acc.v[u], cnt);
    }
    acc.store(out + (int64_t)gidx * ldo + c);
  }
}

template <int V, int MODE>
__global__ void __launch_bounds__(kThreads)
pool_bwd_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ x, int64_t ldx,
                const float* __restrict__ out, int64_t ldo, const int* __restrict__ gptr, int B, int chunks,
                float* __restrict__ gx, int64_t ldgx) {
  const int64_t total = (int64_t)B * chunks;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (int64_t)gridDim.x * kThreads) {
    const int gidx = (int)(t / chunks);
    const int c = (int)(t - (int64_t)gidx * chunks) * V;
    const int beg = __ldg(gptr + gidx), end = __ldg(gptr + gidx + 1);
    Vec<V> gv = Vec<V>::load(g + (int64_t)gidx * ldg + c);
    if (MODE == MGS_POOL_MAX) {
      Vec<V> m = Vec<V>::load(out + (int64_t)gidx * ldo + c);
      float ties[V];
#pragma unroll
      for (int u = 0; u < V; ++u) ties[u] = (m.v[u] == 0.f) ? 1.f : 0.f;   // the zero-initialised destination
      for (int r = beg; r < end; ++r) {
        Vec<V> v = Vec<V>::load(x + (int64_t)r * ldx + c);
#pragma unroll
        for (int u = 0; u < V; ++u) ties[u] += (v.v[u] == m.v[u]) ? 1.f : 0.f;
      }
      Vec<V> share;
#pragma unroll
      for (int u = 0; u < V; ++u) share.v[u] = __fdiv_rn(gv.v[u], ties[u]);
      for (int r = beg; r < end; ++r) {
        Vec<V> v = Vec<V>::load(x + (int64_t)r * ldx + c);
        Vec<V> o;
#pragma unroll
        for (int u = 0; u < V; ++u) o.v[u] = (v.v[u] == m.v[u]) ? share.v[u] : 0.f;
        o.store(gx + (int64_t)r * ldgx + c);
      }
    } else {
      if (MODE == MGS_POOL_MEAN) {
        const float cnt = (float)max(end - beg, 1);
#pragma unroll
        for (int u = 0; u < V; ++u) gv.v[u] = __fdiv_rn(gv.v[u], cnt);
      }
      for (int r = beg; r < end; ++r) gv.store(gx + (int64_t)r * ldgx + c);
    }
  }
}

// ---- max and mean in one pass (the readout `cat([gmp(x), gap(x)])` of ablation/model1.py:72) -------------------
// out[b, 0:F] = max, out[b, F:2F] = mean.  The two reference calls are two autograd nodes over the same x: two
// forward passes, two backward passes and an `add` of two [N, F] gradients (0.33 ms per step on B200).  Fused:
// x is read once forward; backward reads x twice (tie count, then write; the second read is an L1 / L2 hit for
// an 11-94 atom molecule) and writes gx = max part + mean part once -- the same fp32 add autograd would do.
template <int V>
__global__ void __launch_bounds__(kThreads)
pool_maxmean_fwd_kernel(const float* __restrict__ x, int64_t ldx, const int* __restrict__ gptr, int B, int chunks,
                        int F, float* __restrict__ out, int64_t ldo, float* __restrict__ ties) {
  const int64_t total = (int64_t)B * chunks;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (int64_t)gridDim.x * kThreads) {
    const int gidx = (int)(t / chunks);
    const int c = (int)(t - (int64_t)gidx * chunks) * V;
    const int beg = __ldg(gptr + gidx), end = __ldg(gptr + gidx + 1);
    Vec<V> mx, sm, tc;                                      // tc: how many rows attain the running maximum
#pragma unroll
    for (int u = 0; u < V; ++u) { mx.v[u] = -INFINITY; sm.v[u] = 0.f; tc.v[u] = 0.f; }
    int r = beg;
    // eight rows in flight, then four, then one (a 32-atom molecule is 4 dependent round trips instead of 8; same row
    // order, same bits)
    for (; r + 8 <= end; r += 8) {
      Vec<V> v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = Vec<V>::load(x + (int64_t)(r + k) * ldx + c);
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int u = 0; u < V; ++u) {
          tc.v[u] = v[k].v[u] > mx.v[u] ? 1.f : (v[k].v[u] == mx.v[u] ? tc.v[u] + 1.f : tc.v[u]);
          mx.v[u] = fmaxf(mx.v[u], v[k].v[u]);
          sm.v[u] = __fadd_rn(sm.v[u], v[k].v[u]);
        }
    }
    for (; r + 4 <= end; r += 4) {
      Vec<V> v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = Vec<V>::load(x + (int64_t)(r + k) * ldx + c);
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int u = 0; u < V; ++u) {
          tc.v[u] = v[k].v[u] > mx.v[u] ? 1.f : (v[k].v[u] == mx.v[u] ? tc.v[u] + 1.f : tc.v[u]);
          mx.v[u] = fmaxf(mx.v[u], v[k].v[u]);
          sm.v[u] = __fadd_rn(sm.v[u], v[k].v[u]);
        }
    }
    for (; r < end; ++r) {
      Vec<V> v = Vec<V>::load(x + (int64_t)r * ldx + c);
#pragma unroll
      for (int u = 0; u < V; ++u) {
        tc.v[u] = v.v[u] > mx.v[u] ? 1.f : (v.v[u] == mx.v[u] ? tc.v[u] + 1.f : tc.v[u]);
        mx.v[u] = fmaxf(mx.v[u], v.v[u]);
        sm.v[u] = __fadd_rn(sm.v[u], v.v[u]);
      }
    }
    if (end == beg) { mx = vzero<V>(); tc = vzero<V>(); }
#pragma unroll
    for (int u = 0; u < V; ++u)
      if (mx.v[u] == 0.f) tc.v[u] += 1.f;                   // the zero-initialised destination of amax ties too
    const float cnt = (float)max(end - beg, 1);
#pragma unroll
    for (int u = 0; u < V; ++u) sm.v[u] = __fdiv_rn(sm.v[u], cnt);
    mx.store(out + (int64_t)gidx * ldo + c);
    sm.store(out + (int64_t)gidx * ldo + F + c);
    if (ties != nullptr) tc.store(ties + (int64_t)gidx * F + c);
  }
}

template <int V>
__global__ void __launch_bounds__(kThreads)
pool_maxmean_bwd_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ x, int64_t ldx,
                        const float* __restrict__ out, int64_t ldo, const int* __restrict__ gptr, int B, int chunks,
                        int F, float* __restrict__ gx, int64_t ldgx, const float* __restrict__ ties_in, int relu_mask) {
  const int64_t total = (int64_t)B * chunks;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (int64_t)gridDim.x * kThreads) {
    const int gidx = (int)(t / chunks);
    const int c = (int)(t - (int64_t)gidx * chunks) * V;
    const int beg = __ldg(gptr + gidx), end = __ldg(gptr + gidx + 1);
    const Vec<V> gmax = Vec<V>::load(g + (int64_t)gidx * ldg + c);
    Vec<V> gmean = Vec<V>::load(g + (int64_t)gidx * ldg + F + c);
    const Vec<V> m = Vec<V>::load(out + (int64_t)gidx * ldo + c);
    const float cnt = (float)max(end - beg, 1);
    float ties[V];
#pragma unroll
    for (int u = 0; u < V; ++u) {
      ties[u] = (m.v[u] == 0.f) ? 1.f : 0.f;                 // the zero-initialised destination of amax
      gmean.v[u] = __fdiv_rn(gmean.v[u], cnt);
    }
    int r = beg;
    if (ties_in != nullptr) {                               // counted by the forward pass: x is read once here
      const Vec<V> t = Vec<V>::load(ties_in + (int64_t)gidx * F + c);
#pragma unroll
      for (int u = 0; u < V; ++u) ties[u] = t.v[u];
      r = end;
    }
    for (; r + 4 <= end; r += 4) {
      Vec<V> v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = Vec<V>::load(x + (int64_t)(r + k) * ldx + c);
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int u = 0; u < V; ++u) ties[u] += (v[k].v[u] == m.v[u]) ? 1.f : 0.f;
    }
    for (; r < end; ++r) {
      Vec<V> v = Vec<V>::load(x + (int64_t)r * ldx + c);
#pragma unroll
      for (int u = 0; u < V; ++u) ties[u] += (v.v[u] == m.v[u]) ? 1.f : 0.f;
    }
    Vec<V> share;
#pragma unroll
    for (int u = 0; u < V; ++u) share.v[u] = __fdiv_rn(gmax.v[u], ties[u]);
    r = beg;
    for (; r + 8 <= end; r += 8) {                           // eight rows in flight (see the forward kernel)
      Vec<V> v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = Vec<V>::load(x + (int64_t)(r + k) * ldx + c);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        Vec<V> o;
#pragma unroll
        for (int u = 0; u < V; ++u) {
          o.v[u] = __fadd_rn((v[k].v[u] == m.v[u]) ? share.v[u] : 0.f, gmean.v[u]);
          if (relu_mask && v[k].v[u] <= 0.f) o.v[u] = 0.f;
        }
        o.store(gx + (int64_t)(r + k) * ldgx + c);
      }
    }
    for (; r + 4 <= end; r += 4) {
      Vec<V> v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = Vec<V>::load(x + (int64_t)(r + k) * ldx + c);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        Vec<V> o;
#pragma unroll
        for (int u = 0; u < V; ++u) {
          o.v[u] = __fadd_rn((v[k].v[u] == m.v[u]) ? share.v[u] : 0.f, gmean.v[u]);
          if (relu_mask && v[k].v[u] <= 0.f) o.v[u] = 0.f;      // backward of the ReLU that produced x (threshold_backward)
        }
        o.store(gx + (int64_t)(r + k) * ldgx + c);
      }
    }
    for (; r < end; ++r) {
      Vec<V> v = Vec<V>::load(x + (int64_t)r * ldx + c);
      Vec<V> o;
#pragma unroll
      for (int u = 0; u < V; ++u) {
        o.v[u] = __fadd_rn((v.v[u] == m.v[u]) ? share.v[u] : 0.f, gmean.v[u]);
        if (relu_mask && v.v[u] <= 0.f) o.v[u] = 0.f;
      }
      o.store(gx + (int64_t)r * ldgx + c);
    }
  }
}

}  // namespace
}  // namespace mgs

using namespace mgs;

#define MGS_POOL_DISPATCH(KERNEL, ...)                                                          \
  do {                                                                                          \
    if (V == 4) {                                                                               \
      if (mode == MGS_POOL_MAX) KERNEL<4, MGS_POOL_MAX><<<grid, kThreads, 0, stream>>>(__VA_ARGS__);        \
      else if (mode == MGS_POOL_MEAN) KERNEL<4, MGS_POOL_MEAN><<<grid, kThreads, 0, stream>>>(__VA_ARGS__); \
      else KERNEL<4, MGS_POOL_ADD><<<grid, kThreads, 0, stream>>>(__VA_ARGS__);                 \
    } else if (V == 2) {                                                                        \
      if (mode == MGS_POOL_MAX) KERNEL<2, MGS_POOL_MAX><<<grid, kThreads, 0, stream>>>(__VA_ARGS__);        \
      else if (mode == MGS_POOL_MEAN) KERNEL<2, MGS_POOL_MEAN><<<grid, kThreads, 0, stream>>>(__VA_ARGS__); \
      else KERNEL<2, MGS_POOL_ADD><<<grid, kThreads, 0, stream>>>(__VA_ARGS__);                 \
    } else {                                                                                    \
      if (mode == MGS_POOL_MAX) KERNEL<1, MGS_POOL_MAX><<<grid, kThreads, 0, stream>>>(__VA_ARGS__);        \
      else if (mode == MGS_POOL_MEAN) KERNEL<1, MGS_POOL_MEAN><<<grid, kThreads, 0, stream>>>(__VA_ARGS__); \
      else KERNEL<1, MGS_POOL_ADD><<<grid, kThreads, 0, stream>>>(__VA_ARGS__);                 \
    }                                                                                           \
  } while (0)

extern "C" int mgs_pool_fwd(const float* x, int64_t ldx, const int32_t* gptr, int64_t num_graphs,
                            int32_t num_feat, int32_t mode, float* out, int64_t ldo, mgs_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MGS_REQUIRE(num_graphs >= 0 && num_graphs < 0x7fffffff && num_feat > 0, "mgs_pool_fwd: bad sizes");
  MGS_REQUIRE(mode >= MGS_POOL_MAX && mode <= MGS_POOL_ADD, "mgs_pool_fwd: unknown mode %d", mode);
  MGS_REQUIRE(ldx >= num_feat && ldo >= num_feat, "mgs_pool_fwd: leading dimension < num_feat");
  if (num_graphs == 0) return MGS_OK;
  MGS_REQUIRE(gptr && out, "mgs_pool_fwd: null pointer");
  const int V = min_int(vec_width(x, ldx, num_feat), vec_width(out, ldo, num_feat));
  const int chunks = num_feat / V;
  const int B = (int)num_graphs;
  const int grid = grid_for(num_graphs * chunks, kThreads, 8);
  MGS_POOL_DISPATCH(pool_fwd_kernel, x, ldx, gptr, B, chunks, out, ldo);
  return check_launch("pool_fwd_kernel");
}

extern "C" int mgs_pool_bwd(const float* g, int64_t ldg, const float* x, int64_t ldx, const float* out,
                            int64_t ldo, const int32_t* gptr, int64_t num_graphs, int32_t num_feat,
                            int32_t mode, float* gx, int64_t ldgx, mgs_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MGS_REQUIRE(num_graphs >= 0 && num_graphs < 0x7fffffff && num_feat > 0, "mgs_pool_bwd: bad sizes");
  MGS_REQUIRE(mode >= MGS_POOL_MAX && mode <= MGS_POOL_ADD, "mgs_pool_bwd: unknown mode %d", mode);
  MGS_REQUIRE(ldg >= num_feat && ldgx >= num_feat, "mgs_pool_bwd: leading dimension < num_feat");
  if (num_graphs == 0) return MGS_OK;
  MGS_REQUIRE(g && gptr && gx, "mgs_pool_bwd: null pointer");
  int V = min_int(vec_width(g, ldg, num_feat), vec_width(gx, ldgx, num_feat));
  if (mode == MGS_POOL_MAX) {
    MGS_REQUIRE(x && out && ldx >= num_feat && ldo >= num_feat, "mgs_pool_bwd: max mode needs x and out");
    V = min_int(V, min_int(vec_width(x, ldx, num_feat), vec_width(out, ldo, num_feat)));
  }
  const int chunks = num_feat / V;
  const int B = (int)num_graphs;
  const int grid = grid_for(num_graphs * chunks, kThreads, 8);
  MGS_POOL_DISPATCH(pool_bwd_kernel, g, ldg, x, ldx, out, ldo, gptr, B, chunks, gx, ldgx);
  return check_launch("pool_bwd_kernel");
}

extern "C" int mgs_pool_maxmean_fwd(const float* x, int64_t ldx, const int32_t* gptr, int64_t num_graphs,
                                    int32_t num_feat, float* out, int64_t ldo, float* ties, mgs_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MGS_REQUIRE(num_graphs >= 0 && num_graphs < 0x7fffffff && num_feat > 0, "mgs_pool_maxmean_fwd: bad sizes");
  MGS_REQUIRE(ldx >= num_feat && ldo >= 2 * (int64_t)num_feat, "mgs_pool_maxmean_fwd: leading dimension too small");
  if (num_graphs == 0) return MGS_OK;
  MGS_REQUIRE(gptr && out, "mgs_pool_maxmean_fwd: null pointer");
  int V = min_int(vec_width(x, ldx, num_feat), vec_width(out, ldo, num_feat));
  if (ties) V = min_int(V, vec_width(ties, num_feat, num_feat));
  const int chunks = num_feat / V;
  const int grid = grid_for(num_graphs * chunks, kThreads, 8);
  if (V == 4) pool_maxmean_fwd_kernel<4><<<grid, kThreads, 0, stream>>>(x, ldx, gptr, (int)num_graphs, chunks, num_feat, out, ldo, ties);
  else if (V == 2) pool_maxmean_fwd_kernel<2><<<grid, kThreads, 0, stream>>>(x, ldx, gptr, (int)num_graphs, chunks, num_feat, out, ldo, ties);
  else pool_maxmean_fwd_kernel<1><<<grid, kThreads, 0, stream>>>(x, ldx, gptr, (int)num_graphs, chunks, num_feat, out, ldo, ties);
  return check_launch("pool_maxmean_fwd_kernel");
}

extern "C" int mgs_pool_maxmean_bwd(const float* g, int64_t ldg, const float* x, int64_t ldx, const float* out,
                                    int64_t ldo, const int32_t* gptr, int64_t num_graphs, int32_t num_feat,
                                    float* gx, int64_t ldgx, const float* ties, int32_t relu_mask,
                                    mgs_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MGS_REQUIRE(num_graphs >= 0 && num_graphs < 0x7fffffff && num_feat > 0, "mgs_pool_maxmean_bwd: bad sizes");
  MGS_REQUIRE(ldg >= 2 * (int64_t)num_feat && ldo >= 2 * (int64_t)num_feat && ldx >= num_feat && ldgx >= num_feat,
              "mgs_pool_maxmean_bwd: leading dimension too small");
  if (num_graphs == 0) return MGS_OK;
  MGS_REQUIRE(g && x && out && gptr && gx, "mgs_pool_maxmean_bwd: null pointer");
  int V = min_int(min_int(vec_width(g, ldg, num_feat), vec_width(gx, ldgx, num_feat)),
                  min_int(vec_width(x, ldx, num_feat), vec_width(out, ldo, num_feat)));
  if (ties) V = min_int(V, vec_width(ties, num_feat, num_feat));
  const int chunks = num_feat / V;
  const int grid = grid_for(num_graphs * chunks, kThreads, 8);
  if (V == 4) pool_maxmean_bwd_kernel<4><<<grid, kThreads, 0, stream>>>(g, ldg, x, ldx, out, ldo, gptr, (int)num_graphs, chunks, num_feat, gx, ldgx, ties, relu_mask);
  else if (V == 2) pool_maxmean_bwd_kernel<2><<<grid, kThreads, 0, stream>>>(g, ldg, x, ldx, out, ldo, gptr, (int)num_graphs, chunks, num_feat, gx, ldgx, ties, relu_mask);
  else pool_maxmean_bwd_kernel<1><<<grid, kThreads, 0, stream>>>(g, ldg, x, ldx, out, ldo, gptr, (int)num_graphs, chunks, num_feat, gx, ldgx, ties, relu_mask);
  return check_launch("pool_maxmean_bwd_kernel");
}
