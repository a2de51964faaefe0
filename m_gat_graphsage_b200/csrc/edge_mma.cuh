// K2 backward, edge part, tensor-core variant (round 2): the per-head dot products of edge.cuh as a small MMA.
//
//   dalpha[slot, h] = < g_i[h, :], xh_j[h, :] >  =  sum_col  P[slot, col] * Ind[col, h],   P[slot, col] = g_i[col] xh_j[col],
//   Ind[col, h] = 1 if col / C == h else 0
//
// edge.cuh keeps the gathered rows in the coalesced lane mapping and has to send them through shared memory to line
// them up with the heads (12 scalar LDS + FMA per neighbour and lane, ~600 dependent instructions per destination row:
// 28 % of HBM peak, ncu: 96 M warp instructions, issue 39 % at 16 warps / SM).  Here the head sums are done by the tensor
// core: a warp owns 16 consecutive SLOTS (the per-edge arrays are in slot order, so a tile is a contiguous run of in-edges
// + self loops of consecutive destination atoms); for every 8-column slab each thread loads 2 + 2 float2 (its two slots'
// xh_j and g_i, columns 2 tig, 2 tig + 1 -- the slab's K indices are permuted so that a thread's two K values are adjacent
// columns), multiplies, splits the products into TF32 hi + lo and issues mma.sync m16n8k8 against the 0 / 1 indicator
// fragment (exact in TF32), which it computes arithmetically.  No shared memory, no shuffles in the main loop; every 32-byte
// sector fetched is used completely.  The softmax Jacobian / LeakyReLU' / da_dst need whole destination rows and run as a
// second, small thread-per-(atom, head) kernel over the slot-ordered dalpha.
// Used when there are no edge weights / attention-dropout mask (training and importance passes); edge.cuh otherwise.
#pragma once

#include "common.cuh"

namespace mgs {
namespace emma {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t tf32_hi(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }

// NT = number of 8-head column tiles (heads <= 8 NT)
template <int NT>
__global__ void __launch_bounds__(kThreads)
gat_bwd_edge_mma_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ xh, int64_t ld, int N, int H, int C,
                        const int* __restrict__ rowptr, const int* __restrict__ col, float* __restrict__ dalpha) {
  const int S = __ldg(rowptr + N) + N;                          // slots: every in-edge + one self loop per atom
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  const int warp_global = blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * kWarps;
  const int HC = H * C;
  const int ntiles = (S + 15) / 16;
  const int nslabs = (HC + 7) / 8;

  for (int tile = warp_global; tile < ntiles; tile += nwarps) {
    // ---- my two slots: destination atom i (binary search over slot starts rowptr[i] + i) and source atom j ----
    const float* gp[2];
    const float* xp[2];
    bool ok[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int s = tile * 16 + gid + 8 * r;
      ok[r] = s < S;
      int i = 0, j = 0;
      if (ok[r]) {
        int lo = 0, hi = N;
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (__ldg(rowptr + mid) + mid <= s) lo = mid; else hi = mid;
        }
        i = lo;
        const int beg = __ldg(rowptr + i);
        const int k = s - (beg + i);
        j = (k == __ldg(rowptr + i + 1) - beg) ? i : __ldg(col + beg + k);     // last slot of a row: the self loop
      }
      gp[r] = g + (int64_t)i * ldg + 2 * tig;
      xp[r] = xh + (int64_t)j * ld + 2 * tig;
    }
    float acc[NT][4];
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[t][e] = 0.f;

    // head of column c = 8 slab + 2 tig (and of c + 1), tracked incrementally
    int h0 = (2 * tig) / C, r0 = (2 * tig) - h0 * C;              // column c:     head h0, channel r0
    int h1 = (2 * tig + 1) / C, r1 = (2 * tig + 1) - h1 * C;      // column c + 1
    constexpr int U = 4;                                          // slabs in flight
    for (int sb = 0; sb < nslabs; sb += U) {
      float2 gv[U][2], xv[U][2];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = 8 * (sb + u) + 2 * tig;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          gv[u][r] = make_float2(0.f, 0.f);
          xv[u][r] = make_float2(0.f, 0.f);
          if (ok[r] && c < HC) {                                  // HC even (checked by the launcher): c + 1 < HC too
            gv[u][r] = __ldg(reinterpret_cast<const float2*>(gp[r] + 8 * (sb + u)));
            xv[u][r] = __ldg(reinterpret_cast<const float2*>(xp[r] + 8 * (sb + u)));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (sb + u < nslabs) {
          // A fragment: a0 (row gid, k tig) a1 (row gid + 8, k tig) a2 (row gid, k tig + 4) a3 (row gid + 8, k tig + 4);
          // K index tig <-> column c, K index tig + 4 <-> column c + 1
          const float p[4] = {gv[u][0].x * xv[u][0].x, gv[u][1].x * xv[u][1].x, gv[u][0].y * xv[u][0].y, gv[u][1].y * xv[u][1].y};
          uint32_t ahi[4], alo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            ahi[e] = tf32_hi(p[e]);
            alo[e] = __float_as_uint(p[e] - __uint_as_float(ahi[e]));
          }
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            // B fragment: b0 (k tig, n gid) b1 (k tig + 4, n gid): the indicator of "column's head == 8 t + gid"
            const uint32_t b0 = (h0 == 8 * t + gid) ? 0x3f800000u : 0u;
            const uint32_t b1 = (h1 == 8 * t + gid) ? 0x3f800000u : 0u;
            mma_tf32(acc[t], ahi, b0, b1);
            mma_tf32(acc[t], alo, b0, b1);
          }
        }
        r0 += 8; while (r0 >= C) { r0 -= C; ++h0; }
        r1 += 8; while (r1 >= C) { r1 -= C; ++h1; }
      }
    }
    // ---- C fragment: c0 (row gid, n 2 tig) c1 (row gid, n 2 tig + 1) c2 / c3 (row gid + 8) ----
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int s = tile * 16 + gid + 8 * (e >> 1);
        const int h = 8 * t + 2 * tig + (e & 1);
        if (s < S && h < H) dalpha[(int64_t)s * H + h] = acc[t][e];
      }
  }
}

// dr[slot, h] (in: d alpha, out: d raw score) and da_dst[i, h]; thread per (atom, head).  Same arithmetic and summation
// order as edge.cuh's row epilogue.
__global__ void __launch_bounds__(256)
gat_bwd_edge_softmax_kernel(const float* __restrict__ alpha, const float* __restrict__ a_src, const float* __restrict__ a_dst,
                            float slope, const int* __restrict__ rowptr, const int* __restrict__ col, int N, int H,
                            float* __restrict__ dr, float* __restrict__ da_dst) {
  const int64_t total = (int64_t)N * H;
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
    const int i = (int)(t / H), h = (int)(t - (int64_t)i * H);
    const int beg = __ldg(rowptr + i), deg = __ldg(rowptr + i + 1) - beg;
    const int64_t s0 = ((int64_t)beg + i) * H + h;
    const float ad = __ldg(a_dst + t);
    float sum = 0.f;
    for (int k = 0; k <= deg; ++k) {
      const bool removed = k < deg && __ldg(col + beg + k) == i;        // pre-existing self loop: removed by GATConv
      const float da = removed ? 0.f : dr[s0 + (int64_t)k * H];
      sum = fmaf(__ldg(alpha + s0 + (int64_t)k * H), da, sum);
    }
    float acc = 0.f;
    for (int k = 0; k <= deg; ++k) {
      const int j = k < deg ? __ldg(col + beg + k) : i;
      const bool removed = k < deg && j == i;
      const float da = removed ? 0.f : dr[s0 + (int64_t)k * H];
      const float fac = removed ? 0.f : ((__ldg(a_src + (int64_t)j * H + h) + ad > 0.f) ? 1.f : slope);
      const float d = __ldg(alpha + s0 + (int64_t)k * H) * (da - sum) * fac;
      dr[s0 + (int64_t)k * H] = d;
      acc = __fadd_rn(acc, d);
    }
    da_dst[t] = acc;
  }
}

}  // namespace emma
}  // namespace mgs
