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

// NT = number of 8-head column tiles (heads <= 8 NT).  W = floats per load: 2 (one 8-column slab per load) or 4 (rows 16-byte
// aligned and padded to a multiple of 4 floats: one 128-bit load feeds TWO slabs -- the first kernel, 64-bit loads only,
// was bound by L1 tag look-ups: every load instruction touches 8 different lines, and ran no faster than edge.cuh).
template <int NT, int W>
__global__ void __launch_bounds__(kThreads)
gat_bwd_edge_mma_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ xh, int64_t ld, int N, int H, int C,
                        const int* __restrict__ rowptr, const int* __restrict__ col, float* __restrict__ dalpha) {
  // head of every column, 0xff beyond H * C (padding columns match no head): one byte per column, read W at a time
  extern __shared__ __align__(16) unsigned char s_head[];
  const int HC = H * C;
  constexpr int kCols = 4 * W;                                  // columns one load step of a warp covers: 8 or 16
  constexpr int kSub = W / 2;                                   // 8-column MMA slabs per load step
  const int nsteps = (HC + kCols - 1) / kCols;
  for (int c = threadIdx.x; c < nsteps * kCols; c += kThreads) s_head[c] = c < HC ? (unsigned char)(c / C) : (unsigned char)0xff;
  __syncthreads();

  const int S = __ldg(rowptr + N) + N;                          // slots: every in-edge + one self loop per atom
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  const int warp_global = blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * kWarps;
  const int ntiles = (S + 15) / 16;
  const int nfull = HC / kCols;                                 // steps without padding / out-of-range columns

  for (int tile = warp_global; tile < ntiles; tile += nwarps) {
    // ---- my two slots: destination atom i (binary search over slot starts rowptr[i] + i) and source atom j; slots past the
    //      end compute on row 0 and are not stored ----
    const float* gp[2];
    const float* xp[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int s = tile * 16 + gid + 8 * r;
      int i = 0, j = 0;
      if (s < S) {
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
      gp[r] = g + (int64_t)i * ldg + W * tig;
      xp[r] = xh + (int64_t)j * ld + W * tig;
    }
    float acc[NT][4];
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[t][e] = 0.f;

    auto load = [&](int step, float (&gv)[2][W], float (&xv)[2][W]) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if constexpr (W == 4) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(gp[r] + kCols * step));
          const float4 b = __ldg(reinterpret_cast<const float4*>(xp[r] + kCols * step));
          gv[r][0] = a.x; gv[r][1] = a.y; gv[r][2] = a.z; gv[r][3] = a.w;
          xv[r][0] = b.x; xv[r][1] = b.y; xv[r][2] = b.z; xv[r][3] = b.w;
        } else {
          const float2 a = __ldg(reinterpret_cast<const float2*>(gp[r] + kCols * step));
          const float2 b = __ldg(reinterpret_cast<const float2*>(xp[r] + kCols * step));
          gv[r][0] = a.x; gv[r][1] = a.y;
          xv[r][0] = b.x; xv[r][1] = b.y;
        }
      }
    };
    // products of one step -> MMAs.  Slab q of the step: K index tig <-> my column 2 q, K index tig + 4 <-> my column 2 q + 1.
    // A fragment: a0 (row gid, k tig) a1 (row gid + 8, k tig) a2 (row gid, k tig + 4) a3 (row gid + 8, k tig + 4);
    // B fragment: b0 (k tig, n gid) b1 (k tig + 4, n gid) = indicator of "column's head == 8 t + gid" (exact in TF32)
    auto consume = [&](int step, const float (&gv)[2][W], const float (&xv)[2][W], bool tail) {
      int t_lo = 0, t_hi = 0;
      if (NT > 1) {
        const int first = s_head[kCols * step];
        int last = s_head[kCols * step + kCols - 1];
        if (last == 0xff) last = H - 1;
        t_lo = first >> 3;
        t_hi = last >> 3;
      }
      uint32_t heads;                                             // W head bytes of my columns
      if constexpr (W == 4) heads = *reinterpret_cast<const uint32_t*>(s_head + kCols * step + 4 * tig);
      else heads = *reinterpret_cast<const unsigned short*>(s_head + kCols * step + 2 * tig);
#pragma unroll
      for (int q = 0; q < kSub; ++q) {
        float p[4];
        p[0] = gv[0][2 * q] * xv[0][2 * q];
        p[1] = gv[1][2 * q] * xv[1][2 * q];
        p[2] = gv[0][2 * q + 1] * xv[0][2 * q + 1];
        p[3] = gv[1][2 * q + 1] * xv[1][2 * q + 1];
        const uint32_t ha = (heads >> (16 * q)) & 0xffu, hb = (heads >> (16 * q + 8)) & 0xffu;
        if (tail) {                                               // padding columns may hold anything (NaN x 0 = NaN)
          if (ha == 0xffu) { p[0] = 0.f; p[1] = 0.f; }
          if (hb == 0xffu) { p[2] = 0.f; p[3] = 0.f; }
        }
        uint32_t ahi[4], alo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          ahi[e] = tf32_hi(p[e]);
          alo[e] = __float_as_uint(p[e] - __uint_as_float(ahi[e]));
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          // a step's columns belong to one or two adjacent heads: only the head tile(s) they fall into get MMAs
          // (warp-uniform; H = 10: columns 0..279 feed tile 0 only, 280..349 tile 1 only)
          if (NT > 1 && (t < t_lo || t > t_hi)) continue;
          const uint32_t b0 = (ha == (uint32_t)(8 * t + gid)) ? 0x3f800000u : 0u;
          const uint32_t b1 = (hb == (uint32_t)(8 * t + gid)) ? 0x3f800000u : 0u;
          mma_tf32(acc[t], ahi, b0, b1);
          mma_tf32(acc[t], alo, b0, b1);
        }
      }
    };

    // two steps in flight (W = 4: four 128-bit loads per row pair); the tail step (columns up to the padded row end) is
    // read only when the rows are padded that far (W = 4: the launcher checks ld >= roundup(HC, 4)) or exact (W = 2: HC even)
    // software pipeline over PAIRS of steps: the loads of pair p + 1 are issued before pair p is consumed, so every thread
    // keeps 8-16 128-bit loads in flight while it computes (the first version waited for each pair: 2.0 TB/s, ncu:
    // long-scoreboard stalls 4.4 per issue at 44 % issue utilisation)
    // (two head tiles only: with one tile the extra registers cost more occupancy than the overlap returns -- stress
    // shape 251 -> 345 us -- and the plain pair loop below is used)
    constexpr bool kPipe = NT > 1;
    float gA[2][W], xA[2][W], gB[2][W], xB[2][W], gC[kPipe ? 2 : 1][W], xC[kPipe ? 2 : 1][W], gD[kPipe ? 2 : 1][W],
        xD[kPipe ? 2 : 1][W];
    int st = 0;
    if constexpr (!kPipe) {
      for (; st + 2 <= nfull; st += 2) {
        load(st, gA, xA);
        load(st + 1, gB, xB);
        consume(st, gA, xA, false);
        consume(st + 1, gB, xB, false);
      }
    } else if (nfull >= 2) {
      load(0, gA, xA);
      load(1, gB, xB);
      for (; st + 4 <= nfull; st += 4) {
        load(st + 2, gC, xC);
        load(st + 3, gD, xD);
        consume(st, gA, xA, false);
        consume(st + 1, gB, xB, false);
        if (st + 6 <= nfull) {
          load(st + 4, gA, xA);
          load(st + 5, gB, xB);
        }
        consume(st + 2, gC, xC, false);
        consume(st + 3, gD, xD, false);
      }
      if (st + 2 <= nfull) {               // one pair left, already loaded (the loop's refill condition mirrors this)
        consume(st, gA, xA, false);
        consume(st + 1, gB, xB, false);
        st += 2;
      }
    }
    for (; st < nsteps; ++st) {
      const int c = kCols * st + W * tig;
      if (c < HC) {
        load(st, gA, xA);
      } else {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int w = 0; w < W; ++w) { gA[r][w] = 0.f; xA[r][w] = 0.f; }
      }
      consume(st, gA, xA, true);
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
__global__ void __launch_bounds__(256, 4)   // 64 registers: 32 instead of 24 warps / SM (ncu: 80 registers, warps 32 % active, issue 33 %, DRAM 13 %)
gat_bwd_edge_softmax_kernel(const float* __restrict__ alpha, const float* __restrict__ a_src, const float* __restrict__ a_dst,
                            float slope, const int* __restrict__ rowptr, const int* __restrict__ col, int N, int H,
                            float* __restrict__ dr, float* __restrict__ da_dst) {
  constexpr int R = 8;                                          // slots of a row kept in registers (longer rows re-read)
  const int64_t total = (int64_t)N * H;
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
    const int i = (int)(t / H), h = (int)(t - (int64_t)i * H);
    const int beg = __ldg(rowptr + i), deg = __ldg(rowptr + i + 1) - beg;
    const int64_t s0 = ((int64_t)beg + i) * H + h;
    const float ad = __ldg(a_dst + t);
    float al[R], da[R], fc[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      al[k] = 0.f; da[k] = 0.f; fc[k] = 0.f;
      if (k <= deg) {
        const int j = k < deg ? __ldg(col + beg + k) : i;
        const bool removed = k < deg && j == i;                         // pre-existing self loop: removed by GATConv
        al[k] = __ldg(alpha + s0 + (int64_t)k * H);
        da[k] = removed ? 0.f : dr[s0 + (int64_t)k * H];
        fc[k] = removed ? 0.f : ((__ldg(a_src + (int64_t)j * H + h) + ad > 0.f) ? 1.f : slope);
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < R; ++k)
      if (k <= deg) sum = fmaf(al[k], da[k], sum);
    for (int k = R; k <= deg; ++k) {
      const bool removed = k < deg && __ldg(col + beg + k) == i;
      sum = fmaf(__ldg(alpha + s0 + (int64_t)k * H), removed ? 0.f : dr[s0 + (int64_t)k * H], sum);
    }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < R; ++k)
      if (k <= deg) {
        const float d = al[k] * (da[k] - sum) * fc[k];
        dr[s0 + (int64_t)k * H] = d;
        acc = __fadd_rn(acc, d);
      }
    for (int k = R; k <= deg; ++k) {
      const int j = k < deg ? __ldg(col + beg + k) : i;
      const bool removed = k < deg && j == i;
      const float dak = removed ? 0.f : dr[s0 + (int64_t)k * H];
      const float fac = removed ? 0.f : ((__ldg(a_src + (int64_t)j * H + h) + ad > 0.f) ? 1.f : slope);
      const float d = __ldg(alpha + s0 + (int64_t)k * H) * (dak - sum) * fac;
      dr[s0 + (int64_t)k * H] = d;
      acc = __fadd_rn(acc, d);
    }
    da_dst[t] = acc;
  }
}

}  // namespace emma
}  // namespace mgs
