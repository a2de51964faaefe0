// K2 backward, edge part, compile-time-shape variant (round 2): plain FMAs, four lanes per slot.
//
//   dalpha[slot, h] = < g_i[h, :], xh_j[h, :] >
//
// edge_mma.cuh does the per-head sums on the tensor core and pays for it in instructions: the fp32 products have to be split
// into TF32 hi + lo, the indicator fragments rebuilt for every slab -- 62 M warp instructions for the model1 batch, issue
// slots 38 % busy at 16 warps / SM with 6 cycles of long-scoreboard stall per issue (ncu, profiles/round2_ncu_edge.txt),
// DRAM at 29 %, L2 at 20 %: an instruction / latency problem, not a bandwidth problem.  With H and C known at compile time
// the head of every column is a constant after unrolling: a warp owns 16 consecutive slots, lanes (gid, tig) as in the MMA
// kernel (slot rows gid and gid + 8, columns 16 st + 4 tig .. + 3 of step st: one 128-bit load of g_i and one of xh_j per
// row and step, 64 contiguous bytes per slot and instruction), and a step whose 16 columns lie in ONE head is four FMAs
// into that head's accumulator; a step that straddles a head boundary (9 of 22 for C = 35) selects the multiplicand per
// column (col < boundary ? x : 0).  The four lanes of a slot are combined with two shuffles per head at the end.
// ~16 M warp instructions instead of 62 M.  CTAs own CONTIGUOUS runs of tiles: the destination-atom search of a tile is
// confined to the CTA's atom range (9 instead of 17 dependent look-ups) and a molecule's rows stay in one SM's L1.
// The softmax Jacobian runs as the second kernel of edge_mma.cuh.
#pragma once

#include <type_traits>

#include "common.cuh"

namespace mgs {
namespace efma {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

template <int I, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < E) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, E>(f);
  }
}

// largest i in [lo, hi) with rowptr[i] + i <= s
__device__ __forceinline__ int slot_atom(const int* __restrict__ rowptr, int lo, int hi, int s) {
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(rowptr + mid) + mid <= s) lo = mid; else hi = mid;
  }
  return lo;
}

// D = steps of loads in flight per thread (D x 4 128-bit loads).  Rows 16-byte aligned, ld / ldg multiples of 4 floats and
// >= roundup4(H C) (the launcher checks).
template <int H, int C, int D>
__global__ void __launch_bounds__(kThreads, 2)
gat_bwd_edge_fma_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ xh, int64_t ld, int N,
                        const int* __restrict__ rowptr, const int* __restrict__ col, float* __restrict__ dalpha) {
  constexpr int HC = H * C;
  constexpr int NS = (HC + 15) / 16;                            // 16-column steps
  constexpr int kPad4 = (HC + 3) & ~3;
  static_assert(C >= 16, "a 16-column step may span two heads at most");
  static_assert(D >= 1 && D <= NS, "pipeline depth");
  __shared__ int s_range[2];

  const int S = __ldg(rowptr + N) + N;                          // slots: every in-edge + one self loop per atom
  const int ntiles = (S + 15) / 16;
  const int per_cta = (ntiles + gridDim.x - 1) / gridDim.x;
  const int t0 = blockIdx.x * per_cta, t1 = min(ntiles, t0 + per_cta);
  if (t0 >= t1) return;
  if (threadIdx.x < 2) {
    const int s = threadIdx.x == 0 ? t0 * 16 : min(S - 1, t1 * 16 - 1);
    s_range[threadIdx.x] = slot_atom(rowptr, 0, N, s);
  }
  __syncthreads();
  const int i_lo = s_range[0], i_hi = s_range[1] + 1;
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;

  for (int tile = t0 + (threadIdx.x >> 5); tile < t1; tile += kWarps) {
    const float* gp[2];
    const float* xp[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int s = tile * 16 + gid + 8 * r;
      int i = i_lo, j = i_lo;                                   // slots past the end compute on a valid row, not stored
      if (s < S) {
        i = slot_atom(rowptr, i_lo, i_hi, s);
        const int beg = __ldg(rowptr + i);
        const int k = s - (beg + i);
        j = (k == __ldg(rowptr + i + 1) - beg) ? i : __ldg(col + beg + k);     // last slot of a row: the self loop
      }
      gp[r] = g + (int64_t)i * ldg + 4 * tig;
      xp[r] = xh + (int64_t)j * ld + 4 * tig;
    }
    float acc[2][H];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int h = 0; h < H; ++h) acc[r][h] = 0.f;
    float4 bg[D][2], bx[D][2];

    auto load = [&](auto stc, auto bc) {
      constexpr int st = decltype(stc)::value, b = decltype(bc)::value;
      // the last step may reach past the padded row end (only when 16 NS > roundup4(HC)): those groups are not read
      const bool in_row = 16 * (st + 1) <= kPad4 || 16 * st + 4 * tig < kPad4;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (in_row) {
          bg[b][r] = __ldg(reinterpret_cast<const float4*>(gp[r] + 16 * st));
          bx[b][r] = __ldg(reinterpret_cast<const float4*>(xp[r] + 16 * st));
        } else {
          bg[b][r] = make_float4(0.f, 0.f, 0.f, 0.f);
          bx[b][r] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    static_for<0, D>([&](auto stc) { load(stc, stc); });
    static_for<0, NS>([&](auto stc) {
      constexpr int st = decltype(stc)::value, b = st % D;
      constexpr int c0 = 16 * st;
      constexpr int hA = c0 / C;
      constexpr int hB = (c0 + 15 < HC ? c0 + 15 : HC - 1) / C;
      constexpr bool tail = c0 + 16 > HC;                       // columns beyond H C: padding, may hold anything
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float gv[4] = {bg[b][r].x, bg[b][r].y, bg[b][r].z, bg[b][r].w};
        float xv[4] = {bx[b][r].x, bx[b][r].y, bx[b][r].z, bx[b][r].w};
        if constexpr (tail) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool valid = c0 + 4 * tig + u < HC;
            gv[u] = valid ? gv[u] : 0.f;
            xv[u] = valid ? xv[u] : 0.f;
          }
        }
        if constexpr (hA == hB) {
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[r][hA] = fmaf(gv[u], xv[u], acc[r][hA]);
        } else {
          constexpr int bnd = hB * C;                             // first column of head hB
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool in_a = c0 + 4 * tig + u < bnd;
            acc[r][hA] = fmaf(gv[u], in_a ? xv[u] : 0.f, acc[r][hA]);
            acc[r][hB] = fmaf(gv[u], in_a ? 0.f : xv[u], acc[r][hB]);
          }
        }
      }
      if constexpr (st + D < NS) load(std::integral_constant<int, st + D>{}, std::integral_constant<int, b>{});
    });
    // the four lanes of a slot -> lane tig stores heads tig, tig + 4, ...
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int s = tile * 16 + gid + 8 * r;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        float v = acc[r][h];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if ((h & 3) == tig && s < S) dalpha[(int64_t)s * H + h] = v;
      }
    }
  }
}

}  // namespace efma
}  // namespace mgs
