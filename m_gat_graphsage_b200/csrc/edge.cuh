// K2 backward, edge part: d alpha / softmax Jacobian / leaky_relu' per (slot, head)  
//   dalpha_e[h] = < g_i[h,:], xh_j[h,:] > * w_e * mask_e          e = (j -> i), slots of i = in-edges + self loop
//   dr_e[h]     = alpha_e[h] * (dalpha_e[h] - sum_k alpha_k dalpha_k) * leaky'(a_src[j,h] + a_dst[i,h])
//   da_dst[i,h] = sum_e dr_e[h]
//
// Same traffic as the forward aggregation (gather every xh_j once per in-edge, read g_i once), but the
// arithmetic is a per-HEAD dot product: the coalesced mapping of stream.cuh (lane l owns 8-byte chunks l,
// l+32, ...) scatters every head over all 32 lanes, while a head-aligned mapping (P = 32/H lanes per head)
// needs scalar loads with a 48-byte lane stride -- measured on B200: 12 wavefronts per LDG, 144 per gathered
// row, L1 88 % hit rate but the LSU pipe saturated (0.28 ms, 23 % of HBM peak).  Here both mappings are
// used for what they are good at.  A warp owns a contiguous range of destination rows; the CSR column slice
// of 32 rows at a time is copied to shared memory in one coalesced pass, then per row:
//   A. g_i and up to G = 4 gathered xh_j rows are requested with coalesced 64/128-bit loads (lane l owns
//      chunks l, l+32, ...), together with the per-slot scalars (alpha, mask, a_src) of the head leaders;
//   B. the rows are parked in a per-warp staging buffer (coalesced STS);
//   C. lane (head h, part p) reads ITS channels of head h back with conflict-free LDS (channel order and
//      rotation chosen on the host by brute force over bank conflicts), Q FMAs against its slice of g_i in
//      registers, a P-lane shuffle tree, and the leader finishes the softmax Jacobian through a small
//      shared-memory ring when the row ends (rows with more than 8 slots spill d alpha to the dr buffer).
// B200, model1 batch: 0.28 ms -> 0.23 ms.  ncu: still latency bound (16 warps / SM, ~900 -> ~600 dependent
// instructions per row); issuing the next row's gathers before phase C made it slower (the loads' registers
// spill); the next step is a TMA-fed row ring so that no registers sit between HBM and shared memory.
#pragma once

#include "common.cuh"
#include "stream.cuh"

namespace mgs {
namespace edge {

using stream::Row;

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kListMax = 256;
constexpr int kRing = 8;
constexpr int G = 4;

struct Args {
  const float* g; int64_t ldg;
  const float* xh; int64_t ld;
  int N, H, C, P, Q, chunks;
  int interleaved, rot_a, rot_b;          // channel order of the head-aligned read-back (host-chosen)
  const float* alpha; const float* amask; const float* a_src; const float* a_dst;
  float slope;
  const int* rowptr; const int* col; const int* perm;
  const float* ew;
  float* dr; float* da_dst; float* dew;
};

// one staged row: every chunk the coalesced mapping can touch + a zero pad read by unused channel steps
template <int V, int ITERS> struct Stage {
  static constexpr int kFloats = 32 * V * ITERS + 4;
  static constexpr int kZero = 32 * V * ITERS;
};

template <int V, int ITERS>
inline size_t smem_bytes() {
  return (size_t)kWarps * ((G + 1) * Stage<V, ITERS>::kFloats * 4 + kRing * 3 * 32 * 4 + kListMax * 4 * 3);
}

// EXTRA = explainer / attention-dropout instantiation (edge weights, their gradient, alpha mask); the training
// and inference steps run the plain one.  All element indices are 32-bit (the launcher checks the sizes).
template <int V, int ITERS, int QT, bool EXTRA>
__global__ void __launch_bounds__(kThreads, 2) gat_bwd_edge_kernel(const Args a) {
  constexpr int SF = Stage<V, ITERS>::kFloats;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* stage = reinterpret_cast<float*>(smem_raw) + (size_t)warp * (G + 1) * SF;       // [0] = g_i, [1..G] = xh_j
  float* ring = reinterpret_cast<float*>(smem_raw) + (size_t)kWarps * (G + 1) * SF + (size_t)warp * kRing * 3 * 32;
  int* lists = reinterpret_cast<int*>(reinterpret_cast<float*>(smem_raw) + (size_t)kWarps * ((G + 1) * SF + kRing * 3 * 32));
  int* s_col = lists + (size_t)warp * kListMax;
  float* s_w = reinterpret_cast<float*>(lists + (size_t)(kWarps + warp) * kListMax);
  int* s_eid = lists + (size_t)(2 * kWarps + warp) * kListMax;

  const int nwarps = gridDim.x * kWarps;
  const int rows_per_warp = (a.N + nwarps - 1) / nwarps;
  const int row_lo = (blockIdx.x * kWarps + warp) * rows_per_warp;
  const int row_hi = min(a.N, row_lo + rows_per_warp);
  const unsigned H = (unsigned)a.H;
  const bool weighted = EXTRA && a.ew != nullptr;
  const bool want_dew = EXTRA && a.dew != nullptr;
  const bool masked = EXTRA && a.amask != nullptr;
  const float slope = a.slope;
  const unsigned ldx = (unsigned)a.ld, ldg = (unsigned)a.ldg;

  // coalesced mapping (phases A / B)
  const int lane_off = lane * V;
  const bool tail_a = lane + 32 * (ITERS - 2) < a.chunks;
  const bool tail_b = lane + 32 * (ITERS - 1) < a.chunks;
  auto act = [&](int t) { return t < ITERS - 2 ? true : (t == ITERS - 2 ? tail_a : tail_b); };
  const float* xh_lane = a.xh + lane_off;
  const float* g_lane = a.g + lane_off;
  float* stage_lane = stage + lane_off;

  // head-aligned mapping (phase C): shared-memory address of this lane's channels in staged row 0; unused
  // steps read the zero pad
  const int P = a.P;
  const unsigned hl = lane / P, pl = lane - hl * P;
  const bool active = hl < H;
  const bool leader = active && pl == 0;
  uint32_t oq[QT];
  {
    int nq = 0;
    if (active) nq = a.interleaved ? ((int)pl < a.C ? (a.C - pl + P - 1) / P : 0) : max(0, min(a.Q, a.C - (int)pl * a.Q));
    const int rot = nq > 0 ? (a.rot_a * hl + a.rot_b * pl) % nq : 0;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(stage);
#pragma unroll
    for (int q = 0; q < QT; ++q) {
      int idx = Stage<V, ITERS>::kZero;
      if (q < nq) {
        int qq = q + rot;
        if (qq >= nq) qq -= nq;
        idx = hl * a.C + (a.interleaved ? pl + P * qq : pl * a.Q + qq);
      }
      oq[q] = base + idx * 4;
    }
  }
  auto lds = [](uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
  };
  if (lane < 4)
    for (int k = 0; k <= G; ++k) stage[k * SF + Stage<V, ITERS>::kZero + lane] = 0.f;
  // tree sum over the P consecutive lanes of a head (any P); valid in the leader
  int Pc = 1;
  while (Pc < P) Pc <<= 1;
  const bool st16 = Pc > 16, st8 = Pc > 8, st4 = Pc > 4, st2 = Pc > 2, st1 = Pc > 1;
  const bool in16 = (int)pl + 16 < P, in8 = (int)pl + 8 < P, in4 = (int)pl + 4 < P, in2 = (int)pl + 2 < P, in1 = (int)pl + 1 < P;
  auto group_sum = [&](float t) {
    float u;
    if (st16) { u = __shfl_down_sync(0xffffffffu, t, 16); if (in16) t += u; }
    if (st8) { u = __shfl_down_sync(0xffffffffu, t, 8); if (in8) t += u; }
    if (st4) { u = __shfl_down_sync(0xffffffffu, t, 4); if (in4) t += u; }
    if (st2) { u = __shfl_down_sync(0xffffffffu, t, 2); if (in2) t += u; }
    if (st1) { u = __shfl_down_sync(0xffffffffu, t, 1); if (in1) t += u; }
    return t;
  };
  float* ring_lane = ring + lane;

  for (int i0 = row_lo; i0 < row_hi; i0 += 32) {
    const int nrows = min(32, row_hi - i0);
    // ---- the block's slice of the CSR column array (contiguous) goes to shared memory in one coalesced pass
    int beg = 0, elen = 0;
    if (lane < nrows) {
      beg = __ldg(a.rowptr + i0 + lane);
      elen = __ldg(a.rowptr + i0 + lane + 1) - beg;
    }
    const int beg0 = __shfl_sync(0xffffffffu, beg, 0);
    const int cnt = __shfl_sync(0xffffffffu, beg + elen, nrows - 1) - beg0;
    const bool listed = cnt <= kListMax;                    // else (very high degrees): read col from global
    __syncwarp();
    if (listed) {
      for (int t = lane; t < cnt; t += 32) {
        s_col[t] = __ldg(a.col + beg0 + t);
        if (weighted || want_dew) {
          const int e = __ldg(a.perm + beg0 + t);
          s_eid[t] = e;
          if (weighted) s_w[t] = __ldg(a.ew + e);
        }
      }
    }
    __syncwarp();

    for (int r = 0; r < nrows; ++r) {
      const unsigned i = (unsigned)(i0 + r);
      const int beg_r = __shfl_sync(0xffffffffu, beg, r);
      const int nsl = __shfl_sync(0xffffffffu, elen, r) + 1;          // in-edges + self loop (last)
      const unsigned slotH = ((unsigned)beg_r + i) * H + hl;            // (first slot of the row, my head)
      const int* colp = listed ? nullptr : a.col + beg_r;
      const int lbase = beg_r - beg0;
      float gq[QT];
      float adst = 0.f, sum = 0.f;

      for (int s0 = 0; s0 < nsl; s0 += G) {
        int j[G];
        float al[G], as[G], mk[G], w[G];
        Row<V> v[G][ITERS], vg[ITERS];
        // ---- A: coalesced gathers + per-slot scalars, everything in flight together ----
        if (s0 == 0) {
          const float* rowp = g_lane + i * ldg;
#pragma unroll
          for (int it = 0; it < ITERS; ++it)
            if (act(it)) vg[it] = Row<V>::load(rowp + 32 * V * it);
          if (leader) adst = __ldg(a.a_dst + (i * H + hl));
        }
#pragma unroll
        for (int k = 0; k < G; ++k) {
          j[k] = -2;                                        // -2: no slot, -1: slot of a removed self loop
          al[k] = 0.f; as[k] = 0.f; mk[k] = 1.f; w[k] = 1.f;
          const int s = s0 + k;
          if (s < nsl) {
            if (s == nsl - 1) {
              j[k] = (int)i;
            } else {
              j[k] = listed ? s_col[lbase + s] : __ldg(colp + s);
              if (weighted) w[k] = listed ? s_w[lbase + s] : __ldg(a.ew + __ldg(a.perm + beg_r + s));
              if (j[k] == (int)i) j[k] = -1;                // pre-existing self loop: removed by GATConv
            }
            if (j[k] >= 0) {
              const float* rowp = xh_lane + (unsigned)j[k] * ldx;
#pragma unroll
              for (int it = 0; it < ITERS; ++it)
                if (act(it)) v[k][it] = Row<V>::load(rowp + 32 * V * it);
              if (leader) as[k] = __ldg(a.a_src + ((unsigned)j[k] * H + hl));
            }
            if (leader) {
              al[k] = __ldg(a.alpha + (slotH + (unsigned)s * H));
              if (masked) mk[k] = __ldg(a.amask + (slotH + (unsigned)s * H));
            }
          }
        }
        // ---- B: park the rows in the staging buffer ----
        __syncwarp();
        if (s0 == 0) {
#pragma unroll
          for (int it = 0; it < ITERS; ++it)
            if (act(it)) vg[it].store(stage_lane + 32 * V * it);
        }
#pragma unroll
        for (int k = 0; k < G; ++k) {
          if (j[k] >= 0) {
#pragma unroll
            for (int it = 0; it < ITERS; ++it)
              if (act(it)) v[k][it].store(stage_lane + (k + 1) * SF + 32 * V * it);
          }
        }
        __syncwarp();
        // ---- C: head-aligned dot products ----
        if (s0 == 0) {
#pragma unroll
          for (int q = 0; q < QT; ++q) gq[q] = lds(oq[q]);
        }
#pragma unroll
        for (int k = 0; k < G; ++k) {
          if (j[k] >= -1) {
            const int kin = s0 + k;
            float d0 = 0.f, d1 = 0.f;                       // two chains: half the dependent-FMA latency
            if (j[k] >= 0) {
#pragma unroll
              for (int q = 0; q < QT; q += 2) {
                d0 = fmaf(gq[q], lds(oq[q] + (k + 1) * SF * 4), d0);
                if (q + 1 < QT) d1 = fmaf(gq[q + 1], lds(oq[q + 1] + (k + 1) * SF * 4), d1);
              }
            }
            const float dot = group_sum(d0 + d1);
            if (want_dew) {                                 // d w_e = sum_h alpha_used * <g_i, xh_j>
              float c = (leader && j[k] >= 0) ? al[k] * mk[k] * dot : 0.f;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
              if (lane == 0 && kin < nsl - 1)
                a.dew[listed ? s_eid[lbase + kin] : __ldg(a.perm + beg_r + kin)] = c;
            }
            if (leader) {
              float da = (j[k] >= 0) ? dot : 0.f;
              if (EXTRA) da = da * w[k] * mk[k];
              const float raw = as[k] + adst;
              const float fac = (j[k] >= 0) ? (raw > 0.f ? 1.f : slope) : 0.f;
              sum = fmaf(al[k], da, sum);
              if (kin < kRing) {
                float* rp = ring_lane + kin * 96;
                rp[0] = al[k]; rp[32] = da; rp[64] = fac;
              } else {
                a.dr[slotH + (unsigned)kin * H] = da;
              }
            }
          }
        }
      }
      // ---- row complete: softmax Jacobian, leaky_relu', da_dst ----
      if (leader) {
        float acc = 0.f;
        float* drp = a.dr + slotH;
        for (int s = 0; s < nsl; ++s) {
          float al_s, da_s, fac_s;
          if (s < kRing) {
            const float* rp = ring_lane + s * 96;
            al_s = rp[0]; da_s = rp[32]; fac_s = rp[64];
          } else {                                          // long row: recompute from the spilled d alpha
            const int jj = (s == nsl - 1) ? (int)i : __ldg(a.col + beg_r + s);
            al_s = __ldg(a.alpha + (slotH + (unsigned)s * H));
            da_s = *drp;
            const float raw = __ldg(a.a_src + ((unsigned)jj * H + hl)) + adst;
            fac_s = (s != nsl - 1 && jj == (int)i) ? 0.f : (raw > 0.f ? 1.f : slope);
          }
          const float d = al_s * (da_s - sum) * fac_s;
          *drp = d;
          drp += H;
          acc = __fadd_rn(acc, d);
        }
        a.da_dst[i * H + hl] = acc;
      }
    }
  }
}

// Channel order of the read-back: contiguous (lane p of a head owns channels [pQ, pQ+Q)) or interleaved
// (channels p, p+P, ...), each lane starting at a rotated position; pick the variant with the fewest
// shared-memory wavefronts for this (H, C).  Cheap (<= 2 * 12 * 12 * 32 * Q steps); cached per thread.
inline void pick_order(int H, int C, int P, int Q, int* interleaved, int* rot_a, int* rot_b) {
  thread_local int cH = -1, cC = -1, cI = 0, cA = 0, cB = 0;
  if (cH == H && cC == C) { *interleaved = cI; *rot_a = cA; *rot_b = cB; return; }
  long best = -1;
  const int R = Q < 12 ? Q : 12;
  for (int il = 0; il < 2; ++il)
    for (int ra = 0; ra < R; ++ra)
      for (int rb = 0; rb < R; ++rb) {
        long tot = 0;
        for (int q = 0; q < Q; ++q) {
          int cnt[32] = {0};
          int addrs[32][32];
          int worst = 0;
          for (int lane = 0; lane < 32; ++lane) {
            const int hl = lane / P, pl = lane - hl * P;
            if (hl >= H) continue;
            int nq = il ? (pl < C ? (C - pl + P - 1) / P : 0) : (C - pl * Q < 0 ? 0 : (C - pl * Q < Q ? C - pl * Q : Q));
            if (q >= nq) continue;
            int qq = q + (ra * hl + rb * pl) % nq;
            if (qq >= nq) qq -= nq;
            const int addr = hl * C + (il ? pl + P * qq : pl * Q + qq);
            const int b = addr & 31;
            bool dup = false;
            for (int z = 0; z < cnt[b]; ++z) dup |= addrs[b][z] == addr;
            if (!dup) addrs[b][cnt[b]++] = addr;
            if (cnt[b] > worst) worst = cnt[b];
          }
          tot += worst;
        }
        if (best < 0 || tot < best) { best = tot; cI = il; cA = ra; cB = rb; }
      }
  cH = H; cC = C;
  *interleaved = cI; *rot_a = cA; *rot_b = cB;
}

}  // namespace edge
}  // namespace mgs
