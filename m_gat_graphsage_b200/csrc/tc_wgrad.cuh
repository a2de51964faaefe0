// K4, weight gradient on the TMA-fed tcgen05 path:  dw[o][i] = sum_r g[r][o] * a[r][i]   (contraction over the atoms r).
//
// Both operands are activations whose CONTRACTION index is the slow one in memory ("MN-major" in UMMA terms).  The cp.async
// kernel of tc_linear.cuh transposes them through registers into K-major tiles (16 producer warps, L1TEX 90 % busy, tensor
// pipe 48 % active: 0.22 ms per SAGE weight gradient against 0.165 ms for the same flops in the forward kernel).  Here
// nothing is transposed by a thread:
//   * A = g^T goes through TENSOR MEMORY (TS form of tcgen05.mma, tc_tma.cuh): one tensor-map TMA drops the raw
//     [16 atoms x 128 channels] tile into shared memory as it lies in HBM; converter thread m reads COLUMN m of it (16
//     conflict-free LDS.32: consecutive threads, consecutive words) and stores hi (truncated) and lo as 16 + 16 columns of
//     TMEM lane m -- the transposition is the addressing of tcgen05.st;
//   * B = a stays in shared memory as an MN-major operand (instruction descriptor bit 16).  For 32-bit elements the tensor
//     core accepts exactly one MN-major layout, SWIZZLE_128B_BASE32B (layout type 1; with the K-major SWIZZLE_64B type the
//     MMA silently produces zeros -- measured): rows of 32 channels (128 bytes), 32-byte chunks XORed with (row & 3), the
//     canonical  Swizzle<2,5,2> o ((8,n),(4,k)) : ((1,LBO),(8,SBO))  in uint128 units (cute::UMMA::Layout_MN_SW128_32B_Atom).
//     That is what a tensor-map TMA with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B and a box of [16 atoms x 32 channels] writes:
//     channels 32 j .. 32 j + 31 at byte offset 2048 j (LBO), atoms 4 q .. 4 q + 3 at 512 q (SBO); six boxes per K block for
//     a 176-wide tile (the last one half used).  The raw tile is the hi operand (the tensor core truncates), lo is computed
//     element-wise by the converters exactly as in the forward kernel (layout-agnostic);
//   * the contraction is split over the CTAs (one (tile, split) per CTA, all of them resident: 6 tiles x 24 splits for the
//     SAGE layer), partial tiles summed in split order by splitk_reduce_kernel: deterministic.
// Same three products and accumulator pairing as every K4 kernel: hi*hi -> main, lo*hi + hi*lo -> correction.
// Measured and dropped: a whole-span cp.async.bulk.prefetch.L2 of a split's rows ahead of the boxes (no change at 130 k
// atoms, 20 % slower at 4096 where everything is L2-resident anyway).
#pragma once

#include "tc_tma.cuh"

namespace mgs {
namespace tma {

constexpr int kBoxN = 32;                                   // channels per B box = one 128-byte swizzle row
constexpr int kBoxBytes = kBoxN * 4 * BK;                   // 2048: [16 atoms][32 channels]

// MN-major, SWIZZLE_128B_BASE32B operand: LBO = distance between 32-channel groups, SBO = distance between 4-atom groups
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(kBoxBytes >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

// shared-memory plan of the weight-gradient kernel: tensor-memory plan of Cfg<BN, true>, B tiles in whole 32-channel boxes
template <int BN> struct WCfg {
  using Base = Cfg<BN, true>;
  static constexpr int kBoxes = (BN + kBoxN - 1) / kBoxN;
  static constexpr int kABytes = BM * kRowBytes;             // 8192: [16 atoms][128 channels] of g
  static constexpr int kBBytes = kBoxes * kBoxBytes;
  static constexpr int kRawBytes = kABytes + kBBytes;
  static constexpr int kLoBytes = kBBytes;
  static constexpr int kLoStages = 4;
  static constexpr int kBudget = 227 * 1024 - 1024 - 512;
  static constexpr int kRawFit = (kBudget - kLoStages * kLoBytes) / kRawBytes;
  static constexpr int kRawStages = kRawFit > 10 ? 10 : kRawFit;
  static_assert(kRawStages >= 4, "raw ring too short");
  static constexpr int kSmemBytes = kRawStages * kRawBytes + kLoStages * kLoBytes + 1024 + 512;
  static constexpr int kConvWarps = Base::kConvWarps, kGroups = Base::kGroups, kEpiWarps = Base::kEpiWarps;
  static constexpr int kTmemCols = Base::kTmemCols, kCorrCol = Base::kCorrCol;
  __host__ __device__ static constexpr uint32_t a_col(int l) { return Base::a_col(l); }
};
__host__ __device__ constexpr uint32_t make_idesc_bmn(int n) { return make_idesc(n) | (1u << 16); }

// map_g: dims {Nout, M}, box {128, 16}, no swizzle.   map_a: dims {K, M}, box {32, 16}, SWIZZLE_128B_ATOM_32B.   Zero fill.
// Work item w = (split z, tile mn): atoms [z * nb_split * 16, ...), output rows m0 .. m0 + 127, columns n0 .. n0 + BN - 1;
// written to c + z * split_stride (ldc).  A split without atoms writes zeros.
template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tma_wgrad_kernel(const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_a, int M /* atoms */,
                      int Nout, int K, float* __restrict__ c, int64_t ldc, int splits, int64_t split_stride, int dbg) {
  using C = WCfg<BN>;
  static_assert(kThreads == (C::kConvWarps + C::kEpiWarps + 2) * 32, "warp roles");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int R = C::kRawStages, L = C::kLoStages;
  uint8_t* raw_ring = smem;                                   // R slots of [g raw: 16 x 128 | a raw: BN / 16 boxes]
  uint8_t* lo_ring = smem + R * C::kRawBytes;                 // L slots of [a lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + R * C::kRawBytes + L * C::kLoBytes);
  auto bar_raw_full = [&](int i) { return smem_u32(bars + i); };
  auto bar_raw_free = [&](int i) { return smem_u32(bars + R + i); };
  auto bar_lo_full = [&](int i) { return smem_u32(bars + 2 * R + i); };
  auto bar_lo_free = [&](int i) { return smem_u32(bars + 2 * R + L + i); };
  const uint32_t bar_acc_full = smem_u32(bars + 2 * R + 2 * L);
  const uint32_t bar_tmem_free = smem_u32(bars + 2 * R + 2 * L + 1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * R + 2 * L + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = (M + BK - 1) / BK;                           // K blocks = groups of 16 atoms
  const int ntn = (K + BN - 1) / BN, ntm = (Nout + BM - 1) / BM;
  const int nmn = ntn * ntm;
  const int ntiles = nmn * splits;
  const int nb_split = (nb + splits - 1) / splits;
  constexpr int kMmaWarp = C::kConvWarps + C::kEpiWarps;

  if (threadIdx.x == 0) {
    for (int i = 0; i < R; ++i) {
      mbar_init(bar_raw_full(i), 1);
      mbar_init(bar_raw_free(i), 1);
    }
    for (int i = 0; i < L; ++i) {
      mbar_init(bar_lo_full(i), C::kConvWarps / C::kGroups);
      mbar_init(bar_lo_free(i), 1);
    }
    mbar_init(bar_acc_full, 1);
    mbar_init(bar_tmem_free, C::kEpiWarps);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < C::kConvWarps) {
    // ================= converters: g column -> tensor memory (hi | lo), a tile -> lo slot ================================
    constexpr int kVec = C::kLoBytes / 16;
    constexpr int kSkip = C::kABytes / 16;
    constexpr int kGroupThreads = (C::kConvWarps / C::kGroups) * 32;
    constexpr int kPer = (kVec + kGroupThreads - 1) / kGroupThreads;
    const int grp = warp / (C::kConvWarps / C::kGroups), gt = threadIdx.x % kGroupThreads;
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int z = tile / nmn;
      const int it_lo = z * nb_split, it_hi = min(nb, it_lo + nb_split);
      for (int it = it_lo; it < it_hi; ++it, ++g) {
        if ((int)((g % R) % C::kGroups) != grp) continue;                // a raw slot belongs to one group (tc_tma.cuh)
        mbar_wait(bar_raw_full((int)(g % R)), (g / R) & 1u);
        const int l = (int)(g % L);
        mbar_wait(bar_lo_free(l), ((g / L) & 1u) ^ 1u);
        const uint8_t* slot = raw_ring + (g % R) * C::kRawBytes;
        const uint4* raw = reinterpret_cast<const uint4*>(slot) + kSkip;
        uint4* lo = reinterpret_cast<uint4*>(lo_ring + l * C::kLoBytes);
        if (!(dbg & 8)) {
          tc_fence_after();
          const int row = (warp & 3) * 32 + lane;                        // output channel of the tile = TMEM lane
          const uint32_t* col = reinterpret_cast<const uint32_t*>(slot) + row;
          uint32_t hi[16], lw[16];
#pragma unroll
          for (int k = 0; k < BK; ++k) {
            const uint32_t v = col[k * BM];
            hi[k] = v & 0xffffe000u;
            lw[k] = lo_word(v);
          }
          const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + C::a_col(l);
          tmem_st16(ta, hi);
          tmem_st16(ta + 16, lw);
          uint4 v[kPer];
#pragma unroll
          for (int j = 0; j < kPer; ++j)
            if (gt + j * kGroupThreads < kVec) v[j] = raw[gt + j * kGroupThreads];
#pragma unroll
          for (int j = 0; j < kPer; ++j)
            if (gt + j * kGroupThreads < kVec)
              lo[gt + j * kGroupThreads] = make_uint4(lo_word(v[j].x), lo_word(v[j].y), lo_word(v[j].z), lo_word(v[j].w));
        }
        fence_proxy_async();
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_lo_full(l));
      }
    }
  } else if (warp < kMmaWarp) {
    // ================= epilogue warps: TMEM -> registers (accumulator released) -> global partial tile ===================
    const int ew = warp - C::kConvWarps;
    const int q = warp & 3;
    const int half = ew >> 2;
    const int row_l = q * 32 + lane;
    constexpr int kChunksW = BN / 16;
    constexpr int kPass = kChunksW > 11 ? (kChunksW + 1) / 2 : kChunksW;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(c) | (uintptr_t)(ldc * 4) | (uintptr_t)(split_stride * 4)) & 15u) == 0;
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
      const int z = tile / nmn, mn = tile - z * nmn;
      const int m0 = (mn / ntn) * BM, n0 = (mn % ntn) * BN;
      const bool any = min(nb, z * nb_split + nb_split) > z * nb_split;
      mbar_wait(bar_acc_full, tl & 1u);
      tc_fence_after();
      float* crow = c + (int64_t)z * split_stride + (int64_t)(m0 + row_l) * ldc + n0;
      const bool row_ok = m0 + row_l < Nout;
      for (int p0 = 0; p0 < kChunksW; p0 += kPass) {
        float acc[kPass][8];
#pragma unroll
        for (int j = 0; j < kPass; ++j) {
          if (p0 + j < kChunksW) {
            const uint32_t col = (uint32_t)(8 * (half * kChunksW + p0 + j));
            uint32_t rm[8], rc[8];
            tmem_ld8_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + col, rm);
            tmem_ld8_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)C::kCorrCol + col, rc);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[j][u] = any ? __uint_as_float(rm[u]) + __uint_as_float(rc[u]) : 0.f;
          }
        }
        if (p0 + kPass >= kChunksW) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tmem_free);
        }
#pragma unroll
        for (int j = 0; j < kPass; ++j) {
          if (p0 + j < kChunksW) {
            const int nl = 8 * (half * kChunksW + p0 + j);
            if (row_ok && n0 + nl < K) {
              if (vec_ok && n0 + nl + 8 <= K) {
                *reinterpret_cast<float4*>(crow + nl) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
                *reinterpret_cast<float4*>(crow + nl + 4) = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
              } else {
#pragma unroll
                for (int u = 0; u < 8; ++u)
                  if (n0 + nl + u < K) crow[nl + u] = acc[j][u];
              }
            }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0) {
      // ================= MMA issuer =================
      constexpr uint32_t idesc = make_idesc_bmn(BN);
      uint32_t g = 0, tl = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
        const int z = tile / nmn;
        const int it_lo = z * nb_split, it_hi = min(nb, it_lo + nb_split);
        mbar_wait(bar_tmem_free, (tl & 1u) ^ 1u);
        tc_fence_after();
        for (int it = it_lo; it < it_hi; ++it, ++g) {
          const int r = (int)(g % R), l = (int)(g % L);
          mbar_wait(bar_lo_full(l), (g / L) & 1u);
          tc_fence_after();
          const uint32_t sr = smem_u32(raw_ring + r * C::kRawBytes), sl = smem_u32(lo_ring + l * C::kLoBytes);
          const uint64_t b_hi = make_desc_mn(sr + C::kABytes);
          const uint64_t b_lo = make_desc_mn(sl);
          const uint32_t ta = tmem_base + C::a_col(l);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            if (dbg & 4) break;
            const uint64_t adv = (uint64_t)(k * 1024 >> 4);              // the next 8 atoms (two 4-atom groups) of every box
            const uint32_t first = (it != it_lo || k != 0) ? 1u : 0u;
            umma_tf32_ts(tmem_base, ta + 8 * k, b_hi + adv, idesc, first);
            umma_tf32_ts(tmem_base + C::kCorrCol, ta + 16 + 8 * k, b_hi + adv, idesc, first);
            umma_tf32_ts(tmem_base + C::kCorrCol, ta + 8 * k, b_lo + adv, idesc, 1);
          }
          umma_commit(bar_raw_free(r));
          umma_commit(bar_lo_free(l));
        }
        umma_commit(bar_acc_full);
      }
    }
  } else if (lane == 0) {
    // ================= loader: one g box and ceil(BN / 32) a boxes per K block =================================================
    tma_prefetch_desc(&map_g);
    tma_prefetch_desc(&map_a);
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int z = tile / nmn, mn = tile - z * nmn;
      const int m0 = (mn / ntn) * BM, n0 = (mn % ntn) * BN;
      const int it_lo = z * nb_split, it_hi = min(nb, it_lo + nb_split);
      for (int it = it_lo; it < it_hi; ++it, ++g) {
        const int r = (int)(g % R);
        mbar_wait(bar_raw_free(r), ((g / R) & 1u) ^ 1u);
        const uint32_t full = bar_raw_full(r);
        const uint32_t dst = smem_u32(raw_ring + r * C::kRawBytes);
        mbar_arrive_expect_tx(full, C::kRawBytes);
        tma_load_2d(dst, &map_g, m0, it * BK, full);
#pragma unroll
        for (int j = 0; j < C::kBoxes; ++j)
          tma_load_2d(dst + C::kABytes + j * kBoxBytes, &map_a, n0 + j * kBoxN, it * BK, full);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

}  // namespace tma
}  // namespace mgs
