// K4, TMA-fed variant of the fp32-accurate tcgen05 GEMM (forward / dgrad: K-contiguous activations, weight operand).
//
// Why a second kernel.  ncu on the cp.async kernel of tc_linear.cuh (round 1): tensor pipe 40 % active, L1TEX 69 % busy --
// the 16 producer warps moved every activation element global -> smem (LDGSTS) -> registers -> hi AND lo tiles, and every
// K block pulled 8 KB of activations plus 22.5 KB of pre-split weights (hi | lo) through the L2 -> SM path, which the
// microarchitecture notes cap at ~6300 B / cycle for the whole chip = 42.6 B / cycle / SM: 30.5 KB per K block is 716
// cycles of L2 time for 528 cycles of tensor-core work (BN = 176).  This kernel
//   * lets the TMA engine write the RAW fp32 activation tile (tensor map, SWIZZLE_64B, rows of 16 floats) straight into
//     the UMMA operand slot: kind::tf32 reads only the 19 high bits of each word, so the raw tile IS the `hi` operand
//     (hardware truncation) -- no LDGSTS, no registers, no hi store;
//   * streams the weights RAW as well (packed once per call into the swizzled tile image, one bulk copy per K block):
//     half the bytes of the hi | lo image -> 19.25 KB per K block, 36.5 B / cycle / SM at full tensor rate;
//   * computes both `lo` tiles on chip, element-wise and layout-agnostic: lo = x - trunc_tf32(x) (exact), plus half a
//     TF32 ulp so that the tensor core's own truncation of lo rounds it to nearest (unbiased); one LDS.128 + 12 integer /
//     float instructions + one STS.128 per 16 bytes, the same code for A and B, any swizzle;
//   * keeps the three products hi*hi -> main accumulator, lo*hi + hi*lo -> correction accumulator (tc_linear.cuh).
// Warp roles (persistent, one CTA per SM): 8 converter warps in 2 groups (group q owns raw slots q, q + 2, ...: ~10
// independent LDS.128 / STS.128 pairs per thread and block, two blocks in conversion at any time), 8 epilogue warps
// (TMEM -> registers, accumulator handed back to the MMA warp as soon as it is read, THEN bias / ReLU / stores: the next
// tile's MMAs run under the stores), 1 MMA thread, 1 loader thread.  First version of this kernel (one 5-slot ring of
// raw + lo, all 16 warps converting and then draining): 0.39 ms on the SAGE projection, the same as the cp.async kernel;
// ablation (MGS_TMA_DEBUG): 0.29 ms WITHOUT any MMA, 0.26 ms without loads and MMAs -- barrier round trips and the
// serial epilogue, not bandwidth.  Hence the two rings (raw: 7-10 slots deep) and the dedicated epilogue warps.
//
// Requirements (else the caller falls back to the cp.async kernel): activation base 16-byte aligned and leading
// dimension a multiple of 4 floats (functional.rows() pads 350-float rows to 352).
#pragma once

#include <cuda.h>

#include "tc_linear.cuh"

namespace mgs {
namespace tma {

using namespace tc;

// TS = true: the activation operand of the MMAs lives in TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc): the converter
// thread of row m reads its 64 bytes of the raw tile once and stores hi (truncated) and lo as 16 + 16 columns of lane m
// (tcgen05.st 32x32b.x16).  Both kernels are bound by shared-memory bandwidth: with A in shared
// memory a K block moves 8 (TMA write) + 8 (converter read) + 8 (lo write) + 3 x 8 (UMMA reads) = 48 KB for A, with A in
// TMEM 16 KB; B stays at 6 x 11.25 KB (raw write, converter read, lo write, 3 UMMA reads): 115.6 -> 83.6 KB per block.
// TMEM columns: main accumulator [0, BN), correction [kCorrCol, kCorrCol + BN), 4 A stages of 32 columns in the gaps.
template <int BN, bool TS> struct Cfg {
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128");
  static constexpr int kCorrCol = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int kTmemCols = TS ? 512 : 2 * kCorrCol;
  static_assert(!TS || 2 * kCorrCol + 128 <= 512 || kCorrCol - BN >= 64, "no room for the A stages in tensor memory");
  __host__ __device__ static constexpr uint32_t a_col(int l) {      // TMEM column of A stage l: [hi 16 | lo 16]
    return 2 * kCorrCol + 128 <= 512 ? (uint32_t)(2 * kCorrCol + 32 * l)
                                     : (uint32_t)((l < 2 ? BN : kCorrCol + BN) + 32 * (l & 1));
  }
  static constexpr int kABytes = BM * kRowBytes;           // 8192
  static constexpr int kBBytes = BN * kRowBytes;
  static constexpr int kRawBytes = kABytes + kBBytes;      // one ring slot: [A tile | B tile]
  // Two rings.  RAW slots are filled by the TMA engine many K blocks ahead (they have to cover the load latency: with
  // nothing to compute a block still took ~480 cycles on a 5-slot ring, i.e. a ~2400-cycle round trip per slot); LO slots
  // are written by the converters right before the MMAs that read them and only have to cover the conversion latency.
  static constexpr int kLoStages = 4;
  static constexpr int kBudget = 227 * 1024 - 1024 /* alignment */ - 512 /* barriers */;
  static constexpr int kLoBytes = TS ? kBBytes : kRawBytes;   // TS: the lo slot holds the weight tile only
  static constexpr int kRawFit = (kBudget - kLoStages * kLoBytes) / kRawBytes;
  static constexpr int kRawStages = kRawFit > 10 ? 10 : kRawFit;
  static constexpr int kConvWarps = 8, kGroups = 2;        // converter groups of 4 warps: group q owns raw slots q, q + 2, ...
  static constexpr int kEpiWarps = 8;                      // two per TMEM lane quarter, half of the columns each
  static_assert(!TS || (kLoStages == 4 && kConvWarps / kGroups == 4), "TS: 4 TMEM stages, one converter warp per lane quarter");
  static_assert(kRawStages >= 4, "raw ring too short");
  static constexpr int kSmemBytes = kRawStages * kRawBytes + kLoStages * kLoBytes + 1024 + 512;
  static_assert(kRawBytes % 512 == 0 && kLoBytes % 512 == 0, "SWIZZLE_64B atoms (8 rows x 64 B) must stay aligned");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// D[tmem] (+)= A[tmem: lane = row, one tf32 per column] * B[smem desc]
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}


// Backward of a fused ReLU in a GEMM epilogue: v[u] = 0 where the producer's ReLU output was not positive.  `bits` = the row's
// words of the one-bit-per-element mask the aggregation kernels exchange (stream.cuh: bit l of word t * V + u = column
// (l + 32 t) V + u).
__device__ __forceinline__ void mask_by_bits(float (&v)[8], const uint32_t* __restrict__ bits, int c0, int N, int V) {
  // c0 is a multiple of 8: the eight columns are 8 / V consecutive chunks of ONE 32-chunk group t -> one vector load of the
  // group's V words, then shifts (a first version computed c / V per column: 0.45 instead of 0.34 ms for the SAGE GEMM)
  if (c0 >= N) return;
  if (V == 2) {
    const int ch0 = c0 >> 1, l0 = ch0 & 31;
    const uint2 w = __ldg(reinterpret_cast<const uint2*>(bits + ((ch0 >> 5) << 1)));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (!((w.x >> (l0 + i)) & 1u)) v[2 * i] = 0.f;
      if (!((w.y >> (l0 + i)) & 1u)) v[2 * i + 1] = 0.f;
    }
  } else if (V == 4) {
    const int ch0 = c0 >> 2, l0 = ch0 & 31;
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(bits + ((ch0 >> 5) << 2)));
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (!((w.x >> (l0 + i)) & 1u)) v[4 * i] = 0.f;
      if (!((w.y >> (l0 + i)) & 1u)) v[4 * i + 1] = 0.f;
      if (!((w.z >> (l0 + i)) & 1u)) v[4 * i + 2] = 0.f;
      if (!((w.w >> (l0 + i)) & 1u)) v[4 * i + 3] = 0.f;
    }
  } else {
    const uint32_t w = __ldg(bits + (c0 >> 5));
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (!((w >> ((c0 & 31) + u)) & 1u)) v[u] = 0.f;
  }
}

// lo word of one fp32: exact residual of the TF32 truncation the tensor core applies to the raw word, biased by half a
// TF32 ulp of the residual so that the hardware's truncation of THIS word is a round-to-nearest
__device__ __forceinline__ uint32_t lo_word(uint32_t raw) {
  const float r = __uint_as_float(raw) - __uint_as_float(raw & 0xffffe000u);
  return __float_as_uint(r) + 0x1000u;
}

// Weight packer, raw variant: for every (N tile nt, K block kb) the swizzled shared-memory image of the BN x 16 fp32 tile
// at  out + (nt * nkb + kb) * BN * 64  (rows beyond N / columns beyond K are zero).
__global__ void __launch_bounds__(256)
pack_b_raw_kernel(Operand b0, int K0, Operand b1, int K1, int N, int BN, uint8_t* __restrict__ out) {
  const int nkb0 = (K0 + BK - 1) / BK, nkb1 = (K1 + BK - 1) / BK;
  const int nkb = nkb0 + nkb1;
  const int ntiles = (N + BN - 1) / BN;
  const int64_t total = (int64_t)ntiles * nkb * BN * kChunks;
  for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256) {
    const int c = (int)(idx % kChunks);
    int64_t t = idx / kChunks;
    const int r = (int)(t % BN);
    t /= BN;
    const int kb = (int)(t % nkb);
    const int nt = (int)(t / nkb);
    const bool first = kb < nkb0;
    const Operand& op = first ? b0 : b1;
    const int kend = first ? K0 : K1;
    const int k = (first ? kb : kb - nkb0) * BK + 4 * c;
    const int n = nt * BN + r;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (n < N) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (k + u < kend) v[u] = op.k_contig ? __ldg(op.p + (int64_t)n * op.ld + k + u) : __ldg(op.p + (int64_t)(k + u) * op.ld + n);
    }
    uint8_t* tile = out + ((int64_t)nt * nkb + kb) * ((int64_t)BN * kRowBytes);
    *reinterpret_cast<float4*>(tile + swz_off(r, c)) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// C[m][n] = sum_k A0(m,k) B(k,n) (+ sum_k A1(m,k) B1(k,n)) (+ bias[n]) (ReLU).   A0 / A1 through tensor maps
// (dims {K, M}, box {16, 128}, SWIZZLE_64B, zero fill), B pre-packed by pack_b_raw_kernel.
// splits > 1: K blocks of segment 0 are dealt to `splits` partial outputs c + z * split_stride (no bias / ReLU there).
template <int BN, bool TS>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1, int K0, int K1,
                const float* __restrict__ a0, int64_t lda0, const float* __restrict__ a1, int64_t lda1,
                const uint8_t* __restrict__ packed_b, int M, int N, float* __restrict__ c, int64_t ldc,
                const float* __restrict__ bias, int relu, int splits, int64_t split_stride,
                int dbg /* timing experiments only (results are garbage): 4 no MMA, 8 no conversion, 16 no stores */,
                const uint32_t* __restrict__ mask_bits = nullptr, int mask_words = 0, int mask_v = 0) {
  using C = Cfg<BN, TS>;
  static_assert(kThreads == (C::kConvWarps + C::kEpiWarps + 2) * 32, "warp roles");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int R = C::kRawStages, L = C::kLoStages;
  uint8_t* raw_ring = smem;                                   // R slots of [A raw | B raw]
  uint8_t* lo_ring = smem + R * C::kRawBytes;                 // L slots of [A lo | B lo]  (TS: [B lo], A in tensor memory)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + R * C::kRawBytes + L * C::kLoBytes);
  // bars: [0,R) raw landed (TMA tx) | [R,2R) raw slot free (MMAs retired) | [2R,2R+L) lo written (converter group) |
  //       [2R+L,2R+2L) lo slot free (MMAs retired) | acc_full | tmem_free | TMEM base address
  auto bar_raw_full = [&](int i) { return smem_u32(bars + i); };
  auto bar_raw_free = [&](int i) { return smem_u32(bars + R + i); };
  auto bar_lo_full = [&](int i) { return smem_u32(bars + 2 * R + i); };
  auto bar_lo_free = [&](int i) { return smem_u32(bars + 2 * R + L + i); };
  const uint32_t bar_acc_full = smem_u32(bars + 2 * R + 2 * L);
  const uint32_t bar_tmem_free = smem_u32(bars + 2 * R + 2 * L + 1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * R + 2 * L + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb0 = (K0 + BK - 1) / BK, nb1 = (K1 + BK - 1) / BK, nb = nb0 + nb1;
  const int ntn = (N + BN - 1) / BN, ntm = (M + BM - 1) / BM;
  const int nmn = ntn * ntm;
  const int ntiles = nmn * splits;
  const int nb_split = (nb + splits - 1) / splits;          // K blocks per split (splits > 1: one segment only)
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
    // ================= converters: raw slot -> lo slot, element-wise (the same code for A and B, any swizzle) ==========
    constexpr int kVec = C::kLoBytes / 16;                   // float4 items per lo slot
    constexpr int kSkip = TS ? C::kABytes / 16 : 0;          // TS: only the weight part of the raw slot is converted here
    constexpr int kGroupThreads = (C::kConvWarps / C::kGroups) * 32;
    constexpr int kPer = (kVec + kGroupThreads - 1) / kGroupThreads;
    const int grp = warp / (C::kConvWarps / C::kGroups), gt = threadIdx.x % kGroupThreads;
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int z = tile / nmn;
      const int it_lo = z * nb_split, it_hi = min(nb, it_lo + nb_split);
      for (int it = it_lo; it < it_hi; ++it, ++g) {
        // A raw SLOT belongs to one group (slot % kGroups): a group then meets the phases of its slots' barriers one by
        // one, and a slot is not refilled before its owner has converted it -- a parity wait is only valid for a barrier's
        // current or previous phase.  (History: blocks dealt by g % kGroups with every group observing every block's
        // barrier in order had no such flow control for the OBSERVING group: when it was the late one, the block it
        // had yet to observe could be converted by the other group, multiplied, retired and its slot refilled before the
        // observer polled -- two phases ahead, the wait then never returns.  Seen as a rare hang of the converter-bound
        // BN = 128 kernel, about one launch in twenty at 130 k rows.)
        if ((int)((g % R) % C::kGroups) != grp) continue;
        mbar_wait(bar_raw_full((int)(g % R)), (g / R) & 1u);
        const int l = (int)(g % L);
        mbar_wait(bar_lo_free(l), ((g / L) & 1u) ^ 1u);               // MMAs that read this lo slot have retired
        const uint8_t* slot = raw_ring + (g % R) * C::kRawBytes;
        const uint4* raw = reinterpret_cast<const uint4*>(slot) + kSkip;
        uint4* lo = reinterpret_cast<uint4*>(lo_ring + l * C::kLoBytes);
        if constexpr (TS) {
          if (!(dbg & 8)) {
            // activation row (warp % 4) * 32 + lane -> TMEM lane of the same number: 16 hi + 16 lo columns
            tc_fence_after();
            const int row = (warp & 3) * 32 + lane;
            uint32_t hi[16], lw[16];
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
              const uint4 v = *reinterpret_cast<const uint4*>(slot + swz_off(row, c));
              hi[4 * c + 0] = v.x & 0xffffe000u; hi[4 * c + 1] = v.y & 0xffffe000u;
              hi[4 * c + 2] = v.z & 0xffffe000u; hi[4 * c + 3] = v.w & 0xffffe000u;
              lw[4 * c + 0] = lo_word(v.x); lw[4 * c + 1] = lo_word(v.y);
              lw[4 * c + 2] = lo_word(v.z); lw[4 * c + 3] = lo_word(v.w);
            }
            const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + C::a_col(l);
            tmem_st16(ta, hi);
            tmem_st16(ta + 16, lw);
          }
        }
        if (!(dbg & 8)) {
          uint4 v[kPer];
#pragma unroll
          for (int j = 0; j < kPer; ++j)
            if (gt + j * kGroupThreads < kVec) v[j] = raw[gt + j * kGroupThreads];
#pragma unroll
          for (int j = 0; j < kPer; ++j)
            if (gt + j * kGroupThreads < kVec)
              lo[gt + j * kGroupThreads] = make_uint4(lo_word(v[j].x), lo_word(v[j].y), lo_word(v[j].z), lo_word(v[j].w));
        }
        fence_proxy_async();                                          // generic-proxy writes -> async proxy (UMMA)
        if constexpr (TS) {
          tmem_st_wait();
          tc_fence_before();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_lo_full(l));
      }
    }
  } else if (warp < kMmaWarp) {
    // ================= epilogue warps: TMEM -> registers (accumulator released) -> (+ bias, ReLU) -> global ============
    const int ew = warp - C::kConvWarps;
    const int q = warp & 3;                                  // TMEM lane quarter this warp may access (warp id % 4)
    const int half = ew >> 2;                                // column half
    const int row_l = q * 32 + lane;
    constexpr int kChunksW = BN / 16;                        // 8-column chunks per warp
    constexpr int kPass = kChunksW > 11 ? (kChunksW + 1) / 2 : kChunksW;   // chunks held in registers at a time
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(c) | (uintptr_t)(ldc * 4) | (uintptr_t)(split_stride * 4)) & 15u) == 0;
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
      const int z = tile / nmn, mn = tile - z * nmn;
      const int m0 = (mn / ntn) * BM, n0 = (mn % ntn) * BN;
      const bool any = min(nb, z * nb_split + nb_split) > z * nb_split;
      mbar_wait(bar_acc_full, tl & 1u);
      tc_fence_after();
      float* crow = c + (int64_t)z * split_stride + (int64_t)(m0 + row_l) * ldc + n0;
      const bool row_ok = m0 + row_l < M;
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
        if (p0 + kPass >= kChunksW) {                                  // last pass read: the MMA warp may reuse the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tmem_free);
        }
#pragma unroll
        for (int j = 0; j < kPass; ++j) {
          if (p0 + j < kChunksW) {
            const int nl = 8 * (half * kChunksW + p0 + j);
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              v[u] = acc[j][u];
              if (bias != nullptr && n0 + nl + u < N) v[u] += __ldg(bias + n0 + nl + u);
              if (relu) v[u] = v[u] <= 0.f ? 0.f : v[u];
            }
            if (mask_bits != nullptr && row_ok) mask_by_bits(v, mask_bits + (int64_t)(m0 + row_l) * mask_words, n0 + nl, N, mask_v);
            if (row_ok && n0 + nl < N && !(dbg & 16)) {
              if (vec_ok && n0 + nl + 8 <= N) {
                *reinterpret_cast<float4*>(crow + nl) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(crow + nl + 4) = make_float4(v[4], v[5], v[6], v[7]);
              } else {
#pragma unroll
                for (int u = 0; u < 8; ++u)
                  if (n0 + nl + u < N) crow[nl + u] = v[u];
              }
            }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0) {
      // ================= MMA issuer (one thread) =================
      constexpr uint32_t idesc = make_idesc(BN);
      uint32_t g = 0, tl = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
        const int z = tile / nmn;
        const int it_lo = z * nb_split, it_hi = min(nb, it_lo + nb_split);
        mbar_wait(bar_tmem_free, (tl & 1u) ^ 1u);                     // previous tile drained (first tile: passes)
        tc_fence_after();
        for (int it = it_lo; it < it_hi; ++it, ++g) {
          const int r = (int)(g % R), l = (int)(g % L);
          mbar_wait(bar_lo_full(l), (g / L) & 1u);                    // lo written (its converters saw the raw tiles land)
          tc_fence_after();
          const uint32_t sr = smem_u32(raw_ring + r * C::kRawBytes), sl = smem_u32(lo_ring + l * C::kLoBytes);
          const uint64_t b_hi = make_desc(sr + C::kABytes);
          const uint64_t b_lo = make_desc(TS ? sl : sl + C::kABytes);
          const uint64_t a_hi = make_desc(sr), a_lo = make_desc(sl);  // SS form only
          const uint32_t ta = tmem_base + C::a_col(l);                // TS form only
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            if (dbg & 4) break;
            const uint64_t adv = (uint64_t)(k * 32 >> 4);             // 8 tf32 = 32 bytes along the swizzled row
            const uint32_t first = (it != it_lo || k != 0) ? 1u : 0u;
            if constexpr (TS) {
              umma_tf32_ts(tmem_base, ta + 8 * k, b_hi + adv, idesc, first);
              umma_tf32_ts(tmem_base + C::kCorrCol, ta + 16 + 8 * k, b_hi + adv, idesc, first);
              umma_tf32_ts(tmem_base + C::kCorrCol, ta + 8 * k, b_lo + adv, idesc, 1);
            } else {
              umma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, first);
              umma_tf32(tmem_base + C::kCorrCol, a_lo + adv, b_hi + adv, idesc, first);
              umma_tf32(tmem_base + C::kCorrCol, a_hi + adv, b_lo + adv, idesc, 1);
            }
          }
          umma_commit(bar_raw_free(r));                               // both slots are free when these MMAs retire
          umma_commit(bar_lo_free(l));
        }
        umma_commit(bar_acc_full);
      }
    }
  } else if (lane == 0) {
    // ================= loader (one thread, TMA engine): activation tile by tensor map, weight tile by bulk copy =========
    tma_prefetch_desc(&map_a0);
    if (nb1 > 0) tma_prefetch_desc(&map_a1);
    auto prefetch_rows = [&](int tile) {        // a tile's activation rows -> L2: ONE contiguous span per segment
      if (tile >= ntiles) return;
      const int pm0 = ((tile % nmn) / ntn) * BM;
      const int rows = min(BM, M - pm0);
      const uint32_t b0 = (uint32_t)(((int64_t)(rows - 1) * lda0 + K0) * 4) & ~15u;
      if (b0 >= 16) bulk_prefetch_l2(a0 + (int64_t)pm0 * lda0, b0);
      if (nb1 > 0) {
        const uint32_t b1 = (uint32_t)(((int64_t)(rows - 1) * lda1 + K1) * 4) & ~15u;
        if (b1 >= 16) bulk_prefetch_l2(a1 + (int64_t)pm0 * lda1, b1);
      }
    };
    uint32_t g = 0;
    prefetch_rows(blockIdx.x);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      prefetch_rows(tile + gridDim.x);
      const int z = tile / nmn, mn = tile - z * nmn;
      const int m0 = (mn / ntn) * BM;
      const int it_lo = z * nb_split, it_hi = min(nb, it_lo + nb_split);
      const uint8_t* src = packed_b + (int64_t)(mn % ntn) * nb * C::kBBytes;
      for (int it = it_lo; it < it_hi; ++it, ++g) {
        const int r = (int)(g % R);
        mbar_wait(bar_raw_free(r), ((g / R) & 1u) ^ 1u);              // MMAs that read this slot have retired
        const uint32_t full = bar_raw_full(r);
        const uint32_t dst = smem_u32(raw_ring + r * C::kRawBytes);
        mbar_arrive_expect_tx(full, C::kRawBytes);
        if (it < nb0) tma_load_2d(dst, &map_a0, it * BK, m0, full);
        else tma_load_2d(dst, &map_a1, (it - nb0) * BK, m0, full);
        bulk_g2s(dst + C::kABytes, src + (int64_t)it * C::kBBytes, C::kBBytes, full);
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


// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2, clusters of two CTAs on one TPC).  Section 4.9 of DESIGN.md: the one-CTA kernel
// is bound by shared-memory bandwidth -- every 128 x BN x 8 MMA reads the whole BN x 32-byte weight slab from the SM's
// shared memory (64 B / cycle for the MMA reads alone, + the staging of the raw / lo weight tiles).  A pair works on a
// 256 x BN tile: CTA r owns rows [128 r, 128 r + 128) of it (its activation tile in ITS tensor memory, its half of the
// accumulators in ITS tensor memory) and stages only rows [r BN / 2, (r + 1) BN / 2) of the weight tile; one
// tcgen05.mma.cta_group::2 (M = 256), issued by the leader's MMA thread, drives both tensor cores and lets each read the
// other's half.  Per SM and K block: weight bytes written by TMA, read and re-written by the converters, read three times by
// the MMAs all halve (67.5 -> 33.75 KB), the activation part stays (16 KB): 83.6 -> 49.8 KB per 528 tensor cycles.
// Barriers: raw-landed / raw-free / lo-free / accumulator-full are per CTA (tcgen05.commit multicasts its arrival to both);
// lo-written and accumulator-drained live in the LEADER and count the warps of both CTAs (remote arrivals, release /
// acquire at cluster scope).  Same K order and products as the one-CTA kernel: bit-identical results.
template <int BN> struct Cfg2 {
  using Base = Cfg<BN, true>;
  static_assert(BN % 32 == 0 || (BN / 2) % 8 == 0, "each CTA's half of a weight tile must be whole 8-row swizzle atoms");
  static constexpr int kCorrCol = Base::kCorrCol;
  static constexpr int kTmemCols = 512;
  __host__ __device__ static constexpr uint32_t a_col(int l) { return Base::a_col(l); }
  static constexpr int kABytes = BM * kRowBytes;                  // 8192: this CTA's 128 activation rows
  static constexpr int kBFull = BN * kRowBytes;                   // packed image of one (N tile, K block)
  static constexpr int kBBytes = kBFull / 2;                      // this CTA's half
  static constexpr int kRawBytes = kABytes + kBBytes;
  static constexpr int kLoStages = 4;
  static constexpr int kLoBytes = kBBytes;
  static constexpr int kBudget = 227 * 1024 - 1024 - 512;
  static constexpr int kRawFit = (kBudget - kLoStages * kLoBytes) / kRawBytes;
  static constexpr int kRawStages = kRawFit > 12 ? 12 : kRawFit;
  static constexpr int kConvWarps = 8, kGroups = 2, kEpiWarps = 8;
  static constexpr int kSmemBytes = kRawStages * kRawBytes + kLoStages * kLoBytes + 1024 + 512;
  static_assert(kRawBytes % 512 == 0 && kLoBytes % 512 == 0, "SWIZZLE_64B atoms (8 rows x 64 B) must stay aligned");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Remote arrival with the default semantics (release at CTA scope).  What the arrival publishes -- this CTA's lo tile in
// ITS shared memory (made visible to the async proxy by fence.proxy.async) and its activation rows in ITS tensor memory
// (tcgen05.wait::st) -- is consumed by this SM's tensor core, never by the leader's threads; the leader's MMA thread only
// needs the arrival itself.  (`.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR per arrival: measured 0.50 ms instead
// of 0.35 ms for the SAGE projection, the converters spent a sixth of their samples on the ERRBAR.)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// arrival on the barrier at this CTA-relative offset in BOTH CTAs of the pair once every MMA issued so far has retired
__device__ __forceinline__ void umma_commit2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_tf32_ts2(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor of the pair MMA: M = 256 (128 rows per CTA), N = BN
__host__ __device__ constexpr uint32_t make_idesc2(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tma2_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1, int K0, int K1,
                 const float* __restrict__ a0, int64_t lda0, const float* __restrict__ a1, int64_t lda1,
                 const uint8_t* __restrict__ packed_b, int M, int N, float* __restrict__ c, int64_t ldc,
                 const float* __restrict__ bias, int relu, int splits, int64_t split_stride,
                 int dbg /* timing experiments only: 4 no MMA, 8 no conversion, 32 lo slots released by the raw-slot barrier */,
                 const uint32_t* __restrict__ mask_bits = nullptr, int mask_words = 0, int mask_v = 0,
                 float* __restrict__ colsum_part = nullptr /* [ceil(M / 256) * 8][N]: column sums of every 32-row group */) {
  using C = Cfg2<BN>;
  static_assert(kThreads == (C::kConvWarps + C::kEpiWarps + 2) * 32, "warp roles");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int R = C::kRawStages, L = C::kLoStages;
  uint8_t* raw_ring = smem;                                   // R slots of [A raw (128 rows) | B raw (BN / 2 rows)]
  uint8_t* lo_ring = smem + R * C::kRawBytes;                 // L slots of [B lo (BN / 2 rows)]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + R * C::kRawBytes + L * C::kLoBytes);
  auto bar_raw_full = [&](int i) { return smem_u32(bars + i); };
  auto bar_raw_free = [&](int i) { return smem_u32(bars + R + i); };
  auto bar_lo_full = [&](int i) { return smem_u32(bars + 2 * R + i); };          // used in the leader only
  auto bar_lo_free = [&](int i) { return smem_u32(bars + 2 * R + L + i); };
  const uint32_t bar_acc_full = smem_u32(bars + 2 * R + 2 * L);
  const uint32_t bar_tmem_free = smem_u32(bars + 2 * R + 2 * L + 1);              // used in the leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * R + 2 * L + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int ncl = (int)(gridDim.x >> 1), cl = (int)(blockIdx.x >> 1);
  const int nb0 = (K0 + BK - 1) / BK, nb1 = (K1 + BK - 1) / BK, nb = nb0 + nb1;
  const int ntn = (N + BN - 1) / BN, ntm = (M + 2 * BM - 1) / (2 * BM);
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
      mbar_init(bar_lo_full(i), 2 * (C::kConvWarps / C::kGroups));
      mbar_init(bar_lo_free(i), 1);
    }
    mbar_init(bar_acc_full, 1);
    mbar_init(bar_tmem_free, 2 * C::kEpiWarps);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc2(smem_u32(tmem_slot), C::kTmemCols);
  tc_fence_before();
  cluster_sync_all();                                         // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < C::kConvWarps) {
    // ================= converters: own activation rows -> tensor memory (hi | lo), own weight half -> lo slot ============
    constexpr int kVec = C::kLoBytes / 16;
    constexpr int kSkip = C::kABytes / 16;
    constexpr int kGroupThreads = (C::kConvWarps / C::kGroups) * 32;
    constexpr int kPer = (kVec + kGroupThreads - 1) / kGroupThreads;
    const int grp = warp / (C::kConvWarps / C::kGroups), gt = threadIdx.x % kGroupThreads;
    uint32_t lo_full_leader[L];
#pragma unroll
    for (int i = 0; i < L; ++i) lo_full_leader[i] = mapa_u32(bar_lo_full(i), 0);
    uint32_t g = 0;
    for (int tile = cl; tile < ntiles; tile += ncl) {
      const int z = tile / nmn;
      const int it_lo = z * nb_split, it_hi = min(nb, it_lo + nb_split);
      for (int it = it_lo; it < it_hi; ++it, ++g) {
        if ((int)((g % R) % C::kGroups) != grp) continue;                // a raw slot belongs to one group (see above)
        mbar_wait(bar_raw_full((int)(g % R)), (g / R) & 1u);
        const int l = (int)(g % L);
        if (dbg & 32) {
          // block g - L used this lo slot / A stage; its MMAs have retired once ITS raw slot was released (one commit per block)
          if (g >= (uint32_t)L) mbar_wait(bar_raw_free((int)((g - L) % R)), ((g - L) / R) & 1u);
        } else {
          mbar_wait(bar_lo_free(l), ((g / L) & 1u) ^ 1u);
        }
        const uint8_t* slot = raw_ring + (g % R) * C::kRawBytes;
        const uint4* raw = reinterpret_cast<const uint4*>(slot) + kSkip;
        uint4* lo = reinterpret_cast<uint4*>(lo_ring + l * C::kLoBytes);
        if (!(dbg & 8)) {
          tc_fence_after();
          const int row = (warp & 3) * 32 + lane;
          uint32_t hi[16], lw[16];
#pragma unroll
          for (int cc = 0; cc < kChunks; ++cc) {
            const uint4 v = *reinterpret_cast<const uint4*>(slot + swz_off(row, cc));
            hi[4 * cc + 0] = v.x & 0xffffe000u; hi[4 * cc + 1] = v.y & 0xffffe000u;
            hi[4 * cc + 2] = v.z & 0xffffe000u; hi[4 * cc + 3] = v.w & 0xffffe000u;
            lw[4 * cc + 0] = lo_word(v.x); lw[4 * cc + 1] = lo_word(v.y);
            lw[4 * cc + 2] = lo_word(v.z); lw[4 * cc + 3] = lo_word(v.w);
          }
          const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + C::a_col(l);
          tmem_st16(ta, hi);
          tmem_st16(ta + 16, lw);
        }
        if (!(dbg & 8)) {
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
        if (lane == 0) mbar_arrive_cluster(lo_full_leader[l]);
      }
    }
  } else if (warp < kMmaWarp) {
    // ================= epilogue warps: own 128 rows of the pair tile ====================================================
    const int ew = warp - C::kConvWarps;
    const int q = warp & 3;
    const int half = ew >> 2;
    const int row_l = q * 32 + lane;
    constexpr int kChunksW = BN / 16;
    constexpr int kPass = kChunksW > 11 ? (kChunksW + 1) / 2 : kChunksW;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(c) | (uintptr_t)(ldc * 4) | (uintptr_t)(split_stride * 4)) & 15u) == 0;
    const uint32_t tmem_free_leader = mapa_u32(bar_tmem_free, 0);
    uint32_t tl = 0;
    for (int tile = cl; tile < ntiles; tile += ncl, ++tl) {
      const int z = tile / nmn, mn = tile - z * nmn;
      const int m0 = (mn / ntn) * (2 * BM) + (int)rank * BM, n0 = (mn % ntn) * BN;
      const bool any = min(nb, z * nb_split + nb_split) > z * nb_split;
      mbar_wait(bar_acc_full, tl & 1u);
      tc_fence_after();
      float* crow = c + (int64_t)z * split_stride + (int64_t)(m0 + row_l) * ldc + n0;
      const bool row_ok = m0 + row_l < M;
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
          if (lane == 0) mbar_arrive_cluster(tmem_free_leader);
        }
#pragma unroll
        for (int j = 0; j < kPass; ++j) {
          if (p0 + j < kChunksW) {
            const int nl = 8 * (half * kChunksW + p0 + j);
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              v[u] = acc[j][u];
              if (bias != nullptr && n0 + nl + u < N) v[u] += __ldg(bias + n0 + nl + u);
              if (relu) v[u] = v[u] <= 0.f ? 0.f : v[u];
            }
            if (mask_bits != nullptr && row_ok) mask_by_bits(v, mask_bits + (int64_t)(m0 + row_l) * mask_words, n0 + nl, N, mask_v);
            if (colsum_part != nullptr) {
              // column sums of the output (the bias gradient of the layer BELOW: what a separate pass over the [M, N] result
              // would compute): butterfly over the warp's 32 rows, lane u keeps column u of the chunk; fixed order
              float cs[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) cs[u] = (row_ok && n0 + nl + u < N) ? v[u] : 0.f;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int u = 0; u < 8; ++u) cs[u] += __shfl_xor_sync(0xffffffffu, cs[u], o);
              float mine = cs[0];
#pragma unroll
              for (int u = 1; u < 8; ++u) mine = lane == u ? cs[u] : mine;
              if (lane < 8 && n0 + nl + lane < N)
                colsum_part[(int64_t)((m0 + q * 32) >> 5) * N + n0 + nl + lane] = mine;
            }
            if (row_ok && n0 + nl < N) {
              if (vec_ok && n0 + nl + 8 <= N) {
                *reinterpret_cast<float4*>(crow + nl) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(crow + nl + 4) = make_float4(v[4], v[5], v[6], v[7]);
              } else {
#pragma unroll
                for (int u = 0; u < 8; ++u)
                  if (n0 + nl + u < N) crow[nl + u] = v[u];
              }
            }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0 && rank == 0) {
      // ================= MMA issuer: one thread of the LEADER drives both tensor cores ==================================
      constexpr uint32_t idesc = make_idesc2(BN);
      uint32_t g = 0, tl = 0;
      for (int tile = cl; tile < ntiles; tile += ncl, ++tl) {
        const int z = tile / nmn;
        const int it_lo = z * nb_split, it_hi = min(nb, it_lo + nb_split);
        mbar_wait_cluster(bar_tmem_free, (tl & 1u) ^ 1u);               // both CTAs drained the previous tile
        tc_fence_after();
        for (int it = it_lo; it < it_hi; ++it, ++g) {
          const int r = (int)(g % R), l = (int)(g % L);
          mbar_wait_cluster(bar_lo_full(l), (g / L) & 1u);              // both CTAs: raw tiles landed, lo / A written
          tc_fence_after();
          const uint32_t sr = smem_u32(raw_ring + r * C::kRawBytes), sl = smem_u32(lo_ring + l * C::kLoBytes);
          const uint64_t b_hi = make_desc(sr + C::kABytes);
          const uint64_t b_lo = make_desc(sl);
          const uint32_t ta = tmem_base + C::a_col(l);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            if (dbg & 4) break;
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            const uint32_t first = (it != it_lo || k != 0) ? 1u : 0u;
            umma_tf32_ts2(tmem_base, ta + 8 * k, b_hi + adv, idesc, first);
            umma_tf32_ts2(tmem_base + C::kCorrCol, ta + 16 + 8 * k, b_hi + adv, idesc, first);
            umma_tf32_ts2(tmem_base + C::kCorrCol, ta + 8 * k, b_lo + adv, idesc, 1);
          }
          umma_commit2(bar_raw_free(r));
          if (!(dbg & 32)) umma_commit2(bar_lo_free(l));
        }
        umma_commit2(bar_acc_full);
      }
    }
  } else if (lane == 0) {
    // ================= loader (one thread per CTA): own activation rows, own half of the weight tile ====================
    tma_prefetch_desc(&map_a0);
    if (nb1 > 0) tma_prefetch_desc(&map_a1);
    auto prefetch_rows = [&](int tile) {
      if (tile >= ntiles) return;
      const int pm0 = ((tile % nmn) / ntn) * (2 * BM) + (int)rank * BM;
      const int rows = min(BM, M - pm0);
      if (rows <= 0) return;
      const uint32_t b0 = (uint32_t)(((int64_t)(rows - 1) * lda0 + K0) * 4) & ~15u;
      if (b0 >= 16) bulk_prefetch_l2(a0 + (int64_t)pm0 * lda0, b0);
      if (nb1 > 0) {
        const uint32_t b1 = (uint32_t)(((int64_t)(rows - 1) * lda1 + K1) * 4) & ~15u;
        if (b1 >= 16) bulk_prefetch_l2(a1 + (int64_t)pm0 * lda1, b1);
      }
    };
    uint32_t g = 0;
    prefetch_rows(cl);
    for (int tile = cl; tile < ntiles; tile += ncl) {
      prefetch_rows(tile + ncl);
      const int z = tile / nmn, mn = tile - z * nmn;
      const int m0 = (mn / ntn) * (2 * BM) + (int)rank * BM;
      const int it_lo = z * nb_split, it_hi = min(nb, it_lo + nb_split);
      const uint8_t* src = packed_b + (int64_t)(mn % ntn) * nb * C::kBFull + (int64_t)rank * C::kBBytes;
      for (int it = it_lo; it < it_hi; ++it, ++g) {
        const int r = (int)(g % R);
        mbar_wait(bar_raw_free(r), ((g / R) & 1u) ^ 1u);
        const uint32_t full = bar_raw_full(r);
        const uint32_t dst = smem_u32(raw_ring + r * C::kRawBytes);
        mbar_arrive_expect_tx(full, C::kRawBytes);
        if (it < nb0) tma_load_2d(dst, &map_a0, it * BK, m0, full);
        else tma_load_2d(dst, &map_a1, (it - nb0) * BK, m0, full);
        bulk_g2s(dst + C::kABytes, src + (int64_t)it * C::kBFull, C::kBBytes, full);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();                                         // nobody leaves while the peer may still signal or read it
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, C::kTmemCols);
  }
}

}  // namespace tma
}  // namespace mgs
