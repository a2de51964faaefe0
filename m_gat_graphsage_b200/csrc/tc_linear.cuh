// K4 (tensor-core variant) -- fp32-accurate GEMM on the 5th-generation tensor cores (tcgen05 + TMEM),
// used for every projection large enough to fill a 128-row tile.
//
// Precision: "3xTF32".  Each fp32 operand x is split in registers into hi = rna_tf32(x) and
// lo = rna_tf32(x - hi); the tensor core accumulates hi*hi + lo*hi + hi*lo in fp32 (TMEM).  The dropped
// lo*lo term and the rounding of lo are ~2^-22 relative and unbiased, i.e. fp32-class accuracy -- a
// single TF32 pass (2^-11) would miss the 1e-5 logit tolerance by two orders of magnitude.
// The tensor core's fp32 accumulation TRUNCATES (measured on B200: error grows linearly with the number
// of accumulations into one TMEM tile, ~2e-8 per tcgen05.mma), so the main term hi*hi and the two small
// correction terms go to SEPARATE TMEM accumulators: the main accumulator sees K/8 instead of 3K/8
// truncating accumulations, the corrections (2^-11 of the result) truncate harmlessly; the epilogue adds
// the two tiles in fp32.
//
// Structure (one 128 x BN output tile per CTA, 576 threads):
//   warps 0-15 producers: coalesced LDG of the fp32 A / B tiles straight from global memory (activations
//              have 1400-byte rows: not TMA-able without a padded copy), hi/lo split in registers,
//              conflict-free 128-bit STS into the canonical K-major SWIZZLE_64B layout (both operand
//              orientations are transposed on the fly, so shared memory is always K-major),
//              fence.proxy.async, mbarrier arrive;  after the K loop the same warps are the epilogue:
//              tcgen05.ld TMEM -> registers -> (+bias) -> global.
//   warp 16    lane 0 issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8) x 2 k-steps x 3
//              split products per 16-float K block, tcgen05.commit releases the stage;  also owns the
//              TMEM allocation.
//   warp 17    (B_PACKED) lane 0 streams the weight operand with cp.async.bulk (TMA engine, 1-D):
//              weights are tiny (<= 4 MB), so tc_pack_b_kernel splits and swizzles them ONCE per call
//              into exactly the shared-memory image of every (N tile, K block); a stage's B_hi|B_lo is
//              then one 16-32 KB bulk copy completing on the stage's full barrier (complete_tx).
//              For wgrad both operands are activation-sized and both go through the producer warps.
// Pipeline: 3-5 shared-memory stages of one 16-float K block each (full/empty mbarriers), so the bulk
// copies and the MMAs run several blocks apart; the producers keep the cp.async copies of the next
// kDepth = 6 (4 for wgrad) blocks in flight into a raw fp32 ring while block i is converted.  (A first version with two
// 32-float stages ran at ~3000 cycles per block: every weight copy was issued only when its stage had
// just been released and then sat on the critical path.)
#pragma once

#include "common.cuh"

namespace mgs {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 16;                 // floats per K block = one 64-byte swizzle row (SWIZZLE_64B)
constexpr int kRowBytes = BK * 4;      // 64
constexpr int kChunks = kRowBytes / 16;  // 16-byte chunks per row = 4
constexpr int kProducerWarps = 16;    // the hi/lo conversion is instruction bound: 8 warps needed ~650 issue
                                      // cycles per 16-float K block, more than the 528 tensor-core cycles it feeds
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kThreads = kProducerThreads + 64;   // + MMA warp + bulk-copy warp

struct Operand {
  const float* p;
  int64_t ld;
  int vec;       // widest aligned vector along the contiguous dimension (1, 2 or 4 floats)
  int k_contig;  // 1: elem(r,k) = p[r*ld + k]   0: elem(r,k) = p[k*ld + r]
};
struct Segment {
  Operand a, b;
  int K;
};

template <int BN, bool PACKED> struct Cfg {
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128");
  static constexpr int kCorrCol = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;  // correction tile
  static constexpr int kTmemCols = 2 * kCorrCol;
  static constexpr int kABytes = BM * kRowBytes;
  static constexpr int kBBytes = BN * kRowBytes;
  static constexpr int kStageBytes = 2 * (kABytes + kBBytes);
  // (row, 16-byte chunk) items per producer thread: 512 threads x 4 chunks cover 128 rows of a K-contiguous
  // tile in one pass; MN-contiguous tiles (wgrad) are covered by row PAIRS (2 items): 4 k-groups x 4 groups
  // of 32 pairs = 256 rows
  static constexpr int kItemsA = PACKED ? BM / 128 : 2;
  static constexpr int kItemsB = PACKED ? 0 : 2;
  // raw fp32 staging ring filled by cp.async: one 16-byte slot per (thread, item), kDepth K blocks deep
  static constexpr int kRawBytes = (kItemsA + kItemsB) * kProducerThreads * 16;
  // BN = 128 with packed weights is sized for TWO CTAs per SM (96 KB of shared memory, 256 TMEM columns each):
  // one CTA's prologue / epilogue then overlaps the other's K loop (with one CTA per SM the tensor pipe idles
  // during every tile's TMEM drain and global stores: ~0.27 ms of the 0.47 ms SAGE projection was skeleton).
  static constexpr int kCtasPerSm = 1;
  static constexpr int kDepth = PACKED ? (kCtasPerSm == 2 ? 4 : 6) : 0;
  // non-packed (wgrad): register-staged operands, no raw ring -> the shared memory goes to pipeline stages; the
  // producer groups need kStages >= number of groups (a group may then never wait two barrier phases behind)
  static constexpr int kStages = PACKED ? (kCtasPerSm == 2 ? 2 : BN <= 176 ? 4 : 3) : (BN <= 128 ? 6 : BN <= 176 ? 5 : 4);
  static constexpr int kSmemBytes = kStages * kStageBytes + kDepth * kRawBytes + 1024 /* alignment */ + 256 /* barriers */;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk asynchronous copy global -> shared (TMA engine), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// same, multicast to the same CTA-relative shared-memory offsets of every CTA in `cta_mask`; the
// complete_tx signal is multicast to the barrier at the same offset in each destination CTA
__device__ __forceinline__ void bulk_g2s_mcast(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
      ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mcast(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Ampere-style asynchronous copies global -> shared (LDGSTS): `bytes` of the cp-size are read, the rest of the
// destination is zero-filled.  Completion is tracked per thread in commit groups, NOT on the 6-slot register
// scoreboard -- a register-staged prefetch of many K blocks aliases scoreboard slots and ends up waiting for
// the newest load on every use (measured: ~1500 cycles per K block no matter the prefetch depth).
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// whole-tile L2 prefetch (TMA engine): an A tile of a row-major activation is ONE contiguous span
// (128 rows x K floats), which DRAM streams far better than the 64-byte-per-row pieces the K loop asks for
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_64B shared-memory matrix descriptor (sm_100 format, cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4
//   [46,48) version = 1 | [49,52) base offset = 0 | [61,64) layout type = 4 (SWIZZLE_64B)
// Canonical layout Swizzle<2,4,3> o ((8,m),(T,2)):((4T,SBO),(1,T)), T = 4 tf32: rows are 64 bytes, 8-row
// swizzle atoms are 512 bytes apart (SBO), the 16-byte chunk index is XORed with (row >> 1) & 3; LBO is
// unused for swizzled K-major operands.  One tcgen05.mma consumes K = 8 tf32 = 32 bytes of every row.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__device__ __forceinline__ int swz_off(int r, int c) { return r * kRowBytes + ((c ^ ((r >> 1) & 3)) << 4); }
// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32, a=b=TF32, both K-major, N>>3, M>>4
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// round-to-nearest (ties away) fp32 -> tf32 kept in a b32: add half a tf32 ulp to the magnitude bits and
// clear the 13 low mantissa bits.  Same result as cvt.rna.tf32.f32 for every finite input (a carry into
// the exponent is the correct rounding; the largest finite values round to inf), inf and NaN pass through;
// two integer instructions instead of the 4-5 ptxas emits for the guarded cvt.
__device__ __forceinline__ uint32_t rna_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }

// ---- producer: global -> registers -> split -> swizzled shared memory ----------------------------------
// Per-thread view of one operand tile.  Everything that does not change along K (row pointers, row
// validity, swizzled shared-memory offsets) is computed once; a K block costs PASSES vector loads, the
// hi/lo split and PASSES x 2 128-bit shared stores (the first version recomputed 64-bit addresses and
// bounds for every chunk: ~270 SASS instructions per block and warp, more than the 528 tensor-core cycles
// the block takes).
//   K-contiguous operand  (elem(r,k) = p[r*ld + k]): chunk c = tid & 3 of rows (tid >> 2) + 64 * pass.
//   MN-contiguous operand (elem(r,k) = p[k*ld + r], transposed on the fly): row lane + 32 * (warp >> 2)
//   + 64 * pass, k = 4 * (warp & 3) + {0..3}; every load of a warp is one coalesced 128-byte segment.
template <int PASSES>
struct Lane {
  const float* ptr[PASSES];  // first element of this thread's chunk in K block 0
  int off[PASSES];           // swizzled byte offset inside the stage (hi and lo share it)
  uint32_t ok;               // bit ps: the row of item ps exists in the matrix (load it)
  uint32_t st;               // bit ps: the row of item ps exists in the tile (store it, zeros if !ok)
  int64_t ld;                // leading dimension (MN-contiguous operands step K by ld)
  int kpos;                  // k offset of this thread's chunk inside a block
  int vec;                   // widest aligned vector along the contiguous dimension
};

// Thread -> (row, 16-byte K chunk) items of a tile.
//   KC (K-contiguous):  chunk c = tid & 3, rows (tid >> 2) + 128 * ps.
//   MN (MN-contiguous, transposed on the fly): chunk c = warp & 3 (k = 4c..4c+3); items come in ROW PAIRS
//   (2 * lane, 2 * lane + 1) + 64 * (warp >> 2) + 256 * (ps / 2), so that the four k rows of a pair are four
//   coalesced 64-bit copies (256 contiguous bytes per warp) and a register transpose yields both chunks.
template <int PASSES, bool KC>
__device__ __forceinline__ void lane_init(Lane<PASSES>& L, const Operand& op, int r0, int rows, int tile_rows,
                                          int k_begin) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  L.ok = 0;
  L.st = 0;
  L.ld = op.ld;
  L.vec = op.vec;
  const int c = KC ? (threadIdx.x & 3) : (warp & 3);
  L.kpos = 4 * c;
#pragma unroll
  for (int ps = 0; ps < PASSES; ++ps) {
    const int rl = KC ? (threadIdx.x >> 2) + 128 * ps : 2 * lane + (ps & 1) + 64 * (warp >> 2) + 256 * (ps >> 1);
    const bool ok = rl < tile_rows && r0 + rl < rows;
    L.ok |= (ok ? 1u : 0u) << ps;
    L.st |= (rl < tile_rows ? 1u : 0u) << ps;
    const int r = ok ? r0 + rl : 0;
    L.ptr[ps] = KC ? op.p + (int64_t)r * op.ld + k_begin + L.kpos : op.p + (int64_t)(k_begin + L.kpos) * op.ld + r;
    L.off[ps] = swz_off(rl, c);
  }
}

// Issue the asynchronous copies of one K block into this thread's raw slots.  base[ps]: the thread's chunk at the
// start of the block; krem: elements of the segment left from there (<= 0: nothing, everything zero-filled);
// raw: shared address of the thread's first slot, item ps lives at raw + ps * kProducerThreads * 16.
template <int PASSES, bool KC, int VEC>
__device__ __forceinline__ void lane_issue(const Lane<PASSES>& L, const float* const (&base)[PASSES], int krem,
                                           uint32_t raw) {
  constexpr uint32_t kItem = kProducerThreads * 16;
  if constexpr (KC) {
    const int valid = max(0, min(4, krem - L.kpos));             // elements of this chunk inside the segment
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const uint32_t nbytes = ((L.ok >> ps) & 1u) ? 4u * valid : 0u;
      const uint32_t dst = raw + ps * kItem;
      const float* src = base[ps];
      if constexpr (VEC == 4) {
        cp_async16(dst, src, nbytes);
      } else if constexpr (VEC == 2) {
        cp_async8(dst, src, min(nbytes, 8u));
        cp_async8(dst + 8, src + 2, nbytes > 8u ? nbytes - 8u : 0u);
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) cp_async4(dst + 4 * u, src + u, nbytes > 4u * u ? 4u : 0u);
      }
    }
  } else {
    static_assert(PASSES % 2 == 0, "MN-contiguous tiles are loaded in row pairs");
#pragma unroll
    for (int pp = 0; pp < PASSES; pp += 2) {
      const bool ok0 = (L.ok >> pp) & 1u, ok1 = (L.ok >> (pp + 1)) & 1u;
      const float* src = base[pp];                                // row 2*lane; its pair is the next float
#pragma unroll
      for (int j = 0; j < 4; ++j) {                               // k = kpos + j: one 8-byte (row pair) piece
        const bool kin = L.kpos + j < krem;
        const uint32_t dst = raw + (pp + (j >> 1)) * kItem + (j & 1) * 8;
        const float* sj = src + (int64_t)j * L.ld;
        if (L.vec >= 2) {
          cp_async8(dst, sj, kin ? (ok0 ? (ok1 ? 8u : 4u) : 0u) : 0u);
        } else {
          cp_async4(dst, sj, kin && ok0 ? 4u : 0u);
          cp_async4(dst + 4, ok1 ? sj + 1 : sj, kin && ok1 ? 4u : 0u);
        }
      }
    }
  }
}

// raw slots -> hi/lo split -> swizzled UMMA tiles
template <int PASSES, bool KC>
__device__ __forceinline__ void lane_convert(const Lane<PASSES>& L, const uint8_t* raw, uint8_t* hi, uint8_t* lo) {
  constexpr int kItem = kProducerThreads * 16;
  auto put = [&](int ps, const float4& v) {
    if ((L.st >> ps) & 1u) {   // rows of the tile that exist in shared memory (zeros beyond the matrix)
      uint4 h, l;
      h.x = rna_tf32(v.x); h.y = rna_tf32(v.y); h.z = rna_tf32(v.z); h.w = rna_tf32(v.w);
      // lo is stored as the exact fp32 residual: the tensor core reads only its 19 high bits, i.e. truncates it
      // to TF32 itself (|error| <= 2^-22 |x|, sign independent of x: same class as the dropped lo*lo term);
      // rounding it here cost 8 of the 20 ALU instructions per 16-byte chunk of a conversion-bound producer
      l.x = __float_as_uint(v.x - __uint_as_float(h.x));
      l.y = __float_as_uint(v.y - __uint_as_float(h.y));
      l.z = __float_as_uint(v.z - __uint_as_float(h.z));
      l.w = __float_as_uint(v.w - __uint_as_float(h.w));
      *reinterpret_cast<uint4*>(hi + L.off[ps]) = h;
      *reinterpret_cast<uint4*>(lo + L.off[ps]) = l;
    }
  };
  if constexpr (KC) {
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) put(ps, *reinterpret_cast<const float4*>(raw + ps * kItem));
  } else {
#pragma unroll
    for (int pp = 0; pp < PASSES; pp += 2) {
      const float4 p01 = *reinterpret_cast<const float4*>(raw + pp * kItem);         // (k0: r,r+1) (k1: r,r+1)
      const float4 p23 = *reinterpret_cast<const float4*>(raw + (pp + 1) * kItem);   // (k2: r,r+1) (k3: r,r+1)
      put(pp, make_float4(p01.x, p01.z, p23.x, p23.z));
      put(pp + 1, make_float4(p01.y, p01.w, p23.y, p23.w));
    }
  }
}

// Register-staged variant for MN-contiguous tiles (wgrad): row pairs, four coalesced 64-bit loads each.
template <int PASSES>
__device__ __forceinline__ void lane_load_mn(const Lane<PASSES>& L, const float* const (&base)[PASSES], int krem,
                                             float4 (&reg)[PASSES]) {
  static_assert(PASSES % 2 == 0, "MN-contiguous tiles are loaded in row pairs");
  const bool full = krem >= BK;
#pragma unroll
  for (int pp = 0; pp < PASSES; pp += 2) {
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    const bool ok0 = (L.ok >> pp) & 1u, ok1 = (L.ok >> (pp + 1)) & 1u;
    const float* src = base[pp];
    if (full && ok0 && ok1 && L.vec >= 2) {
      const float2 a = __ldg(reinterpret_cast<const float2*>(src));
      const float2 b = __ldg(reinterpret_cast<const float2*>(src + L.ld));
      const float2 c = __ldg(reinterpret_cast<const float2*>(src + 2 * L.ld));
      const float2 d = __ldg(reinterpret_cast<const float2*>(src + 3 * L.ld));
      v0 = make_float4(a.x, b.x, c.x, d.x);
      v1 = make_float4(a.y, b.y, c.y, d.y);
    } else {
      if (ok0) {
        if (L.kpos < krem) v0.x = __ldg(src);
        if (L.kpos + 1 < krem) v0.y = __ldg(src + L.ld);
        if (L.kpos + 2 < krem) v0.z = __ldg(src + 2 * L.ld);
        if (L.kpos + 3 < krem) v0.w = __ldg(src + 3 * L.ld);
      }
      if (ok1) {
        if (L.kpos < krem) v1.x = __ldg(src + 1);
        if (L.kpos + 1 < krem) v1.y = __ldg(src + 1 + L.ld);
        if (L.kpos + 2 < krem) v1.z = __ldg(src + 1 + 2 * L.ld);
        if (L.kpos + 3 < krem) v1.w = __ldg(src + 1 + 3 * L.ld);
      }
    }
    reg[pp] = v0;
    reg[pp + 1] = v1;
  }
}
template <int PASSES>
__device__ __forceinline__ void lane_put(const Lane<PASSES>& L, const float4 (&reg)[PASSES], uint8_t* hi, uint8_t* lo) {
#pragma unroll
  for (int ps = 0; ps < PASSES; ++ps) {
    if ((L.st >> ps) & 1u) {
      const float4 v = reg[ps];
      uint4 h, l;
      h.x = rna_tf32(v.x); h.y = rna_tf32(v.y); h.z = rna_tf32(v.z); h.w = rna_tf32(v.w);
      // lo is stored as the exact fp32 residual: the tensor core reads only its 19 high bits, i.e. truncates it
      // to TF32 itself (|error| <= 2^-22 |x|, sign independent of x: same class as the dropped lo*lo term);
      // rounding it here cost 8 of the 20 ALU instructions per 16-byte chunk of a conversion-bound producer
      l.x = __float_as_uint(v.x - __uint_as_float(h.x));
      l.y = __float_as_uint(v.y - __uint_as_float(h.y));
      l.z = __float_as_uint(v.z - __uint_as_float(h.z));
      l.w = __float_as_uint(v.w - __uint_as_float(h.w));
      *reinterpret_cast<uint4*>(hi + L.off[ps]) = h;
      *reinterpret_cast<uint4*>(lo + L.off[ps]) = l;
    }
  }
}

// Weight packer: writes, for every (N tile nt, K block kb), the exact shared-memory image
// [B_hi (BN x 64 B, SWIZZLE_64B) | B_lo] at  out + (nt * nkb + kb) * 2 * BN * 64.
__global__ void __launch_bounds__(256)
tc_pack_b_kernel(Operand b0, int K0, Operand b1, int K1, int N, int BN, uint8_t* __restrict__ out) {
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
    uint4 h, l;
    h.x = rna_tf32(v[0]); h.y = rna_tf32(v[1]); h.z = rna_tf32(v[2]); h.w = rna_tf32(v[3]);
    l.x = rna_tf32(v[0] - __uint_as_float(h.x));
    l.y = rna_tf32(v[1] - __uint_as_float(h.y));
    l.z = rna_tf32(v[2] - __uint_as_float(h.z));
    l.w = rna_tf32(v[3] - __uint_as_float(h.w));
    uint8_t* tile = out + ((int64_t)nt * nkb + kb) * (2 * (int64_t)BN * kRowBytes);
    const int off = swz_off(r, c);
    *reinterpret_cast<uint4*>(tile + off) = h;
    *reinterpret_cast<uint4*>(tile + (int64_t)BN * kRowBytes + off) = l;
  }
}

// C[m][n] = sum over segments, k of A(m,k) * B(k,n)  (+ bias[n]);  blockIdx.z = K split of segment 0.
// B_PACKED: the B operand comes pre-split / pre-swizzled from tc_pack_b_kernel (`packed_b`).
// CL > 1 (B_PACKED only): thread-block cluster of CL CTAs along M that share one N tile.  Every CTA
// bulk-copies 1/CL of each weight stage and MULTICASTS it to all CL CTAs, so the weights cross the L2
// once per cluster instead of once per CTA (with CL = 1 this GEMM saturates L2 bandwidth: ~30 KB of
// operands per 540-cycle K block and SM, 2/3 of it the re-read weights).  A stage may only be overwritten
// once the MMAs of ALL cluster CTAs have retired: tcgen05.commit multicasts its arrival to every CTA's
// empty barrier (count CL).
// B_PACKED kernels take a K-contiguous A whose vector width VEC is a compile-time constant; the
// non-packed kernel is wgrad: A and B both MN-contiguous (VEC unused).
template <int BN, bool B_PACKED, int CL, int VEC>
__global__ void __launch_bounds__(kThreads, Cfg<BN, B_PACKED>::kCtasPerSm)
tc_gemm_kernel(Segment s0, Segment s1, const uint8_t* __restrict__ packed_b, int M, int N, float* __restrict__ c,
               int64_t ldc, int c_vec, const float* __restrict__ bias, int relu, int k_per_split, int64_t split_stride,
               int dbg /* timing experiments only: 1 = no A loads, 2 = no B copies, 4 = no MMAs */) {
  using C = Cfg<BN, B_PACKED>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // layout: [pipeline stages][raw cp.async ring][barriers]; the epilogue reuses stages + ring as staging tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes + C::kDepth * C::kRawBytes);
  // bars[0..S) full, bars[S..2S) empty, bars[2S] accumulator ready, then the TMEM base address
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 1);
  uint8_t* raw_ring = smem + C::kStages * C::kStageBytes;         // [kDepth][items][256 threads][16 B]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int k_lo = 0, k_hi = s0.K;
  if (gridDim.z > 1) {
    k_lo = blockIdx.z * k_per_split;
    k_hi = min(s0.K, k_lo + k_per_split);
    c += (int64_t)blockIdx.z * split_stride;
  }
  const int nb0 = k_hi > k_lo ? (k_hi - k_lo + BK - 1) / BK : 0;
  const int nb1 = (s1.K + BK - 1) / BK;
  const int nb = nb0 + nb1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(smem_u32(bars + s), B_PACKED ? kProducerWarps + 1 : ((dbg & 32) ? kProducerWarps : kProducerWarps / 4));
      mbar_init(smem_u32(bars + C::kStages + s), CL);
    }
    mbar_init(smem_u32(bars + 2 * C::kStages), 1);
    fence_barrier_init();
  }
  if (threadIdx.x == 32 && m0 < M) {
    // stream this CTA's activation rows into L2 ahead of the K loop (K-contiguous operands only)
    const int rows = min(BM, M - m0);
    if (s0.a.k_contig && s0.a.vec >= 2) {
      const uint32_t bytes = (uint32_t)(((int64_t)(rows - 1) * s0.a.ld + s0.K) * 4) & ~15u;
      const uintptr_t p = (uintptr_t)(s0.a.p + (int64_t)m0 * s0.a.ld) & ~(uintptr_t)15;
      if (bytes >= 16) bulk_prefetch_l2((const void*)p, bytes);
    }
    if (s1.K > 0 && s1.a.k_contig && s1.a.vec >= 2) {
      const uint32_t bytes = (uint32_t)(((int64_t)(rows - 1) * s1.a.ld + s1.K) * 4) & ~15u;
      const uintptr_t p = (uintptr_t)(s1.a.p + (int64_t)m0 * s1.a.ld) & ~(uintptr_t)15;
      if (bytes >= 16) bulk_prefetch_l2((const void*)p, bytes);
    }
  }
  if (warp == kProducerWarps) tmem_alloc(smem_u32(tmem_slot), C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();     // peers' barriers are initialised before any remote arrival
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint16_t kClusterMask = (uint16_t)((1u << CL) - 1u);

  if (warp < kProducerWarps) {
    // ================= producers =================
    constexpr int PA = C::kItemsA;
    constexpr int PB = 2;
    constexpr bool KC = B_PACKED;
    Lane<PA> la;
    Lane<PB> lb;
    const float* pa1[PA];                                          // second segment (same thread mapping)
    lane_init<PA, KC>(la, s0.a, m0, M, BM, k_lo);
    if constexpr (B_PACKED) {
#pragma unroll
      for (int ps = 0; ps < PA; ++ps) {
        const int rl = (threadIdx.x >> 2) + 128 * ps;
        const int r = (m0 + rl < M) ? m0 + rl : 0;
        pa1[ps] = s1.K > 0 ? s1.a.p + (int64_t)r * s1.a.ld + la.kpos : la.ptr[ps];
      }
    } else {
      lane_init<PB, false>(lb, s0.b, n0, N, BN, k_lo);
    }
    if constexpr (B_PACKED) {
      uint8_t* my_raw = raw_ring + threadIdx.x * 16;
      auto issue = [&](int it) {                                    // cp.async the A chunk of K block `it`
        if (it < nb && !(dbg & 1)) {
          const uint32_t raw = smem_u32(my_raw + (it % C::kDepth) * C::kRawBytes);
          const float* base[PA];
          if (it < nb0) {
#pragma unroll
            for (int ps = 0; ps < PA; ++ps) base[ps] = la.ptr[ps] + (int64_t)it * BK;
            lane_issue<PA, true, VEC>(la, base, k_hi - k_lo - it * BK, raw);
          } else {
#pragma unroll
            for (int ps = 0; ps < PA; ++ps) base[ps] = pa1[ps] + (it - nb0) * BK;
            lane_issue<PA, true, VEC>(la, base, s1.K - (it - nb0) * BK, raw);
          }
        }
        cp_async_commit();                                          // (empty groups keep the count uniform)
      };
      for (int d = 0; d < C::kDepth; ++d) issue(d);
      for (int it = 0; it < nb; ++it) {
        const int s = it % C::kStages;
        const uint32_t ph = (uint32_t)(it / C::kStages) & 1u;
        cp_async_wait<C::kDepth - 1>();                             // this thread's copies of block `it` landed
        mbar_wait(smem_u32(bars + C::kStages + s), ph ^ 1u);        // slot free (MMAs that read it retired)
        uint8_t* st = smem + s * C::kStageBytes;
        lane_convert<PA, true>(la, my_raw + (it % C::kDepth) * C::kRawBytes, st, st + C::kABytes);
        issue(it + C::kDepth);                                      // refill the raw slot just consumed
        if (!(dbg & 8)) fence_proxy_async();                        // generic-proxy writes -> async proxy (UMMA)
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(bars + s));             // one arrival per producer warp
      }
      cp_async_wait<0>();
    } else {
      // wgrad: both operands MN-contiguous, register double buffer
      float4 ra[2][PA], rb[2][PB];
      auto fetch = [&](int it, float4 (&fa)[PA], float4 (&fb)[PB]) {
        if (it >= nb) return;
        const float* ba[PA];
        const float* bb[PB];
#pragma unroll
        for (int ps = 0; ps < PA; ++ps) ba[ps] = la.ptr[ps] + (int64_t)it * BK * la.ld;
#pragma unroll
        for (int ps = 0; ps < PB; ++ps) bb[ps] = lb.ptr[ps] + (int64_t)it * BK * lb.ld;
        lane_load_mn<PA>(la, ba, k_hi - k_lo - it * BK, fa);
        lane_load_mn<PB>(lb, bb, k_hi - k_lo - it * BK, fb);
      };
      auto stash = [&](int it, float4 (&fa)[PA], float4 (&fb)[PB]) {
        const int s = it % C::kStages;
        const uint32_t ph = (uint32_t)(it / C::kStages) & 1u;
        mbar_wait(smem_u32(bars + C::kStages + s), ph ^ 1u);
        uint8_t* st = smem + s * C::kStageBytes;
        lane_put<PA>(la, fa, st, st + C::kABytes);
        lane_put<PB>(lb, fb, st + 2 * C::kABytes, st + 2 * C::kABytes + C::kBBytes);
        fetch(it + 2, fa, fb);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(bars + s));
      };
      if (dbg & 32) {                                               // all 16 warps on every block (first version)
        fetch(0, ra[0], rb[0]);
        fetch(1, ra[1], rb[1]);
        for (int it = 0; it < nb; it += 2) {
          stash(it, ra[0], rb[0]);
          if (it + 1 < nb) stash(it + 1, ra[1], rb[1]);
        }
      } else {
        // ---- four producer groups of 4 warps, group g owns K blocks g, g+4, ... ----
        // The register-staged loads of a thread alias the six scoreboard slots, so whatever the software prefetch
        // depth a thread waits for its newest load: one K block per memory round trip (measured 2300 cycles per
        // block with all 16 warps on every block, 4x the tensor-core time; a cp.async ring with real depth 3 was
        // slower still: 16 eight-byte LDGSTS per thread and block).  Groups that own every G-th block overlap G
        // round trips: 2 groups 0.415 -> 0.247 ms, 4 groups (one block in flight each) see DESIGN.md.
        // kStages >= groups, so a group never waits more than one barrier phase behind (no parity aliasing).
        constexpr int G = 4;
        static_assert(C::kStages >= G, "producer groups must not outnumber the pipeline stages");
        constexpr int PG_A = 2 * (BM / 64), PG_B = 2 * ((BN + 63) / 64);   // row pairs: 64 rows per pass pair
        const int grp = warp >> 2, wg = warp & 3;
        Lane<PG_A> ga;
        Lane<PG_B> gb;
        auto ginit = [&](auto& L, const Operand& op, int r0, int rows, int tile_rows, int passes) {
          L.ok = 0; L.st = 0; L.ld = op.ld; L.vec = op.vec;
          L.kpos = 4 * wg;
          for (int ps = 0; ps < passes; ++ps) {
            const int rl = 2 * lane + (ps & 1) + 64 * (ps >> 1);
            const bool ok = rl < tile_rows && r0 + rl < rows;
            L.ok |= (ok ? 1u : 0u) << ps;
            L.st |= (rl < tile_rows ? 1u : 0u) << ps;
            const int r = ok ? r0 + rl : 0;
            L.ptr[ps] = op.p + (int64_t)(k_lo + L.kpos) * op.ld + r;
            L.off[ps] = swz_off(rl, wg);
          }
        };
        ginit(ga, s0.a, m0, M, BM, PG_A);
        ginit(gb, s0.b, n0, N, BN, PG_B);
        float4 fa[PG_A], fb[PG_B];
        auto gfetch = [&](int it) {
          if (it >= nb) return;
          const float* ba[PG_A];
          const float* bb[PG_B];
#pragma unroll
          for (int ps = 0; ps < PG_A; ++ps) ba[ps] = ga.ptr[ps] + (int64_t)it * BK * ga.ld;
#pragma unroll
          for (int ps = 0; ps < PG_B; ++ps) bb[ps] = gb.ptr[ps] + (int64_t)it * BK * gb.ld;
          lane_load_mn<PG_A>(ga, ba, k_hi - k_lo - it * BK, fa);
          lane_load_mn<PG_B>(gb, bb, k_hi - k_lo - it * BK, fb);
        };
        gfetch(grp);
        for (int it = grp; it < nb; it += G) {
          const int s = it % C::kStages;
          const uint32_t ph = (uint32_t)(it / C::kStages) & 1u;
          mbar_wait(smem_u32(bars + C::kStages + s), ph ^ 1u);
          uint8_t* st = smem + s * C::kStageBytes;
          lane_put<PG_A>(ga, fa, st, st + C::kABytes);
          lane_put<PG_B>(gb, fb, st + 2 * C::kABytes, st + 2 * C::kABytes + C::kBBytes);
          gfetch(it + G);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(bars + s));
        }
      }
    }
    // ================= epilogue =================
    // TMEM -> registers (4 column chunks per tcgen05.wait) -> (+ correction tile, + bias) -> shared memory
    // (the pipeline stages are free now) -> fully coalesced row stores.
    mbar_wait(smem_u32(bars + 2 * C::kStages), 0);
    tc_fence_after();
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = warp >> 2;                // column half
    constexpr int kHalf = BN / 2;
    static_assert(kHalf % 8 == 0, "BN / 2 must be a multiple of 8");
    constexpr int kLdS = BN + 4;               // padded staging row (floats): conflict-free 128-bit writes
    static_assert(BM * kLdS * 4 <= C::kStages * C::kStageBytes + C::kDepth * C::kRawBytes,
                  "staging tile must fit in the pipeline stages + raw ring");
    asm volatile("bar.sync 1, %0;" ::"n"(kProducerThreads) : "memory");   // every producer is done with its raw slots
    float* stage_c = reinterpret_cast<float*>(smem);
    const int row_l = q * 32 + lane;
    for (int cc = 0; warp < 8 && cc < kHalf; cc += 32) {   // warps 0-7 drain TMEM (lane quarter x column half)
      uint32_t rm[4][8], rc[4][8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (cc + 8 * j < kHalf) {
          const uint32_t col = (uint32_t)(half * kHalf + cc + 8 * j);
          tmem_ld8_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + col, rm[j]);
          tmem_ld8_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)C::kCorrCol + col, rc[j]);
        }
      }
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (cc + 8 * j < kHalf) {
          const int nl = half * kHalf + cc + 8 * j;
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            v[u] = nb == 0 ? 0.f : __uint_as_float(rm[j][u]) + __uint_as_float(rc[j][u]);
            if (bias != nullptr && n0 + nl + u < N) v[u] += __ldg(bias + n0 + nl + u);
            if (relu) v[u] = v[u] <= 0.f ? 0.f : v[u];
          }
          float* d = stage_c + row_l * kLdS + nl;
          *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(d + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
    }
    tc_fence_before();
    asm volatile("bar.sync 1, %0;" ::"n"(kProducerThreads) : "memory");   // all producer / epilogue warps
    // coalesced write-out: warp w owns rows w, w+8, ...; lanes sweep the row
    const int ncols = min(BN, N - n0);
    for (int r = warp; r < BM; r += kProducerWarps) {
      const int m = m0 + r;
      if (m >= M) break;
      const float* srow = stage_c + r * kLdS;
      float* drow = c + (int64_t)m * ldc + n0;
      if (c_vec == 4) {
        for (int col = lane * 4; col < ncols; col += 128) {
          if (col + 4 <= ncols) {
            *reinterpret_cast<float4*>(drow + col) = *reinterpret_cast<const float4*>(srow + col);
          } else {
            for (int u = 0; col + u < ncols; ++u) drow[col + u] = srow[col + u];
          }
        }
      } else if (c_vec == 2) {
        for (int col = lane * 2; col < ncols; col += 64) {
          if (col + 2 <= ncols) *reinterpret_cast<float2*>(drow + col) = *reinterpret_cast<const float2*>(srow + col);
          else drow[col] = srow[col];
        }
      } else {
        for (int col = lane; col < ncols; col += 32) drow[col] = srow[col];
      }
    }
  } else if (warp == kProducerWarps) {
    if (lane == 0) {
      // ================= MMA issuer (one thread) =================
      constexpr uint32_t idesc = make_idesc(BN);
      for (int it = 0; it < nb; ++it) {
        const int s = it % C::kStages;
        const uint32_t ph = (uint32_t)(it / C::kStages) & 1u;
        mbar_wait(smem_u32(bars + s), ph);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + s * C::kStageBytes);
        const uint64_t a_hi = make_desc(st), a_lo = make_desc(st + C::kABytes);
        const uint64_t b_hi = make_desc(st + 2 * C::kABytes), b_lo = make_desc(st + 2 * C::kABytes + C::kBBytes);
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          if (dbg & 4) break;
          const uint64_t adv = (uint64_t)(k * 32 >> 4);             // 8 tf32 = 32 bytes along the swizzled row
          umma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, (it | k) != 0);
          umma_tf32(tmem_base + C::kCorrCol, a_lo + adv, b_hi + adv, idesc, (it | k) != 0);
          umma_tf32(tmem_base + C::kCorrCol, a_hi + adv, b_lo + adv, idesc, 1);
        }
        // frees the stage (in every cluster CTA) when these MMAs retire
        if constexpr (CL > 1) umma_commit_mcast(smem_u32(bars + C::kStages + s), kClusterMask);
        else umma_commit(smem_u32(bars + C::kStages + s));
      }
      umma_commit(smem_u32(bars + 2 * C::kStages));                 // accumulator complete
    }
  } else if (B_PACKED && lane == 0) {
    // ================= weight loader (one thread, TMA engine) =================
    const uint8_t* src = packed_b + (int64_t)blockIdx.x * nb * (2 * C::kBBytes);
    static_assert((2 * C::kBBytes / CL) % 16 == 0, "bulk copy slices must be multiples of 16 bytes");
    constexpr uint32_t kSlice = 2 * C::kBBytes / CL;
    const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
    for (int it = 0; it < nb; ++it) {
      const int s = it % C::kStages;
      const uint32_t ph = (uint32_t)(it / C::kStages) & 1u;
      mbar_wait(smem_u32(bars + C::kStages + s), ph ^ 1u);          // stage free in ALL cluster CTAs
      const uint32_t full = smem_u32(bars + s);
      if (dbg & 2) { mbar_arrive(full); continue; }
      mbar_arrive_expect_tx(full, 2 * C::kBBytes);                  // own barrier: all CL slices land here
      const uint32_t dst = smem_u32(smem + s * C::kStageBytes + 2 * C::kABytes) + rank * kSlice;
      const uint8_t* from = src + (int64_t)it * (2 * C::kBBytes) + rank * kSlice;
      if constexpr (CL > 1) bulk_g2s_mcast(dst, from, kSlice, full, kClusterMask);
      else bulk_g2s(dst, from, kSlice, full);
    }
  }
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();     // no CTA leaves while peers may still signal its barriers
  if (warp == kProducerWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}


// ---------------------------------------------------------------------------------------------------------
// Persistent variant of the packed-weights GEMM (forward and dgrad): one CTA per SM walks tiles t, t + grid, ...
// Time vs K of the one-tile-per-CTA kernel is a line with a ~0.17 ms intercept for 2042 tiles: ~12 us per tile of
// CTA launch, TMEM allocation, barrier set-up, pipeline fill (one DRAM round trip with nothing to do) and drain.
// Here the barriers, the TMEM allocation and the thread-to-chunk mapping live for the whole kernel, and the
// producers issue the cp.async copies of the NEXT tile's first blocks BEFORE they run the epilogue of the
// current one, so the next K loop starts on data that is already in the raw ring.
// (Dedicated epilogue warps and a second TMEM accumulator set were tried first and were slower: four warps
// draining 32-column slabs held the accumulator longer than the 16 producer warps need for the whole tile.)
// Extra barriers: acc_full (MMA -> epilogue), tmem_free (8 draining warps -> MMA), epi_done (16 warps -> weight
// stream: the epilogue's staging tile lives in the pipeline stages the bulk copies write).
// ---------------------------------------------------------------------------------------------------------
template <int BN, int VEC>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_persistent_kernel(Segment s0, Segment s1, const uint8_t* __restrict__ packed_b, int M, int N,
                          float* __restrict__ c, int64_t ldc, int c_vec, const float* __restrict__ bias, int relu) {
  using C = Cfg<BN, true>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes + C::kDepth * C::kRawBytes);
  // bars[0..S) full, [S..2S) empty, [2S] acc_full, [2S+1] tmem_free, [2S+2] epi_done, then the TMEM base address
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 3);
  const uint32_t bar_acc_full = smem_u32(bars + 2 * C::kStages);
  const uint32_t bar_tmem_free = smem_u32(bars + 2 * C::kStages + 1);
  const uint32_t bar_epi_done = smem_u32(bars + 2 * C::kStages + 2);
  uint8_t* raw_ring = smem + C::kStages * C::kStageBytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb0 = (s0.K + BK - 1) / BK, nb1 = (s1.K + BK - 1) / BK, nb = nb0 + nb1;
  const int ntn = (N + BN - 1) / BN, ntm = (M + BM - 1) / BM;
  const int ntiles = ntn * ntm;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(smem_u32(bars + s), kProducerWarps + 1);
      mbar_init(smem_u32(bars + C::kStages + s), 1);
    }
    mbar_init(bar_acc_full, 1);
    mbar_init(bar_tmem_free, 8);
    mbar_init(bar_epi_done, kProducerWarps);
    fence_barrier_init();
  }
  if (warp == kProducerWarps) tmem_alloc(smem_u32(tmem_slot), C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto prefetch_rows = [&](int tile) {          // this tile's activation rows -> L2 (one contiguous span per segment)
    if (tile >= ntiles) return;
    const int m0 = (tile / ntn) * BM;
    const int rows = min(BM, M - m0);
    if (s0.a.vec >= 2) {
      const uint32_t bytes = (uint32_t)(((int64_t)(rows - 1) * s0.a.ld + s0.K) * 4) & ~15u;
      const uintptr_t p = (uintptr_t)(s0.a.p + (int64_t)m0 * s0.a.ld) & ~(uintptr_t)15;
      if (bytes >= 16) bulk_prefetch_l2((const void*)p, bytes);
    }
    if (s1.K > 0 && s1.a.vec >= 2) {
      const uint32_t bytes = (uint32_t)(((int64_t)(rows - 1) * s1.a.ld + s1.K) * 4) & ~15u;
      const uintptr_t p = (uintptr_t)(s1.a.p + (int64_t)m0 * s1.a.ld) & ~(uintptr_t)15;
      if (bytes >= 16) bulk_prefetch_l2((const void*)p, bytes);
    }
  };

  if (warp < kProducerWarps) {
    // ================= producers (and epilogue) =================
    constexpr int PA = C::kItemsA;
    static_assert(PA == 1, "one 16-byte chunk per producer thread and K block");
    uint8_t* my_raw = raw_ring + threadIdx.x * 16;
    const int rl = threadIdx.x >> 2, cch = threadIdx.x & 3, kpos = 4 * cch;
    const int off = swz_off(rl, cch);
    const float* p0 = nullptr;
    const float* p1 = nullptr;
    bool rok = false;
    auto setup = [&](int tile) {                                    // row pointers of this thread's chunk
      const int m0 = (tile / ntn) * BM;
      rok = m0 + rl < M;
      const int r = rok ? m0 + rl : 0;
      p0 = s0.a.p + (int64_t)r * s0.a.ld + kpos;
      p1 = s1.K > 0 ? s1.a.p + (int64_t)r * s1.a.ld + kpos : p0;
    };
    auto issue = [&](int it) {                                      // cp.async the A chunk of K block `it`
      if (it < nb) {
        const bool first = it < nb0;
        const int krem = first ? s0.K - it * BK : s1.K - (it - nb0) * BK;
        const int valid = max(0, min(4, krem - kpos));
        const uint32_t nbytes = rok ? 4u * valid : 0u;
        const uint32_t dst = smem_u32(my_raw + (it % C::kDepth) * C::kRawBytes);
        const float* src = first ? p0 + (int64_t)it * BK : p1 + (it - nb0) * BK;
        if constexpr (VEC == 4) {
          cp_async16(dst, src, nbytes);
        } else if constexpr (VEC == 2) {
          cp_async8(dst, src, min(nbytes, 8u));
          cp_async8(dst + 8, src + 2, nbytes > 8u ? nbytes - 8u : 0u);
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) cp_async4(dst + 4 * u, src + u, nbytes > 4u * u ? 4u : 0u);
        }
      }
      cp_async_commit();
    };
    const int q = warp & 3, half = warp >> 2;
    constexpr int kHalf = BN / 2;
    constexpr int kLdS = BN + 4;
    static_assert(BM * kLdS * 4 <= C::kStages * C::kStageBytes, "staging tile must fit in the pipeline stages");
    float* stage_c = reinterpret_cast<float*>(smem);
    const int row_l = q * 32 + lane;

    uint32_t g = 0, tl = 0;
    int tile = blockIdx.x;
    if (tile < ntiles) {
      setup(tile);
      for (int d = 0; d < C::kDepth; ++d) issue(d);
    }
    for (; tile < ntiles; tile += gridDim.x, ++tl) {
      const int m0 = (tile / ntn) * BM, n0 = (tile % ntn) * BN;
      for (int it = 0; it < nb; ++it, ++g) {
        const int s = (int)(g % C::kStages);
        const uint32_t ph = (g / C::kStages) & 1u;
        cp_async_wait<C::kDepth - 1>();
        mbar_wait(smem_u32(bars + C::kStages + s), ph ^ 1u);
        uint8_t* hi = smem + s * C::kStageBytes;
        const float4 v = *reinterpret_cast<const float4*>(my_raw + (it % C::kDepth) * C::kRawBytes);
        uint4 h, l;
        h.x = rna_tf32(v.x); h.y = rna_tf32(v.y); h.z = rna_tf32(v.z); h.w = rna_tf32(v.w);
        l.x = __float_as_uint(v.x - __uint_as_float(h.x));          // exact residual, see lane_put
        l.y = __float_as_uint(v.y - __uint_as_float(h.y));
        l.z = __float_as_uint(v.z - __uint_as_float(h.z));
        l.w = __float_as_uint(v.w - __uint_as_float(h.w));
        *reinterpret_cast<uint4*>(hi + off) = h;
        *reinterpret_cast<uint4*>(hi + C::kABytes + off) = l;
        issue(it + C::kDepth);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(bars + s));
      }
      cp_async_wait<0>();
      // ---- next tile's first blocks go in flight now: they land during the epilogue ----
      const int next = tile + gridDim.x;
      if (next < ntiles) {
        setup(next);
        for (int d = 0; d < C::kDepth; ++d) issue(d);
      }
      // ---- epilogue: TMEM -> registers -> staging tile in the (idle) pipeline stages -> coalesced row stores ----
      mbar_wait(bar_acc_full, tl & 1u);
      tc_fence_after();
      for (int cc = 0; warp < 8 && cc < kHalf; cc += 32) {          // warps 0-7 drain TMEM (lane quarter x column half)
        uint32_t rm[4][8], rc[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (cc + 8 * j < kHalf) {
            const uint32_t col = (uint32_t)(half * kHalf + cc + 8 * j);
            tmem_ld8_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + col, rm[j]);
            tmem_ld8_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)C::kCorrCol + col, rc[j]);
          }
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (cc + 8 * j < kHalf) {
            const int nl = half * kHalf + cc + 8 * j;
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              v[u] = __uint_as_float(rm[j][u]) + __uint_as_float(rc[j][u]);
              if (bias != nullptr && n0 + nl + u < N) v[u] += __ldg(bias + n0 + nl + u);
              if (relu) v[u] = v[u] <= 0.f ? 0.f : v[u];
            }
            float* d = stage_c + row_l * kLdS + nl;
            *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(d + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
      }
      tc_fence_before();
      if (warp < 8) {                                               // accumulator drained: the MMA warp may reuse it
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tmem_free);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kProducerThreads) : "memory");   // staging tile complete
      const int ncols = min(BN, N - n0);
      for (int r = warp; r < BM; r += kProducerWarps) {
        const int m = m0 + r;
        if (m >= M) break;
        const float* srow = stage_c + r * kLdS;
        float* drow = c + (int64_t)m * ldc + n0;
        if (c_vec == 4) {
          for (int col = lane * 4; col < ncols; col += 128) {
            if (col + 4 <= ncols) {
              *reinterpret_cast<float4*>(drow + col) = *reinterpret_cast<const float4*>(srow + col);
            } else {
              for (int u = 0; col + u < ncols; ++u) drow[col + u] = srow[col + u];
            }
          }
        } else if (c_vec == 2) {
          for (int col = lane * 2; col < ncols; col += 64) {
            if (col + 2 <= ncols) *reinterpret_cast<float2*>(drow + col) = *reinterpret_cast<const float2*>(srow + col);
            else drow[col] = srow[col];
          }
        } else {
          for (int col = lane; col < ncols; col += 32) drow[col] = srow[col];
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_epi_done);                     // this warp no longer reads the staging tile
      asm volatile("bar.sync 1, %0;" ::"n"(kProducerThreads) : "memory");   // ... and nobody writes stages before that
    }
  } else if (warp == kProducerWarps) {
    if (lane == 0) {
      // ================= MMA issuer (one thread) =================
      constexpr uint32_t idesc = make_idesc(BN);
      uint32_t g = 0, tl = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
        mbar_wait(bar_tmem_free, (tl & 1u) ^ 1u);                   // previous tile drained (first tile: passes)
        tc_fence_after();
        for (int it = 0; it < nb; ++it, ++g) {
          const int s = (int)(g % C::kStages);
          const uint32_t ph = (g / C::kStages) & 1u;
          mbar_wait(smem_u32(bars + s), ph);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + s * C::kStageBytes);
          const uint64_t a_hi = make_desc(st), a_lo = make_desc(st + C::kABytes);
          const uint64_t b_hi = make_desc(st + 2 * C::kABytes), b_lo = make_desc(st + 2 * C::kABytes + C::kBBytes);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            umma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, (it | k) != 0);
            umma_tf32(tmem_base + C::kCorrCol, a_lo + adv, b_hi + adv, idesc, (it | k) != 0);
            umma_tf32(tmem_base + C::kCorrCol, a_hi + adv, b_lo + adv, idesc, 1);
          }
          umma_commit(smem_u32(bars + C::kStages + s));
        }
        umma_commit(bar_acc_full);
      }
    }
  } else if (lane == 0) {
    // ================= weight stream + L2 prefetch (one thread, TMA engine) =================
    uint32_t g = 0, tl = 0;
    prefetch_rows(blockIdx.x);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
      prefetch_rows(tile + gridDim.x);
      mbar_wait(bar_epi_done, (tl & 1u) ^ 1u);                      // staging tile of the previous tile is dead
      const uint8_t* src = packed_b + (int64_t)(tile % ntn) * nb * (2 * C::kBBytes);
      for (int it = 0; it < nb; ++it, ++g) {
        const int s = (int)(g % C::kStages);
        const uint32_t ph = (g / C::kStages) & 1u;
        mbar_wait(smem_u32(bars + C::kStages + s), ph ^ 1u);
        const uint32_t full = smem_u32(bars + s);
        mbar_arrive_expect_tx(full, 2 * C::kBBytes);
        bulk_g2s(smem_u32(smem + s * C::kStageBytes + 2 * C::kABytes), src + (int64_t)it * (2 * C::kBBytes),
                 2 * C::kBBytes, full);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kProducerWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

}  // namespace tc
}  // namespace mgs
