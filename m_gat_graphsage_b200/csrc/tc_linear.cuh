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
// Structure (one 128 x BN output tile per CTA, 320 threads):
//   warps 0-7  producers: coalesced LDG of the fp32 A / B tiles straight from global memory (activations
//              have 1400-byte rows: not TMA-able without a padded copy), hi/lo split in registers,
//              conflict-free 128-bit STS into the canonical K-major SWIZZLE_128B layout (both operand
//              orientations are transposed on the fly, so shared memory is always K-major),
//              fence.proxy.async, mbarrier arrive;  after the K loop the same warps are the epilogue:
//              tcgen05.ld TMEM -> registers -> (+bias) -> global.
//   warp 8     lane 0 issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8) x 4 k-steps x 3
//              split products per 32-float K block, tcgen05.commit releases the stage;  also owns the
//              TMEM allocation.
//   warp 9     (B_PACKED) lane 0 streams the weight operand with cp.async.bulk (TMA engine, 1-D):
//              weights are tiny (<= 4 MB), so tc_pack_b_kernel splits and swizzles them ONCE per call
//              into exactly the shared-memory image of every (N tile, K block); a stage's B_hi|B_lo is
//              then one 32-64 KB bulk copy completing on the stage's full barrier (complete_tx).
//              For wgrad both operands are activation-sized and both go through the producer warps.
// Pipeline: `stages` shared-memory stages with full/empty mbarriers; the producers keep the global
// loads of blocks i+1 and i+2 in flight (two register buffers) while block i is converted.
#pragma once

#include "common.cuh"

namespace mgs {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;                 // floats per K block = one 128-byte swizzle row
constexpr int kProducerWarps = 8;
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

template <int BN> struct Cfg {
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128");
  static constexpr int kStages = BN <= 128 ? 3 : 2;
  static constexpr int kCorrCol = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;  // correction tile
  static constexpr int kTmemCols = 2 * kCorrCol;
  static constexpr int kABytes = BM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = 2 * (kABytes + kBBytes);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /* alignment slack */ + 256 /* barriers */;
  static constexpr int kPassesA = BM / 32;
  static constexpr int kPassesB = (BN + 31) / 32;
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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format, cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4
//   [46,48) version = 1 | [49,52) base offset = 0 | [61,64) layout type = 2 (SWIZZLE_128B)
// Rows are 128 bytes, 8-row swizzle atoms are 1024 bytes apart (SBO); LBO is unused for swizzled K-major.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32, a=b=TF32, both K-major, N>>3, M>>4
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ uint32_t rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// ---- producer: global -> registers -------------------------------------------------------------------
// K-contiguous operand: thread owns 16-byte chunk c = tid & 7 of rows (tid >> 3) + 32 * pass.
template <int PASSES>
__device__ __forceinline__ void load_kc(const Operand& op, int r0, int rows, int tile_rows, int k0, int kend,
                                        float4 (&reg)[PASSES]) {
  const int c = threadIdx.x & 7;
  const int k = k0 + 4 * c;
#pragma unroll
  for (int ps = 0; ps < PASSES; ++ps) {
    const int rl = (threadIdx.x >> 3) + 32 * ps;
    const int r = r0 + rl;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rl < tile_rows && r < rows && k < kend) {
      const float* src = op.p + (int64_t)r * op.ld + k;
      if (op.vec == 4 && k + 4 <= kend) {
        v = __ldg(reinterpret_cast<const float4*>(src));
      } else if (op.vec >= 2 && k + 4 <= kend) {
        const float2 a = __ldg(reinterpret_cast<const float2*>(src));
        const float2 b = __ldg(reinterpret_cast<const float2*>(src) + 1);
        v = make_float4(a.x, a.y, b.x, b.y);
      } else {
        v.x = __ldg(src);
        if (k + 1 < kend) v.y = __ldg(src + 1);
        if (k + 2 < kend) v.z = __ldg(src + 2);
        if (k + 3 < kend) v.w = __ldg(src + 3);
      }
    }
    reg[ps] = v;
  }
}
// MN-contiguous operand: thread owns row lane + 32 * pass, k = 4 * warp + {0..3} (transposed on the fly)
template <int PASSES>
__device__ __forceinline__ void load_mn(const Operand& op, int r0, int rows, int tile_rows, int k0, int kend,
                                        float4 (&reg)[PASSES]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = k0 + 4 * warp;
#pragma unroll
  for (int ps = 0; ps < PASSES; ++ps) {
    const int rl = lane + 32 * ps;
    const int r = r0 + rl;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rl < tile_rows && r < rows) {
      const float* src = op.p + (int64_t)k * op.ld + r;
      if (k < kend) v.x = __ldg(src);
      if (k + 1 < kend) v.y = __ldg(src + op.ld);
      if (k + 2 < kend) v.z = __ldg(src + 2 * op.ld);
      if (k + 3 < kend) v.w = __ldg(src + 3 * op.ld);
    }
    reg[ps] = v;
  }
}
// registers -> split -> swizzled shared memory (row r, 16-byte chunk c -> r * 128 + ((c ^ (r & 7)) << 4))
template <int PASSES>
__device__ __forceinline__ void store_split(uint8_t* hi, uint8_t* lo, int k_contig, int tile_rows,
                                            const float4 (&reg)[PASSES]) {
#pragma unroll
  for (int ps = 0; ps < PASSES; ++ps) {
    int rl, c;
    if (k_contig) {
      rl = (threadIdx.x >> 3) + 32 * ps;
      c = threadIdx.x & 7;
    } else {
      rl = (threadIdx.x & 31) + 32 * ps;
      c = threadIdx.x >> 5;
    }
    if (rl < tile_rows) {
      const float4 v = reg[ps];
      uint4 h, l;
      h.x = rna_tf32(v.x); h.y = rna_tf32(v.y); h.z = rna_tf32(v.z); h.w = rna_tf32(v.w);
      l.x = rna_tf32(v.x - __uint_as_float(h.x));
      l.y = rna_tf32(v.y - __uint_as_float(h.y));
      l.z = rna_tf32(v.z - __uint_as_float(h.z));
      l.w = rna_tf32(v.w - __uint_as_float(h.w));
      const int off = rl * 128 + ((c ^ (rl & 7)) << 4);
      *reinterpret_cast<uint4*>(hi + off) = h;
      *reinterpret_cast<uint4*>(lo + off) = l;
    }
  }
}

// Weight packer: writes, for every (N tile nt, K block kb), the exact shared-memory image
// [B_hi (BN x 128 B, SWIZZLE_128B) | B_lo] at  out + (nt * nkb + kb) * 2 * BN * 128.
__global__ void __launch_bounds__(256)
tc_pack_b_kernel(Operand b0, int K0, Operand b1, int K1, int N, int BN, uint8_t* __restrict__ out) {
  const int nkb0 = (K0 + BK - 1) / BK, nkb1 = (K1 + BK - 1) / BK;
  const int nkb = nkb0 + nkb1;
  const int ntiles = (N + BN - 1) / BN;
  const int64_t total = (int64_t)ntiles * nkb * BN * 8;
  for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256) {
    const int c = (int)(idx & 7);
    int64_t t = idx >> 3;
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
    uint8_t* tile = out + ((int64_t)nt * nkb + kb) * (2 * (int64_t)BN * 128);
    const int off = r * 128 + ((c ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(tile + off) = h;
    *reinterpret_cast<uint4*>(tile + (int64_t)BN * 128 + off) = l;
  }
}

// C[m][n] = sum over segments, k of A(m,k) * B(k,n)  (+ bias[n]);  blockIdx.z = K split of segment 0.
// B_PACKED: the B operand comes pre-split / pre-swizzled from tc_pack_b_kernel (`packed_b`).
template <int BN, bool B_PACKED>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(Segment s0, Segment s1, const uint8_t* __restrict__ packed_b, int M, int N, float* __restrict__ c,
               int64_t ldc, int c_vec, const float* __restrict__ bias, int k_per_split, int64_t split_stride) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  // bars[0..S) full, bars[S..2S) empty, bars[2S] accumulator ready, then the TMEM base address
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 1);

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
      mbar_init(smem_u32(bars + s), kProducerThreads + (B_PACKED ? 1 : 0));
      mbar_init(smem_u32(bars + C::kStages + s), 1);
    }
    mbar_init(smem_u32(bars + 2 * C::kStages), 1);
    fence_barrier_init();
  }
  if (warp == kProducerWarps) tmem_alloc(smem_u32(tmem_slot), C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kProducerWarps) {
    // ================= producers =================
    float4 ra[2][C::kPassesA];
    float4 rb[2][B_PACKED ? 1 : C::kPassesB];
    auto fetch = [&](int it, float4 (&fa)[C::kPassesA], float4 (&fb)[B_PACKED ? 1 : C::kPassesB]) {
      if (it >= nb) return;
      const bool first = it < nb0;
      const Segment& s = first ? s0 : s1;
      const int k0 = first ? k_lo + it * BK : (it - nb0) * BK;
      const int kend = first ? k_hi : s1.K;
      if (s.a.k_contig) load_kc<C::kPassesA>(s.a, m0, M, BM, k0, kend, fa);
      else load_mn<C::kPassesA>(s.a, m0, M, BM, k0, kend, fa);
      if constexpr (!B_PACKED) {
        if (s.b.k_contig) load_kc<C::kPassesB>(s.b, n0, N, BN, k0, kend, fb);
        else load_mn<C::kPassesB>(s.b, n0, N, BN, k0, kend, fb);
      }
    };
    auto stash = [&](int it, float4 (&fa)[C::kPassesA], float4 (&fb)[B_PACKED ? 1 : C::kPassesB]) {
      const int s = it % C::kStages;
      const uint32_t ph = (uint32_t)(it / C::kStages) & 1u;
      const Segment& seg = it < nb0 ? s0 : s1;
      mbar_wait(smem_u32(bars + C::kStages + s), ph ^ 1u);          // slot free (MMAs that read it retired)
      uint8_t* st = smem + s * C::kStageBytes;
      store_split<C::kPassesA>(st, st + C::kABytes, seg.a.k_contig, BM, fa);
      if constexpr (!B_PACKED)
        store_split<C::kPassesB>(st + 2 * C::kABytes, st + 2 * C::kABytes + C::kBBytes, seg.b.k_contig, BN, fb);
      fetch(it + 2, fa, fb);                                        // refill this register buffer: 2 blocks ahead
      fence_proxy_async();                                          // generic-proxy writes -> async proxy (UMMA)
      mbar_arrive(smem_u32(bars + s));
    };
    fetch(0, ra[0], rb[0]);
    fetch(1, ra[1], rb[1]);
    for (int it = 0; it < nb; it += 2) {
      stash(it, ra[0], rb[0]);
      if (it + 1 < nb) stash(it + 1, ra[1], rb[1]);
    }
    // ================= epilogue =================
    mbar_wait(smem_u32(bars + 2 * C::kStages), 0);
    tc_fence_after();
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = warp >> 2;                // column half
    constexpr int kHalf = BN / 2;
    static_assert(kHalf % 8 == 0, "BN / 2 must be a multiple of 8");
    const int m = m0 + q * 32 + lane;
    for (int cc = 0; cc < kHalf; cc += 8) {
      const int nl = half * kHalf + cc;
      float v[8];
      float vc[8];
      tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)nl, v);
      tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(C::kCorrCol + nl), vc);
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] += vc[u];
      const int n = n0 + nl;
      if (nb == 0) {
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = 0.f;
      }
      if (m < M && n < N) {
        if (bias != nullptr) {
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (n + u < N) v[u] += __ldg(bias + n + u);
        }
        float* dst = c + (int64_t)m * ldc + n;
        if (c_vec == 4 && n + 8 <= N) {
          *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
        } else if (c_vec >= 2 && n + 8 <= N) {
#pragma unroll
          for (int u = 0; u < 8; u += 2) *reinterpret_cast<float2*>(dst + u) = make_float2(v[u], v[u + 1]);
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (n + u < N) dst[u] = v[u];
        }
      }
    }
    tc_fence_before();
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
          const uint64_t adv = (uint64_t)(k * 32 >> 4);             // 8 tf32 = 32 bytes along the swizzled row
          umma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, (it | k) != 0);
          umma_tf32(tmem_base + C::kCorrCol, a_lo + adv, b_hi + adv, idesc, (it | k) != 0);
          umma_tf32(tmem_base + C::kCorrCol, a_hi + adv, b_lo + adv, idesc, 1);
        }
        umma_commit(smem_u32(bars + C::kStages + s));               // frees the stage when these MMAs retire
      }
      umma_commit(smem_u32(bars + 2 * C::kStages));                 // accumulator complete
    }
  } else if (B_PACKED && lane == 0) {
    // ================= weight loader (one thread, TMA engine) =================
    const uint8_t* src = packed_b + (int64_t)blockIdx.x * nb * (2 * C::kBBytes);
    for (int it = 0; it < nb; ++it) {
      const int s = it % C::kStages;
      const uint32_t ph = (uint32_t)(it / C::kStages) & 1u;
      mbar_wait(smem_u32(bars + C::kStages + s), ph ^ 1u);
      const uint32_t full = smem_u32(bars + s);
      mbar_arrive_expect_tx(full, 2 * C::kBBytes);
      bulk_g2s(smem_u32(smem + s * C::kStageBytes + 2 * C::kABytes), src + (int64_t)it * (2 * C::kBBytes),
               2 * C::kBBytes, full);
    }
  }
  __syncthreads();
  if (warp == kProducerWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

}  // namespace tc
}  // namespace mgs
