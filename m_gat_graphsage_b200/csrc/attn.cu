// K5 -- streaming all-pairs attention for the reference's ModifiedGATLayer (train.py:87-99, SURVEY.md
// section 8 rows a11 / f-1):
//
//     out[b] = sum_i softmax_i( <qry[b], key[i]> * scale ) val[i]              (qry = K_new, key = Q, val = V)
//
// over every atom of the batch ("global": what the reference computes with a dense [N, N] score matrix, 68 GB
// at N = 130 k) or only over the atoms of b's own molecule ("segmented": what test.py / gnnexplainer.py get
// one molecule at a time).  Flash-style: the score matrix never exists; a CTA owns BM query rows, streams
// 64-row tiles of key / val through shared memory and keeps a running (max, sum, accumulator) per row.
// fp32 throughout on the CUDA cores (d = 35: the tensor pipe would need a 3xTF32 split of both products;
// the CUDA-core form is exact fp32 and bounded by FFMA issue, 83 FFMA per score).
//
// One tile engine, three modes:
//   FWD     rows = queries  cols = keys      acc  += P  val[cols]                      -> out, lse
//   BWD_Q   rows = queries  cols = keys      acc  += dS key[cols]                      -> dqry
//   BWD_KV  rows = keys     cols = queries   accV += P^T gout[cols], accK += dS^T qry[cols]  -> dval, dkey
// with P = exp2(s*scale*log2e - lse2[query]), dP = <gout[query], val[key]>, dS = P (dP - delta[query]) scale.
// No atomics: every output element is owned by one thread, results are run-to-run reproducible.
//
// Thread layout: 256 threads = 16 (ty) x 16 (tx); a thread owns RM rows (ty*RM..) x 4 cols (tx*4..) of the
// score tile and RM rows x {tx, tx+16, tx+32, ..} of the accumulator.  The 16 threads that share a row are one
// half-warp, so row reductions are shuffles and the P tile goes through shared memory under __syncwarp only.
#include "common.cuh"

#include <cmath>

namespace mgs {
namespace {

constexpr int kT = 256;
constexpr int BN = 64;          // columns per streamed tile
constexpr int CS = BN + 4;      // d-major column tile stride (16 B aligned, 4-way conflicts on the transposing store)
constexpr int WS = BN + 4;      // P / dS tile row stride
enum { A_FWD = 0, A_BWD_Q = 1, A_BWD_KV = 2 };

struct AttnArgs {
  const float* X; int64_t ldx;      // row operand of the scores (FWD/BWD_Q: qry, BWD_KV: key)
  const float* Y; int64_t ldy;      // column operand           (FWD/BWD_Q: key, BWD_KV: qry)
  const float* V; int64_t ldv;      // val
  const float* G; int64_t ldg;      // gout (backward)
  const float* lse2;                // per query: log2 of the softmax denominator (incl. the running max)
  const float* delta;               // per query: <gout, attention output>
  const int* seg; const int* gptr;  // molecule of every atom / atom range of every molecule; null = global
  float* O1; int64_t ldo1;          // FWD: out, BWD_Q: dqry, BWD_KV: dval
  float* O2; int64_t ldo2;          // BWD_KV: dkey
  float* lse_out;                   // FWD
  int N; int d;
  float scale, c2;                  // c2 = scale * log2(e)
};

// Accumulator geometry for head width D: NC "main" column groups of 16 (lane tx owns columns tx + 16u) and, when
// D mod 16 <= 4, EX "extra" columns kept next to them in the row-major tile (D = 35: 32 + 3 instead of 48 with 13
// idle columns, 27 % of the accumulate FFMAs).  The extra columns of row a are summed by the lanes tx = a (mod RM),
// each over its own 4*RM of the 64 tile columns, and reduced across those lanes once at the end.
template <int D> struct Geo {
  static constexpr int EX = (D % 16 != 0 && D % 16 <= 4) ? D % 16 : 0;
  static constexpr int NC = EX ? D / 16 : (D + 15) / 16;
  static_assert(NC % 2 == 0, "main columns come in groups of 32");
  static constexpr int NU = NC / 2;                   // 4-column groups per lane
  static constexpr int ZS = NC * 16;                  // row-major tile stride (main columns)
  // the extra columns of tile row j live behind the main block at ((j + j/16) * 4 + c): the +j/16 skew puts the
  // rows jb = 0, 16, 32, 48 that the lanes of a half-warp read together into different banks (inside the main
  // block, 16 rows apart is always the same bank: 4-way conflicts, seen as +25 % shared wavefronts in BWD_KV)
  static constexpr int ZE = EX ? (BN + BN / 16) * 4 : 0;
  static constexpr int ZT = BN * ZS + ZE;             // floats per row-major column tile
};

__device__ __forceinline__ float ex2(float x) {         // 2^x, flush-to-zero, 2 ulp: one MUFU
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// rows [r0, r0+ROWS) x [0, D) of a row-major matrix -> registers (fetch) -> shared memory (commit), d-major
// (dm[k*DS + r]) and / or row-major (rm[r*ZS + k]); rows >= N and columns >= d read as zero.  Split in two so that
// the global loads of the next tile are in flight while the current one is computed.
// Mapping: warp w takes rows w, w+8, ..; lane l takes columns l, l+32, .. (a row is one coalesced request); the
// D mod 32 remainder columns of all the warp's rows are packed over the lanes.  Every shared / global address is a
// per-thread base plus a compile-time offset (an element-linear mapping cost ~15 integer instructions per element,
// a quarter of the kernel's issue slots).
template <int D, int ROWS> struct TileRegs {
  static constexpr int RPW = ROWS / 8;                    // rows per warp
  static constexpr int KF = D / 32, KR = D % 32;          // full 32-column passes, remainder columns
  static constexpr int NR = (RPW * KR + 31) / 32;         // registers for the packed remainder
  float v[RPW * KF + (NR ? NR : 1)];
};

template <int D, int ROWS>
__device__ __forceinline__ void fetch_tile(const float* __restrict__ src, int64_t ld, int r0, int N, int d,
                                           TileRegs<D, ROWS>& t) {
  using T = TileRegs<D, ROWS>;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* base = src + (int64_t)(r0 + w) * ld + lane;
  const bool full = r0 + ROWS <= N;
#pragma unroll
  for (int i = 0; i < T::RPW; ++i)
#pragma unroll
    for (int f = 0; f < T::KF; ++f) {
      const bool ok = (full || r0 + w + 8 * i < N) && lane + 32 * f < d;
      t.v[i * T::KF + f] = ok ? __ldg(base + (int64_t)(8 * i) * ld + 32 * f) : 0.f;
    }
#pragma unroll
  for (int q = 0; q < T::NR; ++q) {
    const int e = lane + 32 * q, i = e / (T::KR ? T::KR : 1), k = 32 * T::KF + e - i * T::KR;
    const bool ok = e < T::RPW * T::KR && (full || r0 + w + 8 * i < N) && k < d;
    t.v[T::RPW * T::KF + q] = ok ? __ldg(src + (int64_t)(r0 + w + 8 * i) * ld + k) : 0.f;
  }
}

template <int D, int ROWS, int DS, int ZS>
__device__ __forceinline__ void commit_tile(const TileRegs<D, ROWS>& t, float* __restrict__ dm, float* __restrict__ rm) {
  using T = TileRegs<D, ROWS>;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* dmb = dm ? dm + lane * DS + w : nullptr;
  float* rmb = rm ? rm + w * ZS + lane : nullptr;
#pragma unroll
  for (int i = 0; i < T::RPW; ++i)
#pragma unroll
    for (int f = 0; f < T::KF; ++f) {
      if (dm) dmb[32 * f * DS + 8 * i] = t.v[i * T::KF + f];
      if (rm) rmb[8 * i * ZS + 32 * f] = t.v[i * T::KF + f];
    }
#pragma unroll
  for (int q = 0; q < T::NR; ++q) {
    const int e = lane + 32 * q, i = e / (T::KR ? T::KR : 1), k = 32 * T::KF + e - i * T::KR;
    if (e < T::RPW * T::KR) {
      if (dm) dm[k * DS + w + 8 * i] = t.v[T::RPW * T::KF + q];
      if (rm) {
        if constexpr (Geo<D>::EX > 0) {
          const int row = w + 8 * i;
          rm[ROWS * ZS + (row + (row >> 4)) * 4 + (k - 32 * T::KF)] = t.v[T::RPW * T::KF + q];
        } else {
          rm[(w + 8 * i) * ZS + k] = t.v[T::RPW * T::KF + q];
        }
      }
    }
  }
}

// zero the padding columns [D, ZS) of a row-major tile (once: commits never touch them)
template <int D, int ROWS, int ZS>
__device__ __forceinline__ void zero_pad(float* __restrict__ rm) {
  if constexpr (ZS > D) {
    for (int e = threadIdx.x; e < ROWS * (ZS - D); e += kT) {
      const int r = e / (ZS - D), k = D + (e - r * (ZS - D));
      rm[r * ZS + k] = 0.f;
    }
  }
}

template <int D, int ROWS, int DS, int ZS>
__device__ __forceinline__ void load_tile(const float* __restrict__ src, int64_t ld, int r0, int N, int d,
                                          float* __restrict__ dm, float* __restrict__ rm) {
  TileRegs<D, ROWS> t;
  fetch_tile<D, ROWS>(src, ld, r0, N, d, t);
  commit_tile<D, ROWS, DS, ZS>(t, dm, rm);
}

// t[a][b] = sum_k xs[k][ty*RM + a] * ys[k][tx*4 + b]
template <int D, int RM, int RS>
__device__ __forceinline__ void tile_dot(const float* __restrict__ xs, const float* __restrict__ ys, int ty, int tx,
                                         float (&t)[RM][4]) {
#pragma unroll
  for (int a = 0; a < RM; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) t[a][b] = 0.f;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    float xr[RM];
    if constexpr (RM % 4 == 0) {
#pragma unroll
      for (int a = 0; a < RM; a += 4) {
        const float4 q = *reinterpret_cast<const float4*>(xs + k * RS + ty * RM + a);
        xr[a] = q.x; xr[a + 1] = q.y; xr[a + 2] = q.z; xr[a + 3] = q.w;
      }
    } else {
#pragma unroll
      for (int a = 0; a < RM; a += 2) {
        const float2 q = *reinterpret_cast<const float2*>(xs + k * RS + ty * RM + a);
        xr[a] = q.x; xr[a + 1] = q.y;
      }
    }
    const float4 y = *reinterpret_cast<const float4*>(ys + k * CS + tx * 4);
#pragma unroll
    for (int a = 0; a < RM; ++a) {
      t[a][0] = fmaf(xr[a], y.x, t[a][0]);
      t[a][1] = fmaf(xr[a], y.y, t[a][1]);
      t[a][2] = fmaf(xr[a], y.z, t[a][2]);
      t[a][3] = fmaf(xr[a], y.w, t[a][3]);
    }
  }
}

// Accumulate mapping: the half-warp that owns score rows ty*RM .. ty*RM+RM-1 splits them by parity -- lanes
// tx < 8 take rows ty*RM + 0, 2, 4, .., lanes tx >= 8 rows ty*RM + 1, 3, 5, .. (RM/2 rows each) -- and lane
// (tx & 7) owns the 4 adjacent columns 4 (tx & 7) + 32 u: every operand comes in with LDS.128 (8 FFMA per shared
// memory wavefront; the 8-rows x {tx, tx+16} mapping before it had 4, which tied the LSU pipe with the FMA pipe).
// Rows of a P tile are stored with their column index XOR 16 for odd ty, so that the four rows a warp reads at
// once (2 ty x 2 parities) fall into four different bank groups.
//   acc[a][4u + q] += sum_j ws[ty*RM + h + 2a][j] * zs[j][4 (tx&7) + 32u + q]
// The 64 products of a tile are summed on their own and then added to the long-running accumulator: a two-level
// sum whose rounding error grows with sqrt(64) + sqrt(N / 64) instead of sqrt(N) (N = 130 k terms at B = 4096).
template <int RM, int NU, int ZS>
__device__ __forceinline__ void tile_accumulate(const float* __restrict__ ws, const float* __restrict__ zs, int ty,
                                                int tx, float (&acc)[RM / 2][NU * 4]) {
  constexpr int RP = RM / 2;
  const int h = tx >> 3, cq = (tx & 7) * 4, sw = (ty & 1) << 4;
  float t[RP][NU * 4];
#pragma unroll
  for (int a = 0; a < RP; ++a)
#pragma unroll
    for (int u = 0; u < NU * 4; ++u) t[a][u] = 0.f;
#pragma unroll 4
  for (int j0 = 0; j0 < BN; j0 += 4) {
    float4 w[RP];
#pragma unroll
    for (int a = 0; a < RP; ++a)
      w[a] = *reinterpret_cast<const float4*>(ws + (ty * RM + h + 2 * a) * WS + (j0 ^ sw));
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const float4 z = *reinterpret_cast<const float4*>(zs + (j0 + jj) * ZS + cq + 32 * u);
#pragma unroll
        for (int a = 0; a < RP; ++a) {
          const float wv = jj == 0 ? w[a].x : jj == 1 ? w[a].y : jj == 2 ? w[a].z : w[a].w;
          t[a][4 * u + 0] = fmaf(wv, z.x, t[a][4 * u + 0]);
          t[a][4 * u + 1] = fmaf(wv, z.y, t[a][4 * u + 1]);
          t[a][4 * u + 2] = fmaf(wv, z.z, t[a][4 * u + 2]);
          t[a][4 * u + 3] = fmaf(wv, z.w, t[a][4 * u + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < RP; ++a)
#pragma unroll
    for (int u = 0; u < NU * 4; ++u) acc[a][u] += t[a][u];
}

// extra columns: acce[c] += sum over this lane's 4*RM tile columns j of ws[ty*RM + (tx mod RM)][j] * zs[j][16 NC + c]
template <int RM, int NC, int EX, int ZS>
__device__ __forceinline__ void tile_accumulate_extra(const float* __restrict__ ws, const float* __restrict__ zs,
                                                      int ty, int tx, float (&acce)[EX ? EX : 1]) {
  if constexpr (EX > 0) {
    const int a = tx % RM, jb = (tx / RM) * (4 * RM), sw = (ty & 1) << 4;
    float t[EX];
#pragma unroll
    for (int c = 0; c < EX; ++c) t[c] = 0.f;
#pragma unroll
    for (int j0 = 0; j0 < 4 * RM; j0 += 4) {
      const float4 w = *reinterpret_cast<const float4*>(ws + (ty * RM + a) * WS + ((jb + j0) ^ sw));
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = jb + j0 + jj;
        const float4 z = *reinterpret_cast<const float4*>(zs + BN * ZS + (j + (j >> 4)) * 4);
        const float wv = jj == 0 ? w.x : jj == 1 ? w.y : jj == 2 ? w.z : w.w;
        t[0] = fmaf(wv, z.x, t[0]);
        if constexpr (EX > 1) t[1] = fmaf(wv, z.y, t[1]);
        if constexpr (EX > 2) t[2] = fmaf(wv, z.z, t[2]);
        if constexpr (EX > 3) t[3] = fmaf(wv, z.w, t[3]);
      }
    }
#pragma unroll
    for (int c = 0; c < EX; ++c) acce[c] += t[c];
  }
}

// sum of the extra-column partials over the lanes that share a row (tx = a mod RM); valid in every such lane
template <int RM>
__device__ __forceinline__ float extra_row_sum(float v) {
#pragma unroll
  for (int o = 8; o >= RM; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int RM>
__device__ __forceinline__ void store_w(float* __restrict__ ws, int ty, int tx, const float (&t)[RM][4]) {
#pragma unroll
  for (int a = 0; a < RM; ++a)
    *reinterpret_cast<float4*>(ws + (ty * RM + a) * WS + ((tx * 4) ^ ((ty & 1) << 4))) =
        make_float4(t[a][0], t[a][1], t[a][2], t[a][3]);
}

__device__ __forceinline__ float half_warp_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int MODE, int D, int RM>
constexpr size_t attn_smem_floats() {
  constexpr int BM = 16 * RM, RS = BM + 4;
  size_t n = (size_t)D * RS + (size_t)D * CS + (size_t)Geo<D>::ZT + (size_t)BM * WS;
  if (MODE != A_FWD) n += (size_t)D * RS + (size_t)D * CS;
  if (MODE == A_BWD_KV) n += (size_t)Geo<D>::ZT + (size_t)BM * WS + 2 * BN;
  return n + 2 * BM;                                // per-row valid column range
}

template <int MODE, int D, int RM, bool PF>
__global__ void __launch_bounds__(kT, 2) attn_kernel(const AttnArgs p) {
  constexpr int BM = 16 * RM, RS = BM + 4, NC = Geo<D>::NC, ZS = Geo<D>::ZS, EX = Geo<D>::EX, EXN = EX ? EX : 1,
                NU = Geo<D>::NU, RP = RM / 2, AC = NU * 4;
  constexpr bool BWD = MODE != A_FWD, KV = MODE == A_BWD_KV;
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;                               // [D][RS]   row operand, d-major
  float* X2s = Xs + D * RS;                       // [D][RS]   BWD_Q: gout rows, BWD_KV: val rows
  float* Ys = X2s + (BWD ? D * RS : 0);           // [D][CS]   column operand, d-major
  float* Y2s = Ys + D * CS;                       // [D][CS]   BWD_Q: val cols, BWD_KV: gout cols
  float* Zs = Y2s + (BWD ? D * CS : 0);           // [BN][ZS]  FWD: val, BWD_Q: key, BWD_KV: gout (row-major)
  float* Z2s = Zs + Geo<D>::ZT;                   // [BN][ZS]  BWD_KV: qry (row-major)
  float* Ws = Z2s + (KV ? Geo<D>::ZT : 0);           // [BM][WS]  P (FWD, BWD_KV) or dS (BWD_Q)
  float* W2s = Ws + BM * WS;                      // [BM][WS]  BWD_KV: dS
  float* stat = W2s + (KV ? BM * WS : 0);         // [2][BN]   BWD_KV: lse2, delta of the column queries

  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int N = p.N, d = p.d;
  const int r0 = blockIdx.x * BM;
  const bool segmented = p.seg != nullptr;

  // valid column range of every row (shared memory: only boundary tiles read it), of the whole CTA (cbeg, cend)
  // and the range valid for all rows (all_lo, all_hi: tiles inside it skip the mask)
  int* rlo = reinterpret_cast<int*>(stat + (KV ? 2 * BN : 0));
  int* rhi = rlo + BM;
  if (tid < BM) {
    const int r = r0 + tid;
    int lo_ = 0, hi_ = 0;
    if (r < N) {
      if (segmented) { const int g = __ldg(p.seg + r); lo_ = __ldg(p.gptr + g); hi_ = __ldg(p.gptr + g + 1); }
      else hi_ = N;
    }
    rlo[tid] = lo_; rhi[tid] = hi_;
  }
  int cbeg = 0, cend = N, all_lo = 0, all_hi = N;
  const int rl = min(r0 + BM, N) - 1;
  if (segmented) {
    const int g0 = __ldg(p.seg + r0), g1 = __ldg(p.seg + rl);
    cbeg = __ldg(p.gptr + g0); cend = __ldg(p.gptr + g1 + 1);
    all_lo = __ldg(p.gptr + g1); all_hi = __ldg(p.gptr + g0 + 1);
  }
  if (r0 + BM > N) all_hi = all_lo;               // the tile has rows past the end: always take the masked path

  float lse_r[RM], del_r[RM];
  if constexpr (MODE == A_BWD_Q) {
#pragma unroll
    for (int a = 0; a < RM; ++a) {
      const int r = r0 + ty * RM + a;
      lse_r[a] = r < N ? __ldg(p.lse2 + r) : 0.f;
      del_r[a] = r < N ? __ldg(p.delta + r) : 0.f;
    }
  }

  load_tile<D, BM, RS, ZS>(p.X, p.ldx, r0, N, d, Xs, nullptr);
  if constexpr (MODE == A_BWD_Q) load_tile<D, BM, RS, ZS>(p.G, p.ldg, r0, N, d, X2s, nullptr);
  if constexpr (KV) load_tile<D, BM, RS, ZS>(p.V, p.ldv, r0, N, d, X2s, nullptr);

  float acc[RP][AC], acc2[KV ? RP : 1][KV ? AC : 1];
  float acce[EXN], acce2[EXN];
#pragma unroll
  for (int c = 0; c < EXN; ++c) acce[c] = acce2[c] = 0.f;
  float m[RM], l[RM];
#pragma unroll
  for (int a = 0; a < RM; ++a) { m[a] = -INFINITY; l[a] = 0.f; }
#pragma unroll
  for (int a = 0; a < RP; ++a)
#pragma unroll
    for (int u = 0; u < AC; ++u) { acc[a][u] = 0.f; if constexpr (KV) acc2[a][u] = 0.f; }
  const bool odd = (tx >> 3) != 0;                    // this lane accumulates the odd rows of its ty group

  // column tiles: A = the score operand (key / qry), B = val (FWD, BWD_Q) or gout (BWD_KV)
  TileRegs<D, BN> ta, tb;
  float st_l = 0.f, st_d = 0.f;
  const float* srcB = KV ? p.G : p.V;
  const int64_t ldB = KV ? p.ldg : p.ldv;
  auto fetch = [&](int c0) {
    fetch_tile<D, BN>(p.Y, p.ldy, c0, N, d, ta);
    fetch_tile<D, BN>(srcB, ldB, c0, N, d, tb);
    if constexpr (KV) {
      const int c = c0 + tid;
      const bool ok = tid < BN && c < N;
      st_l = ok ? __ldg(p.lse2 + c) : 0.f;
      st_d = ok ? __ldg(p.delta + c) : 0.f;
    }
  };
  zero_pad<D, BN, ZS>(Zs);
  if constexpr (KV) zero_pad<D, BN, ZS>(Z2s);
  for (int e = tid; e < Geo<D>::ZE; e += kT) {       // 4th lane of the extra-column quads stays zero
    Zs[BN * ZS + e] = 0.f;
    if constexpr (KV) Z2s[BN * ZS + e] = 0.f;
  }
  if (PF && cbeg < cend) fetch(cbeg);

  for (int c0 = cbeg; c0 < cend; c0 += BN) {
    __syncthreads();
    if constexpr (!PF) fetch(c0);
    if constexpr (MODE == A_FWD) {
      commit_tile<D, BN, CS, ZS>(ta, Ys, nullptr);
      commit_tile<D, BN, CS, ZS>(tb, nullptr, Zs);
    } else if constexpr (MODE == A_BWD_Q) {
      commit_tile<D, BN, CS, ZS>(ta, Ys, Zs);
      commit_tile<D, BN, CS, ZS>(tb, Y2s, nullptr);
    } else {
      commit_tile<D, BN, CS, ZS>(ta, Ys, Z2s);
      commit_tile<D, BN, CS, ZS>(tb, Y2s, Zs);
      if (tid < BN) { stat[tid] = st_l; stat[BN + tid] = st_d; }
    }
    __syncthreads();
    if (PF && c0 + BN < cend) fetch(c0 + BN);     // in flight during the compute below

    float s[RM][4];
    tile_dot<D, RM, RS>(Xs, Ys, ty, tx, s);
    const bool interior = c0 >= all_lo && c0 + BN <= all_hi;
    const int cb = c0 + tx * 4;

    if constexpr (MODE == A_FWD) {
      float corr_e = 1.f;                           // rescale factor of the row whose extra columns this lane sums
      float corr_a[RP];                             // ... of the rows whose main columns it sums
#pragma unroll
      for (int a = 0; a < RM; ++a) {
        if (!interior) {
          const int lo = rlo[ty * RM + a], hi = rhi[ty * RM + a];
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (cb + b < lo || cb + b >= hi) s[a][b] = -INFINITY;
        }
        // running maximum in raw score units (c2 > 0 keeps the order); exponent = fma(s, c2, -m c2)
        const float mt = half_warp_max(fmaxf(fmaxf(s[a][0], s[a][1]), fmaxf(s[a][2], s[a][3])));
        const float mn = fmaxf(m[a], mt);
        const float ms = mn == -INFINITY ? 0.f : mn * p.c2;
        const float corr = ex2(fmaf(m[a], p.c2, -ms));
        m[a] = mn;
        float ps = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) { s[a][b] = ex2(fmaf(s[a][b], p.c2, -ms)); ps += s[a][b]; }
        l[a] = fmaf(l[a], corr, ps);
        if ((a & 1) == 0) corr_a[a >> 1] = corr;
        else if (odd) corr_a[a >> 1] = corr;
        if (EX && a == tx % RM) corr_e = corr;
      }
#pragma unroll
      for (int a = 0; a < RP; ++a)
#pragma unroll
        for (int u = 0; u < AC; ++u) acc[a][u] *= corr_a[a];
#pragma unroll
      for (int c = 0; c < EXN; ++c) acce[c] *= corr_e;
      store_w<RM>(Ws, ty, tx, s);
      __syncwarp();
      tile_accumulate<RM, NU, ZS>(Ws, Zs, ty, tx, acc);
      tile_accumulate_extra<RM, NC, EX, ZS>(Ws, Zs, ty, tx, acce);
    } else {
      float dp[RM][4];
      tile_dot<D, RM, RS>(X2s, Y2s, ty, tx, dp);
      float lc[4], dc[4];
      if constexpr (KV) {
        const float4 t0 = *reinterpret_cast<const float4*>(stat + tx * 4);
        const float4 t1 = *reinterpret_cast<const float4*>(stat + BN + tx * 4);
        lc[0] = t0.x; lc[1] = t0.y; lc[2] = t0.z; lc[3] = t0.w;
        dc[0] = t1.x; dc[1] = t1.y; dc[2] = t1.z; dc[3] = t1.w;
      }
#pragma unroll
      for (int a = 0; a < RM; ++a) {
        int lo = 0, hi = 0;
        if (!interior) { lo = rlo[ty * RM + a]; hi = rhi[ty * RM + a]; }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          float lse, del;
          if constexpr (KV) { lse = lc[b]; del = dc[b]; } else { lse = lse_r[a]; del = del_r[a]; }
          float pv = ex2(fmaf(s[a][b], p.c2, -lse));
          if (!interior && (cb + b < lo || cb + b >= hi)) pv = 0.f;
          s[a][b] = pv;
          dp[a][b] = pv * (dp[a][b] - del) * p.scale;
        }
      }
      if constexpr (KV) {
        store_w<RM>(Ws, ty, tx, s);
        store_w<RM>(W2s, ty, tx, dp);
        __syncwarp();
        tile_accumulate<RM, NU, ZS>(Ws, Zs, ty, tx, acc);
        tile_accumulate_extra<RM, NC, EX, ZS>(Ws, Zs, ty, tx, acce);
        tile_accumulate<RM, NU, ZS>(W2s, Z2s, ty, tx, acc2);
        tile_accumulate_extra<RM, NC, EX, ZS>(W2s, Z2s, ty, tx, acce2);
      } else {
        store_w<RM>(Ws, ty, tx, dp);
        __syncwarp();
        tile_accumulate<RM, NU, ZS>(Ws, Zs, ty, tx, acc);
        tile_accumulate_extra<RM, NC, EX, ZS>(Ws, Zs, ty, tx, acce);
      }
    }
  }

  float lt_e = 1.f;                                 // softmax denominator of this lane's extra-column row
  float lt_a[RP];                                   // ... of the rows whose main columns it holds
#pragma unroll
  for (int a = 0; a < RP; ++a) lt_a[a] = 1.f;
  if constexpr (MODE == A_FWD) {
#pragma unroll
    for (int a = 0; a < RM; ++a) {
      const int r = r0 + ty * RM + a;
      const float lt = half_warp_sum(l[a]);
      if ((a & 1) == 0) lt_a[a >> 1] = lt;
      else if (odd) lt_a[a >> 1] = lt;
      if (EX && a == tx % RM) lt_e = lt;
      if (r < N && tx == 0) p.lse_out[r] = fmaf(m[a], p.c2, log2f(lt));
    }
  }
#pragma unroll
  for (int a = 0; a < RP; ++a) {
    const int r = r0 + ty * RM + (tx >> 3) + 2 * a;
    if (r >= N) continue;
#pragma unroll
    for (int u = 0; u < AC; ++u) {
      const int c = (tx & 7) * 4 + 32 * (u >> 2) + (u & 3);
      if (c < d) {
        p.O1[(int64_t)r * p.ldo1 + c] = MODE == A_FWD ? __fdiv_rn(acc[a][u], lt_a[a]) : acc[a][u];
        if constexpr (KV) p.O2[(int64_t)r * p.ldo2 + c] = acc2[a][u];
      }
    }
  }
  if constexpr (EX > 0) {
    const int r = r0 + ty * RM + tx;                // lanes tx < RM write the extra columns of row ty*RM + tx
#pragma unroll
    for (int c = 0; c < EX; ++c) {
      const float v = extra_row_sum<RM>(acce[c]);
      float v2 = 0.f;
      if constexpr (KV) v2 = extra_row_sum<RM>(acce2[c]);
      const int col = 16 * NC + c;
      if (tx < RM && r < N && col < d) {
        p.O1[(int64_t)r * p.ldo1 + col] = MODE == A_FWD ? __fdiv_rn(v, lt_e) : v;
        if constexpr (KV) p.O2[(int64_t)r * p.ldo2 + col] = v2;
      }
    }
  }
}

template <int MODE, int D, int RM, bool PF = true>
int launch(const AttnArgs& a, cudaStream_t stream) {
  constexpr int BM = 16 * RM;
  constexpr size_t bytes = attn_smem_floats<MODE, D, RM>() * sizeof(float);
  static_assert(bytes <= 227 * 1024, "shared memory per CTA");
  auto kern = attn_kernel<MODE, D, RM, PF>;
  MGS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  const int grid = (a.N + BM - 1) / BM;
  kern<<<grid, kT, bytes, stream>>>(a);
  return check_launch("attn_kernel");
}

template <int MODE, int D>
int dispatch_rm(const AttnArgs& a, cudaStream_t stream) {
  // 128-row tiles while they fill two CTAs per SM, then 64-row, then 32-row tiles (more CTAs for small batches)
  if constexpr (MODE == A_FWD) {
    // (no register prefetch at 128 rows: its 18 registers spill, measured 86 ms vs 76 ms at N = 130 k)
    if ((a.N + 127) / 128 >= 2 * sm_count()) return launch<MODE, D, 8, false>(a, stream);
  }
  if ((a.N + 63) / 64 >= sm_count()) return launch<MODE, D, 4>(a, stream);
  return launch<MODE, D, 2>(a, stream);
}

template <int MODE>
int dispatch(const AttnArgs& a, cudaStream_t stream) {
  if (a.d <= 35) return dispatch_rm<MODE, 35>(a, stream);
  return dispatch_rm<MODE, 64>(a, stream);
}

int check_common(const char* who, int64_t N, int32_t d, int64_t ldq, int64_t ldk, int64_t ldv, const int32_t* seg,
                 const int32_t* gptr) {
  MGS_REQUIRE(N >= 0 && N < 0x7fffffff - 256, "%s: bad atom count", who);
  MGS_REQUIRE(d > 0 && d <= 64, "%s: head width %d outside 1..64", who, d);
  MGS_REQUIRE(ldq >= d && ldk >= d && ldv >= d, "%s: leading dimension < d", who);
  MGS_REQUIRE((seg == nullptr) == (gptr == nullptr), "%s: seg and gptr go together", who);
  return MGS_OK;
}

}  // namespace
}  // namespace mgs

using namespace mgs;

extern "C" int mgs_attn_fwd(const float* qry, int64_t ldq, const float* key, int64_t ldk, const float* val,
                            int64_t ldv, int64_t N, int32_t d, float scale, const int32_t* seg,
                            const int32_t* gptr, float* out, int64_t ldo, float* lse2, mgs_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_common("mgs_attn_fwd", N, d, ldq, ldk, ldv, seg, gptr)) return rc;
  MGS_REQUIRE(ldo >= d, "mgs_attn_fwd: leading dimension < d");
  if (N == 0) return MGS_OK;
  MGS_REQUIRE(qry && key && val && out && lse2, "mgs_attn_fwd: null pointer");
  AttnArgs a{};
  a.X = qry; a.ldx = ldq; a.Y = key; a.ldy = ldk; a.V = val; a.ldv = ldv;
  a.seg = seg; a.gptr = gptr; a.O1 = out; a.ldo1 = ldo; a.lse_out = lse2;
  a.N = (int)N; a.d = d; a.scale = scale; a.c2 = scale * 1.4426950408889634f;
  return dispatch<A_FWD>(a, stream);
}

extern "C" int mgs_attn_bwd(const float* qry, int64_t ldq, const float* key, int64_t ldk, const float* val,
                            int64_t ldv, int64_t N, int32_t d, float scale, const int32_t* seg,
                            const int32_t* gptr, const float* lse2, const float* delta, const float* gout,
                            int64_t ldg, float* dqry, int64_t lddq, float* dkey, int64_t lddk, float* dval,
                            int64_t lddv, mgs_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_common("mgs_attn_bwd", N, d, ldq, ldk, ldv, seg, gptr)) return rc;
  MGS_REQUIRE(ldg >= d && lddq >= d && lddk >= d && lddv >= d, "mgs_attn_bwd: leading dimension < d");
  if (N == 0) return MGS_OK;
  MGS_REQUIRE(qry && key && val && lse2 && delta && gout && dqry && dkey && dval, "mgs_attn_bwd: null pointer");
  AttnArgs a{};
  a.V = val; a.ldv = ldv; a.G = gout; a.ldg = ldg; a.lse2 = lse2; a.delta = delta; a.seg = seg; a.gptr = gptr;
  a.N = (int)N; a.d = d; a.scale = scale; a.c2 = scale * 1.4426950408889634f;
  AttnArgs q = a;
  q.X = qry; q.ldx = ldq; q.Y = key; q.ldy = ldk; q.O1 = dqry; q.ldo1 = lddq;
  if (int rc = dispatch<A_BWD_Q>(q, stream)) return rc;
  AttnArgs k = a;
  k.X = key; k.ldx = ldk; k.Y = qry; k.ldy = ldq; k.O1 = dval; k.ldo1 = lddv; k.O2 = dkey; k.ldo2 = lddk;
  return dispatch<A_BWD_KV>(k, stream);
}
