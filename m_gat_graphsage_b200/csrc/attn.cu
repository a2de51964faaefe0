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

template <int D> struct Geo {
  static constexpr int NC = (D + 15) / 16;   // accumulator columns per thread
  static constexpr int ZS = NC * 16;         // row-major tile stride (zero padded)
};

// rows [r0, r0+ROWS) x [0, D) of a row-major matrix -> registers (fetch) -> shared memory (commit), d-major
// (dm[k*DS + r]) and / or row-major (rm[r*ZS + k], zero padded to ZS); rows >= N and columns >= d read as zero.
// Split in two so that the global loads of the next tile are in flight while the current one is computed.
template <int D, int ROWS> struct TileRegs {
  static constexpr int CNT = (ROWS * D + kT - 1) / kT;
  float v[CNT];
};

template <int D, int ROWS>
__device__ __forceinline__ void fetch_tile(const float* __restrict__ src, int64_t ld, int r0, int N, int d,
                                           TileRegs<D, ROWS>& t) {
#pragma unroll
  for (int i = 0; i < TileRegs<D, ROWS>::CNT; ++i) {
    const int e = threadIdx.x + i * kT;
    const int r = e / D, k = e - r * D;
    t.v[i] = (e < ROWS * D && r0 + r < N && k < d) ? __ldg(src + (int64_t)(r0 + r) * ld + k) : 0.f;
  }
}

template <int D, int ROWS, int DS, int ZS>
__device__ __forceinline__ void commit_tile(const TileRegs<D, ROWS>& t, float* __restrict__ dm, float* __restrict__ rm) {
#pragma unroll
  for (int i = 0; i < TileRegs<D, ROWS>::CNT; ++i) {
    const int e = threadIdx.x + i * kT;
    const int r = e / D, k = e - r * D;
    if (e < ROWS * D) {
      if (dm) dm[k * DS + r] = t.v[i];
      if (rm) rm[r * ZS + k] = t.v[i];
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

// acc[a][u] += sum_j ws[ty*RM + a][j] * zs[j][tx + 16u].  The 64 products of a tile are summed on their own and
// then added to the long-running accumulator: a two-level sum whose rounding error grows with
// sqrt(64) + sqrt(N / 64) instead of sqrt(N) (N = 130 k terms at B = 4096).
template <int RM, int NC, int ZS>
__device__ __forceinline__ void tile_accumulate(const float* __restrict__ ws, const float* __restrict__ zs, int ty,
                                                int tx, float (&acc)[RM][NC]) {
  float t[RM][NC];
#pragma unroll
  for (int a = 0; a < RM; ++a)
#pragma unroll
    for (int u = 0; u < NC; ++u) t[a][u] = 0.f;
#pragma unroll 2
  for (int j0 = 0; j0 < BN; j0 += 4) {
    float4 w[RM];
#pragma unroll
    for (int a = 0; a < RM; ++a) w[a] = *reinterpret_cast<const float4*>(ws + (ty * RM + a) * WS + j0);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      float z[NC];
#pragma unroll
      for (int u = 0; u < NC; ++u) z[u] = zs[(j0 + jj) * ZS + tx + 16 * u];
#pragma unroll
      for (int a = 0; a < RM; ++a) {
        const float wv = jj == 0 ? w[a].x : jj == 1 ? w[a].y : jj == 2 ? w[a].z : w[a].w;
#pragma unroll
        for (int u = 0; u < NC; ++u) t[a][u] = fmaf(wv, z[u], t[a][u]);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < RM; ++a)
#pragma unroll
    for (int u = 0; u < NC; ++u) acc[a][u] += t[a][u];
}

template <int RM>
__device__ __forceinline__ void store_w(float* __restrict__ ws, int ty, int tx, const float (&t)[RM][4]) {
#pragma unroll
  for (int a = 0; a < RM; ++a)
    *reinterpret_cast<float4*>(ws + (ty * RM + a) * WS + tx * 4) = make_float4(t[a][0], t[a][1], t[a][2], t[a][3]);
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
  constexpr int BM = 16 * RM, RS = BM + 4, ZS = Geo<D>::ZS;
  size_t n = (size_t)D * RS + (size_t)D * CS + (size_t)BN * ZS + (size_t)BM * WS;
  if (MODE != A_FWD) n += (size_t)D * RS + (size_t)D * CS;
  if (MODE == A_BWD_KV) n += (size_t)BN * ZS + (size_t)BM * WS + 2 * BN;
  return n;
}

template <int MODE, int D, int RM, bool PF>
__global__ void __launch_bounds__(kT, 2) attn_kernel(const AttnArgs p) {
  constexpr int BM = 16 * RM, RS = BM + 4, NC = Geo<D>::NC, ZS = Geo<D>::ZS;
  constexpr bool BWD = MODE != A_FWD, KV = MODE == A_BWD_KV;
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;                               // [D][RS]   row operand, d-major
  float* X2s = Xs + D * RS;                       // [D][RS]   BWD_Q: gout rows, BWD_KV: val rows
  float* Ys = X2s + (BWD ? D * RS : 0);           // [D][CS]   column operand, d-major
  float* Y2s = Ys + D * CS;                       // [D][CS]   BWD_Q: val cols, BWD_KV: gout cols
  float* Zs = Y2s + (BWD ? D * CS : 0);           // [BN][ZS]  FWD: val, BWD_Q: key, BWD_KV: gout (row-major)
  float* Z2s = Zs + BN * ZS;                      // [BN][ZS]  BWD_KV: qry (row-major)
  float* Ws = Z2s + (KV ? BN * ZS : 0);           // [BM][WS]  P (FWD, BWD_KV) or dS (BWD_Q)
  float* W2s = Ws + BM * WS;                      // [BM][WS]  BWD_KV: dS
  float* stat = W2s + (KV ? BM * WS : 0);         // [2][BN]   BWD_KV: lse2, delta of the column queries

  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int N = p.N, d = p.d;
  const int r0 = blockIdx.x * BM;
  const bool segmented = p.seg != nullptr;

  // valid column range of every owned row, of the whole CTA (cbeg, cend) and the range valid for all rows
  int lo[RM], hi[RM];
#pragma unroll
  for (int a = 0; a < RM; ++a) {
    const int r = r0 + ty * RM + a;
    lo[a] = 0; hi[a] = 0;
    if (r < N) {
      if (segmented) { const int g = __ldg(p.seg + r); lo[a] = __ldg(p.gptr + g); hi[a] = __ldg(p.gptr + g + 1); }
      else hi[a] = N;
    }
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

  float acc[RM][NC], acc2[KV ? RM : 1][KV ? NC : 1];
  float m[RM], l[RM];
#pragma unroll
  for (int a = 0; a < RM; ++a) {
    m[a] = -INFINITY; l[a] = 0.f;
#pragma unroll
    for (int u = 0; u < NC; ++u) { acc[a][u] = 0.f; if constexpr (KV) acc2[a][u] = 0.f; }
  }

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
#pragma unroll
      for (int a = 0; a < RM; ++a) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          s[a][b] *= p.c2;
          if (!interior && (cb + b < lo[a] || cb + b >= hi[a])) s[a][b] = -INFINITY;
        }
        const float mt = half_warp_max(fmaxf(fmaxf(s[a][0], s[a][1]), fmaxf(s[a][2], s[a][3])));
        const float mn = fmaxf(m[a], mt);
        const float ms = mn == -INFINITY ? 0.f : mn;
        const float corr = exp2f(m[a] - ms);
        m[a] = mn;
        float ps = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) { s[a][b] = exp2f(s[a][b] - ms); ps += s[a][b]; }
        l[a] = fmaf(l[a], corr, ps);
#pragma unroll
        for (int u = 0; u < NC; ++u) acc[a][u] *= corr;
      }
      store_w<RM>(Ws, ty, tx, s);
      __syncwarp();
      tile_accumulate<RM, NC, ZS>(Ws, Zs, ty, tx, acc);
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
      for (int a = 0; a < RM; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          float lse, del;
          if constexpr (KV) { lse = lc[b]; del = dc[b]; } else { lse = lse_r[a]; del = del_r[a]; }
          float pv = exp2f(fmaf(s[a][b], p.c2, -lse));
          if (!interior && (cb + b < lo[a] || cb + b >= hi[a])) pv = 0.f;
          s[a][b] = pv;
          dp[a][b] = pv * (dp[a][b] - del) * p.scale;
        }
      if constexpr (KV) {
        store_w<RM>(Ws, ty, tx, s);
        store_w<RM>(W2s, ty, tx, dp);
        __syncwarp();
        tile_accumulate<RM, NC, ZS>(Ws, Zs, ty, tx, acc);
        tile_accumulate<RM, NC, ZS>(W2s, Z2s, ty, tx, acc2);
      } else {
        store_w<RM>(Ws, ty, tx, dp);
        __syncwarp();
        tile_accumulate<RM, NC, ZS>(Ws, Zs, ty, tx, acc);
      }
    }
  }

#pragma unroll
  for (int a = 0; a < RM; ++a) {
    const int r = r0 + ty * RM + a;
    float inv = 1.f;
    if constexpr (MODE == A_FWD) {
      const float lt = half_warp_sum(l[a]);
      inv = lt;
      if (r < N && tx == 0) p.lse_out[r] = m[a] + log2f(lt);
    }
    if (r >= N) continue;
#pragma unroll
    for (int u = 0; u < NC; ++u) {
      const int c = tx + 16 * u;
      if (c < d) {
        p.O1[(int64_t)r * p.ldo1 + c] = MODE == A_FWD ? __fdiv_rn(acc[a][u], inv) : acc[a][u];
        if constexpr (KV) p.O2[(int64_t)r * p.ldo2 + c] = acc2[a][u];
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
  if (a.d <= 16) return dispatch_rm<MODE, 16>(a, stream);
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
