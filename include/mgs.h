/*
 * mgs.h -- C ABI of libmgs.so: the B200 (sm_100a) message-passing hot path of
 * M-GAT-GraphSAGE (GATConv / SAGEConv / global pooling / dense projections and
 * their backward).
 *
 * The reference (JiaCZ-Computational-Biology/M-GAT-GraphSAGE) has no FFI of its
 * own: its hot path is the Python operator API of torch_geometric that its
 * scripts import.  Each entry point below therefore names the reference
 * *call site* whose arithmetic it replaces (paths relative to /root/reference)
 * and the PyG operator chain behind it.
 *
 * Contract
 *   - plain pointers and sizes only; every buffer (inputs, outputs, saved-for-
 *     backward, scratch) is allocated by the caller (PyTorch) in device memory;
 *     the library never allocates, frees or retains pointers;
 *   - every call only enqueues work on the caller's CUDA stream (`stream` is a
 *     cudaStream_t); no internal synchronisation, no host reads of device data;
 *   - return value 0 = enqueued, non-zero = error code below, message through
 *     mgs_last_error_string() (thread-local).  Nothing throws, nothing exits;
 *   - reentrant; the only process-wide state is an atomic launch counter;
 *   - all floating point is fp32, all indices inside the library are int32,
 *     edge_index / batch arrive as int64 exactly as PyG holds them;
 *   - matrices are row-major with an explicit leading dimension in ELEMENTS.
 *
 * Sorted-CSR layout produced by mgs_csr_build
 *   by destination:  rowptr[N+1], col[E] (source of each in-edge), perm[E]
 *                    (original edge id), in-edges of a node in ascending edge id
 *                    == numpy.argsort(dst, kind="stable");
 *   by source:       colptr[N+1], row[E] (destination), permt[E] (original edge
 *                    id), csc_pos[E] (position of the same edge in the
 *                    by-destination order);
 *   "slot" order used by the GAT kernels for per-edge-per-head arrays such as
 *   alpha[(E+N), H]: the k-th in-edge of node i lives at slot rowptr[i] + i + k
 *   and the implicit self loop PyG appends (A.1 step 3) at rowptr[i+1] + i.
 *   In-edges that already are self loops (col == i) are removed by GATConv:
 *   their slot holds 0 and they are skipped.
 */
#ifndef MGS_H_
#define MGS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGS_VERSION 100 /* 0.1.0 */

enum {
  MGS_OK = 0,
  MGS_ERR_INVALID_ARGUMENT = 1,
  MGS_ERR_CUDA = 2,
  MGS_ERR_WORKSPACE_TOO_SMALL = 3,
  MGS_ERR_UNSUPPORTED = 4
};

/* bits set in the device-side `status` word by the index kernels */
#define MGS_STATUS_EDGE_OUT_OF_RANGE 1
#define MGS_STATUS_BATCH_NOT_SORTED 2
#define MGS_STATUS_BATCH_OUT_OF_RANGE 4

typedef void* mgs_stream_t; /* a cudaStream_t */

/* pool modes */
#define MGS_POOL_MAX 0
#define MGS_POOL_MEAN 1
#define MGS_POOL_ADD 2

int mgs_version(void);
const char* mgs_last_error_string(void);
/* number of kernels this library has launched in this process (bench.py's `gpu_launches`) */
uint64_t mgs_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * K0  sorted-CSR / segment-pointer builder (integer, bit-exact).
 * Replaces the COO scatter bookkeeping PyG redoes on every call
 * (torch_geometric.utils.scatter, used by every conv at train.py:117, ablation/model1.py:68,70)
 * and Batch.ptr (train.py:209 DataLoader collation).
 * edge_index: row 0 (sources) at edge_index[0..E), row 1 (destinations) at
 * edge_index[edge_row_stride .. edge_row_stride+E).
 * ------------------------------------------------------------------------------------------ */
size_t mgs_csr_workspace_bytes(int64_t num_nodes, int64_t num_edges);
int mgs_csr_build(const int64_t* edge_index, int64_t edge_row_stride, int64_t num_edges, int64_t num_nodes,
                  int32_t* rowptr, int32_t* col, int32_t* perm,
                  int32_t* colptr, int32_t* row, int32_t* permt, int32_t* csc_pos,
                  int32_t* status, void* workspace, size_t workspace_bytes, mgs_stream_t stream);
/* gptr[g] = first atom of molecule g (batch must be sorted ascending, as Batch collation makes it;
 * test.py:185 / gnnexplainer.py:645 pass zeros).  Empty molecules are allowed. */
int mgs_graph_ptr(const int64_t* batch, int64_t num_nodes, int64_t num_graphs, int32_t* gptr,
                  int32_t* status, mgs_stream_t stream);

/* Device-side expansion of the compact wire format of a batch (m_gat_graphsage_b200/data.py WireBatch; the host ->
 * device copy of the reference's `batch.to(device)`, test.py:188-189): x[n, f] = bit f of bits[n] (the reference's atom
 * features are exactly 0.0 / 1.0: train.py:33-43), edge_index int32 -> int64, batch[n] = molecule of atom n from the
 * segment pointers gptr[B + 1].  One launch. */
int mgs_wire_expand(const uint64_t* bits, int64_t num_nodes, int32_t num_feat, float* x, int64_t ldx,
                    const int32_t* edge_index32, int64_t num_edges, int64_t* edge_index64,
                    const int32_t* gptr, int64_t num_graphs, int64_t* batch, mgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K1  SAGEConv mean aggregation  (train.py:117, ablation/model1.py:70, gnn/graphsage.py:64,67;
 *     PyG: index_select -> scatter_add -> count.clamp(1) -> divide, Appendix A.2).
 * out[i,:] = (sum over in-edges e=(j->i), ascending edge id, of w_e * x[j,:]) / max(indeg(i),1)
 * edge_weight: optional [E] by ORIGINAL edge id (the explainer's sigmoid(edge_mask), A.4), or NULL.
 * ------------------------------------------------------------------------------------------ */
int mgs_sage_aggr_fwd(const float* x, int64_t ldx, int64_t num_nodes, int32_t num_feat,
                      const int32_t* rowptr, const int32_t* col, const int32_t* perm,
                      const float* edge_weight, float* out, int64_t ldo, mgs_stream_t stream);
/* gx[j,:] = sum over out-edges e=(j->i), ascending edge id, of w_e * g[i,:] / max(indeg(i),1)
 * (autograd of the chain above: train.py:248 total_loss.backward(), gnnexplainer.py:650). */
int mgs_sage_aggr_bwd(const float* g, int64_t ldg, int64_t num_nodes, int32_t num_feat,
                      const int32_t* rowptr, const int32_t* colptr, const int32_t* row, const int32_t* permt,
                      const float* edge_weight, float* gx, int64_t ldgx, mgs_stream_t stream);
/* Same, gx = base + (...): the lin_r data gradient of SAGEConv (base, may be gx itself or a strided column block of
 * a wider GEMM output) joins the aggregation gradient in one pass instead of a separate [N, F] add; fp32 addition
 * commutes, so the bits equal autograd's sum of the two gradients. */
int mgs_sage_aggr_bwd_accumulate(const float* g, int64_t ldg, int64_t num_nodes, int32_t num_feat,
                                 const int32_t* rowptr, const int32_t* colptr, const int32_t* row,
                                 const int32_t* permt, const float* edge_weight, const float* base, int64_t ldbase,
                                 const float* relu_mask /* optional: gx = relu_mask <= 0 ? 0 : gx, i.e. the backward of
                                 the ReLU whose OUTPUT (relu_mask) is this layer's input, fused into the producer of its
                                 gradient (model1.py:69 self.relu(conv1(..))) */, int64_t ldmask,
                                 const uint32_t* relu_bits /* the same mask as bits (see mgs_gat_aggr_fwd); at most one of
                                 the two forms */, int32_t bits_words,
                                 float* gx, int64_t ldgx, mgs_stream_t stream);
/* d_edge_weight[e] = < g[i,:] / max(indeg(i),1), x[j,:] >   (explainer edge-mask gradient, A.4) */
int mgs_sage_aggr_bwd_edge_weight(const float* g, int64_t ldg, const float* x, int64_t ldx,
                                  int64_t num_nodes, int32_t num_feat,
                                  const int32_t* rowptr, const int32_t* col, const int32_t* perm,
                                  float* d_edge_weight, mgs_stream_t stream);
/* Neighbourhood SUM (GCNConv gnn/gcn.py:46-48, gnn/gat-gcn.py:58; GINConv gnn/gin.py:64-77; PyG: index_select ->
 * (* edge_weight) -> scatter_add):  dst[i,:] = (base ? base[i,:] : 0) + sum over entries p in [ptr[i], ptr[i+1]), in
 * order, of w[eid[p]] * src[idx[p],:].  Forward: (rowptr, col, perm); backward of the same op: (colptr, row,
 * permt).  `base` may alias `src` (GIN's (1 + eps) x_i with eps = 0, GCN's self loop) or `dst`. */
int mgs_sum_aggr(const float* src, int64_t lds, int64_t num_nodes, int32_t num_feat, const int32_t* ptr,
                 const int32_t* idx, const int32_t* eid, const float* edge_weight, const float* base, int64_t ldbase,
                 float* dst, int64_t ldd, mgs_stream_t stream);
/* Self test of the in-kernel replacement for IEEE division by an in-degree (csrc/common.cuh
 * div_by_count): compares it with __fdiv_rn for every fp32 bit pattern 0, stride, 2*stride, ... as
 * dividend and every integer divisor in [count_lo, count_hi]; *mismatches (device, uint64) receives the
 * number of differing bit patterns (NaN payloads excluded).  Test hook, not on the hot path. */
int mgs_selftest_div(int32_t count_lo, int32_t count_hi, uint64_t stride, unsigned long long* mismatches,
                     mgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K2  GATConv message passing  (ablation/model1.py:68, gnn/gat.py:63,65; Appendix A.1 steps 2-8).
 * xh = lin(x) viewed [N, H, C] comes from mgs_linear_fwd.
 * ------------------------------------------------------------------------------------------ */
/* a_src[n,h] = <xh[n,h,:], att_src[h,:]>, a_dst likewise (A.1 step 2) */
int mgs_gat_scores_fwd(const float* xh, int64_t ld, int64_t num_nodes, int32_t heads, int32_t channels,
                       const float* att_src, const float* att_dst, float* a_src, float* a_dst,
                       mgs_stream_t stream);
/* alpha[slot,h] = softmax over the in-edges (+ self loop) of leaky_relu(a_src[j,h] + a_dst[i,h])
 * with max subtraction and the +1e-16 denominator of torch_geometric.utils.softmax (A.1 steps 3-5) */
int mgs_gat_alpha_fwd(const float* a_src, const float* a_dst, int64_t num_nodes, int32_t heads,
                      const int32_t* rowptr, const int32_t* col, float negative_slope,
                      float* alpha, mgs_stream_t stream);
/* out[i,h,:] = sum over slots of alpha_used[slot,h] * w_e * xh[j,h,:]  (+ bias)   (A.1 steps 7-8, concat layout)
 * alpha_used = alpha, or alpha * dropout keep-mask (A.1 step 6, gnn/gat.py:54-55) */
int mgs_gat_aggr_fwd(const float* xh, int64_t ld, int64_t num_nodes, int32_t heads, int32_t channels,
                     const float* alpha_used, const int32_t* rowptr, const int32_t* col, const int32_t* perm,
                     const float* edge_weight, const float* bias, float* out, int64_t ldo,
                     int32_t activation /* applied to out: 0 none, 1 ReLU (model1.py:68-69), 2 ELU (gnn/gat.py:63) */,
                     uint32_t* relu_bits /* optional, ReLU only: out > 0 as one bit per element, bits_words uint32 per row in the
                     lane order of the aggregation kernels (consumed by mgs_sage_aggr_bwd_accumulate) */, int32_t bits_words,
                     mgs_stream_t stream);
/* backward, stage 1 (per destination): d alpha = <g_i, xh_j> (x mask, x w_e), softmax Jacobian,
 * leaky_relu'  ->  dr[slot,h], da_dst[i,h] = sum_slots dr;  optional d_edge_weight[e] */
int mgs_gat_bwd_edge(const float* g, int64_t ldg, const float* xh, int64_t ld,
                     int64_t num_nodes, int32_t heads, int32_t channels,
                     const float* alpha, const float* alpha_mask, const float* a_src, const float* a_dst,
                     float negative_slope, const int32_t* rowptr, const int32_t* col, const int32_t* perm,
                     const float* edge_weight, float* dr, float* da_dst, float* d_edge_weight,
                     mgs_stream_t stream);
/* backward, stage 2 (per source): da_src[j,h] = sum over out-slots of dr;
 * dxh[j,h,:] = sum over out-slots alpha_used*w_e*g[i,h,:] + da_src[j,h]*att_src[h,:] + da_dst[j,h]*att_dst[h,:]
 * att_src = att_dst = NULL: the scores were inputs of the node (mgs_proj_fwd computed them from x): no rank-1 term */
int mgs_gat_bwd_node(const float* g, int64_t ldg, int64_t num_nodes, int32_t heads, int32_t channels,
                     const float* alpha_used, const float* dr, const float* da_dst,
                     const float* att_src, const float* att_dst,
                     const int32_t* rowptr, const int32_t* colptr, const int32_t* row,
                     const int32_t* csc_pos, const int32_t* permt, const float* edge_weight,
                     float* dxh, int64_t lddxh, float* da_src, mgs_stream_t stream);
/* backward, stage 3: datt_src[h,c] = sum_n da_src[n,h]*xh[n,h,c], datt_dst likewise (deterministic two-stage) */
size_t mgs_gat_bwd_att_workspace_bytes(int32_t heads, int32_t channels);
int mgs_gat_bwd_att(const float* xh, int64_t ld, int64_t num_nodes, int32_t heads, int32_t channels,
                    const float* da_src, const float* da_dst, float* datt_src, float* datt_dst,
                    void* workspace, size_t workspace_bytes, mgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K4 small-K projection with fused attention scores (GATConv(35, 35, heads=10).lin of ablation/model1.py:57,68):
 *   [out0 | out1 | out2] = x[N,K] . [w ; u1 ; u2]^T (+ bias on out0),   K <= 64, n0 + n1 + n2 <= 384
 *   with u1 = U_src, u2 = U_dst (U[h,:] = sum_c att[h,c] W[hC+c,:]) this yields xh, a_src, a_dst in one pass.
 * wgrad: [dw ; du1 ; du2] = [g0 | g1 | g2]^T . x,   K <= 36, n0 + n1 + n2 <= 384 (deterministic split over atoms).
 * ------------------------------------------------------------------------------------------ */
int mgs_proj_fwd(const float* x, int64_t ldx, int64_t num_rows, int32_t K, const float* w, int64_t ldw, int32_t n0,
                 const float* u1, int64_t ldu1, int32_t n1, const float* u2, int64_t ldu2, int32_t n2,
                 const float* bias, float* out0, int64_t ld0, float* out1, int64_t ld1, float* out2, int64_t ld2,
                 mgs_stream_t stream);
size_t mgs_proj_wgrad_workspace_bytes(int32_t K, int32_t n_total);
int mgs_proj_wgrad(const float* g0, int64_t ldg0, int32_t n0, const float* g1, int64_t ldg1, int32_t n1,
                   const float* g2, int64_t ldg2, int32_t n2, const float* x, int64_t ldx, int64_t num_rows,
                   int32_t K, float* dw, int64_t lddw, float* du1, int64_t lddu1, float* du2, int64_t lddu2,
                   void* workspace, size_t workspace_bytes, mgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K3  segmented global pooling  (train.py:119, ablation/model1.py:72, gnn/gat.py:67, gnn/graphsage.py:68;
 *     PyG scatter(reduce='max'|'mean'|'sum') over `batch`, Appendix A.3).  Empty molecule -> 0.
 * ------------------------------------------------------------------------------------------ */
int mgs_pool_fwd(const float* x, int64_t ldx, const int32_t* gptr, int64_t num_graphs, int32_t num_feat,
                 int32_t mode, float* out, int64_t ldo, mgs_stream_t stream);
/* max: gradient split evenly over exact ties, the zero-initialised destination counting as one more
 * tie when the maximum is exactly 0 (ATen scatter_reduce 'amax' backward); mean: g / max(count,1). */
int mgs_pool_bwd(const float* g, int64_t ldg, const float* x, int64_t ldx, const float* out, int64_t ldo,
                 const int32_t* gptr, int64_t num_graphs, int32_t num_feat, int32_t mode,
                 float* gx, int64_t ldgx, mgs_stream_t stream);
/* Max and mean pooling of the same x in one pass: out[b, 0:F] = max, out[b, F:2F] = mean (ldo >= 2F); the
 * backward takes g[b, 0:2F] in the same layout and writes gx = d max + d mean.  Replaces the pair
 * global_max_pool(x, batch) / global_mean_pool(x, batch) of ablation/model1.py:72 (and `model 2.py`, `model 3.py`,
 * gnn/gat-gcn.py:71) when both are applied to the same tensor. */
/* `ties` (optional, [B, F] dense): number of rows attaining the maximum (+1 when the maximum is 0, the amax
 * destination rule), counted by the forward pass so that the backward reads x once instead of twice. */
int mgs_pool_maxmean_fwd(const float* x, int64_t ldx, const int32_t* gptr, int64_t num_graphs, int32_t num_feat,
                         float* out, int64_t ldo, float* ties, mgs_stream_t stream);
int mgs_pool_maxmean_bwd(const float* g, int64_t ldg, const float* x, int64_t ldx, const float* out, int64_t ldo,
                         const int32_t* gptr, int64_t num_graphs, int32_t num_feat, float* gx, int64_t ldgx,
                         const float* ties, int32_t relu_mask /* != 0: gx = x <= 0 ? 0 : gx (x is a ReLU output,
                         model1.py:71-72: the ReLU's backward rides on the pass that reads x anyway) */,
                         mgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K5  streaming all-pairs attention of ModifiedGATLayer (train.py:87-99; copies in test.py:62-84 and
 *     gnnexplainer.py:54-76): replaces `softmax(matmul(Q, K_new^T) / sqrt(d)) @ V`, whose [N, N] score matrix
 *     the reference materialises, by a tiled pass with a running softmax -- nothing of size N x N exists.
 *       out[b, :] = sum_i softmax_i(<qry[b], key[i]> * scale) val[i, :]     qry = K_new, key = Q, val = V
 *     seg / gptr null: over all N atoms of the batch (the reference's semantics); given (seg[N] = molecule of
 *     every atom, ascending; gptr[B+1] = atom range of every molecule): over the atoms of b's molecule only
 *     (what the one-molecule-at-a-time scripts compute).  d <= 64.  lse2[N] (log2 of the softmax denominator)
 *     is saved for the backward; delta[N] = <gout[b], out[b]> is computed by the caller.  The residual `+ V`
 *     of the layer stays with the caller.
 * ------------------------------------------------------------------------------------------ */
int mgs_attn_fwd(const float* qry, int64_t ldq, const float* key, int64_t ldk, const float* val, int64_t ldv,
                 int64_t num_nodes, int32_t d, float scale, const int32_t* seg, const int32_t* gptr,
                 float* out, int64_t ldo, float* lse2, mgs_stream_t stream);
int mgs_attn_bwd(const float* qry, int64_t ldq, const float* key, int64_t ldk, const float* val, int64_t ldv,
                 int64_t num_nodes, int32_t d, float scale, const int32_t* seg, const int32_t* gptr,
                 const float* lse2, const float* delta, const float* gout, int64_t ldg,
                 float* dqry, int64_t lddq, float* dkey, int64_t lddk, float* dval, int64_t lddv,
                 mgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K4  dense projections / readout MLP  (GATConv.lin, SAGEConv.lin_l / lin_r, fc_g1 / fc_g2 / out:
 *     train.py:107-111,120-123, ablation/model1.py:59-64,73-76; ATen addmm).
 * c[M,Nout] = a[M,K] w[Nout,K]^T (+ a2[M,K2] w2[Nout,K2]^T) (+ bias) (ReLU if relu != 0)
 * fp32-accurate (no single-pass TF32/BF16): SURVEY.md section 7 "Tensor cores vs 1e-5".
 * Tiles of >= 128 rows run on tcgen05 tensor cores as 3xTF32 (hi/lo split, fp32 TMEM accumulators); the
 * workspace holds the weight operand pre-split and pre-swizzled for 1-D TMA bulk copies.
 * ------------------------------------------------------------------------------------------ */
size_t mgs_linear_fwd_workspace_bytes(int64_t M, int32_t K, int32_t Nout, int32_t K2);
int mgs_linear_fwd(const float* a, int64_t lda, int64_t M, int32_t K,
                   const float* w, int64_t ldw, int32_t Nout, const float* bias,
                   const float* a2, int64_t lda2, int32_t K2, const float* w2, int64_t ldw2,
                   float* c, int64_t ldc, int32_t relu,
                   void* workspace, size_t workspace_bytes, mgs_stream_t stream);
/* da[M,K] = g[M,Nout] w[Nout,K] */
size_t mgs_linear_dgrad_workspace_bytes(int64_t M, int32_t Nout, int32_t K);
int mgs_linear_dgrad(const float* g, int64_t ldg, int64_t M, int32_t Nout,
                     const float* w, int64_t ldw, int32_t K, float* da, int64_t ldda,
                     void* workspace, size_t workspace_bytes, mgs_stream_t stream);
/* dw[Nout,K] = g[M,Nout]^T a[M,K]   (split over M, deterministic reduction) */
size_t mgs_linear_wgrad_workspace_bytes(int64_t M, int32_t Nout, int32_t K);
int mgs_linear_wgrad(const float* g, int64_t ldg, int64_t M, int32_t Nout,
                     const float* a, int64_t lda, int32_t K, float* dw, int64_t lddw,
                     void* workspace, size_t workspace_bytes, mgs_stream_t stream);
/* out[n] = sum_m g[m,n]   (bias gradients; deterministic two-stage) */
size_t mgs_colsum_workspace_bytes(int32_t Nout);
int mgs_colsum(const float* g, int64_t ldg, int64_t M, int32_t Nout, float* out,
               void* workspace, size_t workspace_bytes, mgs_stream_t stream);

/* da[M, K] = g0[M, N0] w0[N0, K] + g1[M, N1] w1[N1, K]  (two data gradients as ONE GEMM over K' = N0 + N1), optionally
 * followed by the backward of a fused ReLU: da = 0 where the bit of relu_bits ([M, bits_words] uint32 in the row layout of
 * mgs_gat_aggr_fwd's relu_bits, bits_v floats per lane) is clear.  SAGEConv backward (reference: gnn/graphsage.py:52-56,
 * ablation/model1.py:70): gx = relu'(x) (g W_r + mean-aggregation-backward(g) W_l).  TMA-fed kernel only: returns
 * MGS_ERR_UNSUPPORTED when the operands do not qualify (16-byte aligned rows, K a 176-column-tile width).  colsum_out
 * (optional, [K]): column sums of the masked result, from the GEMM epilogue -- the bias gradient of the layer that produced x
 * (GATConv's `bias`, ablation/model1.py:68) without another pass over the [M, K] gradient. */
size_t mgs_linear_dgrad2_workspace_bytes(int64_t M, int32_t N0, int32_t N1, int32_t K);
int mgs_linear_dgrad2(const float* g0, int64_t ldg0, int32_t N0, const float* w0, int64_t ldw0, const float* g1,
                      int64_t ldg1, int32_t N1, const float* w1, int64_t ldw1, int64_t M, int32_t K, float* da,
                      int64_t ldda, const uint32_t* relu_bits, int32_t bits_words, int32_t bits_v, float* colsum_out,
                      void* workspace, size_t workspace_bytes, mgs_stream_t stream);

/* Score weights of the fused GATConv projection (mgs_proj_fwd): U_src[h, :] = sum_c att_src[h, c] W[hC + c, :] (same for
 * dst), [H, K] each, and their backward: dw[hC + c, :] = att_src[h, c] du_src[h, :] + att_dst[h, c] du_dst[h, :] (overwritten),
 * datt_src[h, c] = <du_src[h, :], W[hC + c, :]>.  att / datt are [H * C] contiguous. */
int mgs_gat_u_fwd(const float* w, int64_t ldw, const float* att_src, const float* att_dst, int32_t H, int32_t C, int32_t K,
                  float* u_src, float* u_dst, mgs_stream_t stream);
int mgs_gat_u_bwd(const float* w, int64_t ldw, const float* att_src, const float* att_dst, const float* du_src,
                  const float* du_dst, int32_t H, int32_t C, int32_t K, float* dw, int64_t lddw, float* datt_src,
                  float* datt_dst, mgs_stream_t stream);

/* Optimiser step of the reference scripts (train.py:216-222, ablation/model1.py:113: torch.optim.Adam, L2 weight decay,
 * no amsgrad) over `count` parameter tensors in ONE launch per 24 tensors: params / grads / exp_avg / exp_avg_sq are host
 * arrays of device pointers, numel their element counts; `step` is the 1-based step number (bias corrections). */
int mgs_adam_step(int32_t count, float* const* params, const float* const* grads, float* const* exp_avg,
                  float* const* exp_avg_sq, const int64_t* numel, double lr, double beta1, double beta2, double eps,
                  double weight_decay, int64_t step, mgs_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MGS_H_ */
