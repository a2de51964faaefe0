"""CPU oracle for the M-GAT-GraphSAGE message-passing hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product path (``m_gat_graphsage_b200``)
never imports anything from ``oracle/``.

PARITY UNPINNED.  The arithmetic of the hot path lives in the third-party
package ``torch_geometric`` (PyPI ``torch-geometric``; the reference pins no
version -- ``/root/reference/README.md:22-32`` lists torch 2.4.1 only; the
``torch_geometric.explain`` imports at ``gnnexplainer.py:7-8`` imply >= 2.3).
It is neither vendored under ``/root/reference`` nor installable here (no
network), and the reference ships no tests, golden vectors or weights.  This
file therefore *restates the published PyG algorithm* op for op (same
scatter order, same epsilons, same self-loop placement) in plain PyTorch on
the CPU and is cross-checked by ``oracle/dense_check.py`` (an independent
dense-adjacency formulation) and fp64 ``gradcheck`` in ``tests/``.  What *is*
pinned against the reference's own source is the model wiring and
``ModifiedGATLayer`` (plain-torch code that does run here): see
``tests/golden/make_golden.py``.

Call sites this follows (``/root/reference``):
  * ``GATConv(35, 35, heads=10)``            ablation/model1.py:57,68
  * ``GATConv(.., heads=10, dropout=0.2)``   gnn/gat.py:54-55,63,65
  * ``SAGEConv(in, out)``                    train.py:106,117; ablation/model1.py:58,70;
                                             gnn/graphsage.py:53-54,64,67
  * ``global_max_pool`` / ``global_mean_pool`` train.py:119; ablation/model1.py:72
  * explain hook (A.4)                        gnnexplainer.py:620-631
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# A.0  scatter  (torch_geometric.utils.scatter, torch>=2 CPU path)
# ----------------------------------------------------------------------------
def scatter(src: torch.Tensor, index: torch.Tensor, dim_size: int, reduce: str) -> torch.Tensor:
    """``scatter(src, index, dim=0, dim_size, reduce)`` along dim 0.

    sum  : zeros.scatter_add_  (CPU accumulates in ascending row order)
    mean : sum / clamp(count, 1)       (true fp32 division)
    max  : zeros.scatter_reduce_('amax', include_self=False)  (empty -> 0)
    """
    shape = (dim_size,) + tuple(src.shape[1:])
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    if reduce in ("sum", "add"):
        return src.new_zeros(shape).scatter_add_(0, idx, src)
    if reduce == "mean":
        count = src.new_zeros(dim_size).scatter_add_(0, index, src.new_ones(src.size(0)))
        count = count.clamp(min=1)
        out = src.new_zeros(shape).scatter_add_(0, idx, src)
        return out / count.view(-1, *([1] * (src.dim() - 1)))
    if reduce == "max":
        return src.new_zeros(shape).scatter_reduce_(0, idx, src, reduce="amax", include_self=False)
    raise ValueError(reduce)


def segment_softmax(src: torch.Tensor, index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """``torch_geometric.utils.softmax`` (A.1 step 5)."""
    src_max = scatter(src.detach(), index, num_nodes, "max")
    out = (src - src_max.index_select(0, index)).exp()
    out_sum = scatter(out, index, num_nodes, "sum") + 1e-16
    return out / out_sum.index_select(0, index)


def remove_self_loops(edge_index: torch.Tensor):
    mask = edge_index[0] != edge_index[1]
    return edge_index[:, mask], mask


def add_self_loops(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index, loop.unsqueeze(0).repeat(2, 1)], dim=1)


def glorot_(t: torch.Tensor) -> None:
    stdv = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-stdv, stdv)


class _MessagePassingHooks:
    """The three attributes PyG's ``Explainer`` sets on every MessagePassing module (A.4)."""
    _explain: bool = False
    _edge_mask: Optional[torch.Tensor] = None
    _apply_sigmoid: bool = True

    def _mask_messages(self, msg: torch.Tensor, num_real_edges: int) -> torch.Tensor:
        if not self._explain or self._edge_mask is None:
            return msg
        m = self._edge_mask
        if self._apply_sigmoid:
            m = m.sigmoid()
        # self-loop messages appended by GATConv get weight 1 (A.4)
        if msg.size(0) != m.size(0):
            m = torch.cat([m, m.new_ones(msg.size(0) - m.size(0))])
        return msg * m.view(-1, *([1] * (msg.dim() - 1)))


# ----------------------------------------------------------------------------
# A.2  SAGEConv
# ----------------------------------------------------------------------------
class SAGEConv(nn.Module, _MessagePassingHooks):
    def __init__(self, in_channels: int, out_channels: int, aggr: str = "mean",
                 normalize: bool = False, root_weight: bool = True,
                 project: bool = False, bias: bool = True):
        super().__init__()
        assert aggr == "mean" and not normalize and not project
        self.in_channels, self.out_channels, self.root_weight = in_channels, out_channels, root_weight
        self.lin_l = nn.Linear(in_channels, out_channels, bias=bias)
        if root_weight:
            self.lin_r = nn.Linear(in_channels, out_channels, bias=False)

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor, size=None) -> torch.Tensor:
        src, dst = edge_index[0], edge_index[1]
        msg = x.index_select(0, src)
        msg = self._mask_messages(msg, src.numel())
        agg = scatter(msg, dst, x.size(0), "mean")
        out = self.lin_l(agg)
        if self.root_weight:
            out = out + self.lin_r(x)
        return out


# ----------------------------------------------------------------------------
# GCNConv / GINConv (gnn/gcn.py:46-48, gnn/gat-gcn.py:58, gnn/gin.py:64-77) -- PyG's published algorithms
# ----------------------------------------------------------------------------
def gcn_norm(edge_index: torch.Tensor, edge_weight: Optional[torch.Tensor], num_nodes: int, improved: bool = False,
             add_self_loops_: bool = True, dtype=torch.float32):
    """torch_geometric.nn.conv.gcn_conv.gcn_norm: ``add_remaining_self_loops`` (existing self loops keep their
    weight, every other node gets one of weight 1, or 2 when ``improved``), ``deg = scatter_add(w, target)``,
    ``norm_e = deg^-1/2[source] * w_e * deg^-1/2[target]`` (inf -> 0)."""
    fill = 2.0 if improved else 1.0
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=dtype)
    if add_self_loops_:
        src, dst = edge_index
        loop = src == dst
        loop_w = torch.full((num_nodes,), fill, dtype=edge_weight.dtype)
        loop_w[src[loop]] = edge_weight[loop]
        ar = torch.arange(num_nodes)
        edge_index = torch.cat([edge_index[:, ~loop], torch.stack([ar, ar])], dim=1)
        edge_weight = torch.cat([edge_weight[~loop], loop_w])
    src, dst = edge_index
    deg = scatter(edge_weight, dst, num_nodes, "sum")
    dinv = deg.pow(-0.5)
    dinv = dinv.masked_fill(dinv == float("inf"), 0.0)
    return edge_index, dinv[src] * edge_weight * dinv[dst]


class GCNConv(nn.Module, _MessagePassingHooks):
    def __init__(self, in_channels: int, out_channels: int, improved: bool = False, cached: bool = False,
                 add_self_loops: bool = True, normalize: bool = True, bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.add_self_loops, self.normalize = improved, add_self_loops, normalize
        self.lin = nn.Linear(in_channels, out_channels, bias=False)
        glorot_(self.lin.weight)
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def forward(self, x, edge_index, edge_weight=None):
        if self.normalize:
            edge_index, edge_weight = gcn_norm(edge_index, edge_weight, x.size(0), self.improved, self.add_self_loops,
                                               x.dtype)
        xw = self.lin(x)
        msg = xw.index_select(0, edge_index[0])
        if edge_weight is not None:
            msg = msg * edge_weight.view(-1, 1)
        out = scatter(msg, edge_index[1], x.size(0), "sum")
        return out if self.bias is None else out + self.bias


class GINConv(nn.Module, _MessagePassingHooks):
    def __init__(self, nn_: nn.Module, eps: float = 0.0, train_eps: bool = False):
        super().__init__()
        self.nn = nn_
        if train_eps:
            self.eps = nn.Parameter(torch.tensor([float(eps)]))
        else:
            self.register_buffer("eps", torch.tensor([float(eps)]))

    def forward(self, x, edge_index, size=None):
        out = scatter(x.index_select(0, edge_index[0]), edge_index[1], x.size(0), "sum")
        return self.nn(out + (1.0 + self.eps) * x)


# ----------------------------------------------------------------------------
# A.1  GATConv
# ----------------------------------------------------------------------------
class GATConv(nn.Module, _MessagePassingHooks):
    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, concat: bool = True,
                 negative_slope: float = 0.2, dropout: float = 0.0, add_self_loops: bool = True,
                 edge_dim=None, fill_value="mean", bias: bool = True):
        super().__init__()
        assert edge_dim is None
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.negative_slope, self.dropout = concat, negative_slope, dropout
        self.add_self_loops = add_self_loops
        self.lin = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(heads * out_channels if concat else out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()
        #: optional injected attention-dropout keep-mask [E', H] already scaled by 1/(1-p);
        #: rows ordered like the self-loop-augmented edge list.  Test hook only.
        self._injected_alpha_mask: Optional[torch.Tensor] = None

    def reset_parameters(self) -> None:
        glorot_(self.lin.weight)
        glorot_(self.att_src)
        glorot_(self.att_dst)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor, edge_attr=None, size=None,
                return_attention_weights=None):
        H, C, N = self.heads, self.out_channels, x.size(0)
        xh = self.lin(x).view(N, H, C)
        a_s = (xh * self.att_src).sum(dim=-1)
        a_d = (xh * self.att_dst).sum(dim=-1)
        num_real = edge_index.size(1)
        if self.add_self_loops:
            edge_index, _ = remove_self_loops(edge_index)
            num_real = edge_index.size(1)
            edge_index = add_self_loops(edge_index, N)
        src, dst = edge_index[0], edge_index[1]
        e = F.leaky_relu(a_s.index_select(0, src) + a_d.index_select(0, dst), self.negative_slope)
        alpha = segment_softmax(e, dst, N)
        if self._injected_alpha_mask is not None:
            alpha = alpha * self._injected_alpha_mask
        else:
            alpha = F.dropout(alpha, p=self.dropout, training=self.training)
        msg = alpha.unsqueeze(-1) * xh.index_select(0, src)
        if self._explain and self._edge_mask is not None:
            m = self._edge_mask.sigmoid() if self._apply_sigmoid else self._edge_mask
            if m.size(0) != num_real:
                raise ValueError("edge_mask does not match the self-loop-free edge list")
            m = torch.cat([m, m.new_ones(msg.size(0) - num_real)])
            msg = msg * m.view(-1, 1, 1)
        out = scatter(msg, dst, N, "sum")
        out = out.view(N, H * C) if self.concat else out.mean(dim=1)
        if self.bias is not None:
            out = out + self.bias
        if return_attention_weights:
            return out, (edge_index, alpha)
        return out


# ----------------------------------------------------------------------------
# A.3  pools
# ----------------------------------------------------------------------------
def _num_graphs(batch: torch.Tensor, size: Optional[int]) -> int:
    return int(batch.max()) + 1 if size is None else size


def global_max_pool(x, batch, size=None):
    if batch is None:
        return x.max(dim=0, keepdim=True)[0]
    return scatter(x, batch, _num_graphs(batch, size), "max")


def global_mean_pool(x, batch, size=None):
    if batch is None:
        return x.mean(dim=0, keepdim=True)
    return scatter(x, batch, _num_graphs(batch, size), "mean")


def global_add_pool(x, batch, size=None):
    if batch is None:
        return x.sum(dim=0, keepdim=True)
    return scatter(x, batch, _num_graphs(batch, size), "sum")


# ----------------------------------------------------------------------------
# K0 oracle: integer, bit-exact (numpy restatement of stable argsort/bincount/cumsum)
# ----------------------------------------------------------------------------
def csr_oracle(edge_index: torch.Tensor, num_nodes: int):
    """Sorted-CSR by destination + CSC by source, SURVEY.md section 8 row a2."""
    import numpy as np
    ei = edge_index.cpu().numpy()
    src, dst = ei[0], ei[1]
    perm = np.argsort(dst, kind="stable").astype(np.int32)
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(dst, minlength=num_nodes))]).astype(np.int32)
    col = src[perm].astype(np.int32)
    permt = np.argsort(src, kind="stable").astype(np.int32)
    colptr = np.concatenate([[0], np.cumsum(np.bincount(src, minlength=num_nodes))]).astype(np.int32)
    row = dst[permt].astype(np.int32)
    inv = np.empty(perm.shape[0], dtype=np.int32)
    inv[perm] = np.arange(perm.shape[0], dtype=np.int32)
    csc_pos = inv[permt]
    return dict(rowptr=rowptr, col=col, perm=perm, colptr=colptr, row=row, permt=permt, csc_pos=csc_pos)


def graph_ptr_oracle(batch: torch.Tensor, num_graphs: int):
    import numpy as np
    b = batch.cpu().numpy()
    return np.concatenate([[0], np.cumsum(np.bincount(b, minlength=num_graphs))]).astype(np.int32)


# ----------------------------------------------------------------------------
# a1 oracle: Batch collation (torch_geometric.data.Batch.from_data_list)
# ----------------------------------------------------------------------------
def collate_oracle(graphs):
    """graphs: list of (x[n,F], edge_index[2,e]) -> x, edge_index, batch, ptr."""
    xs, eis, bs, ptr, off = [], [], [], [0], 0
    for g, (x, ei) in enumerate(graphs):
        xs.append(x)
        eis.append(ei + off)
        bs.append(torch.full((x.size(0),), g, dtype=torch.long))
        off += x.size(0)
        ptr.append(off)
    return torch.cat(xs), torch.cat(eis, dim=1), torch.cat(bs), torch.tensor(ptr)


# ------------------------------------------------------------------------------------------------
# ModifiedGATLayer attention core (train.py:96-98), dense restatement
# ------------------------------------------------------------------------------------------------
def modified_gat_attention(q: torch.Tensor, k_new: torch.Tensor, v: torch.Tensor,
                           batch: Optional[torch.Tensor] = None) -> torch.Tensor:
    """train.py:96-98 with the broadcasting written out: ``matmul(Q [N,d], K_new^T [N,d,1])`` is
    ``scores[b, i] = <Q[i], K_new[b]>``, the softmax runs over ``i`` (``dim=-1`` after the squeeze) and the
    output is ``weights @ V + V``.  ``batch`` given: atoms of other molecules are excluded from the softmax (what
    the scripts that feed one molecule at a time compute)."""
    scores = (k_new @ q.t()) / (k_new.size(-1) ** 0.5)
    if batch is not None:
        scores = scores.masked_fill(batch.view(-1, 1) != batch.view(1, -1), float("-inf"))
    return torch.softmax(scores, dim=-1) @ v + v
