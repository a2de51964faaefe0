"""Drop-in mirror of the ``torch_geometric.nn`` names the reference scripts import
(SURVEY.md section 8b): ``GATConv``, ``SAGEConv``, ``global_max_pool``, ``global_mean_pool``,
``global_add_pool``.  Same constructor arguments, same ``forward(x, edge_index)`` /
``pool(x, batch)`` signatures, same parameter names (``lin.weight``, ``att_src``, ``att_dst``,
``bias``; ``lin_l.weight``, ``lin_l.bias``, ``lin_r.weight``) so reference checkpoints load with
``strict=True`` (/root/reference/test.py:160-164, gnnexplainer.py:1354-1360).

All arithmetic goes through ``libmgs.so``; CPU tensors are rejected (no fallback).
"""
from __future__ import annotations

import math
import weakref
from typing import Optional

import torch
import torch.nn as nn

from . import functional as F_
from .graph import graph_index, graph_ptr, require_cuda, resolve_num_graphs
from .lazy import PendingActivation, activation_fusion_enabled, set_activation_fusion  # noqa: F401


def _apply_activation(out: torch.Tensor, activation: Optional[str]) -> torch.Tensor:
    if activation is None:
        return out
    return torch.relu(out) if activation == "relu" else torch.nn.functional.elu(out)


class Linear(nn.Module):
    """``torch_geometric.nn.dense.linear.Linear`` (weight ``[out, in]``, optional bias) on K4."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True,
                 weight_initializer: Optional[str] = None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight_initializer = weight_initializer
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self) -> None:
        if self.weight_initializer == "glorot":
            stdv = math.sqrt(6.0 / (self.weight.size(-2) + self.weight.size(-1)))
            nn.init.uniform_(self.weight, -stdv, stdv)
        else:  # PyG default == torch.nn.Linear default
            nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1.0 / math.sqrt(self.in_channels) if self.in_channels > 0 else 0.0
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F_.linear(x, self.weight, self.bias)

    def extra_repr(self) -> str:
        return f"{self.in_channels}, {self.out_channels}, bias={self.bias is not None}"


class MessagePassing(nn.Module):
    """Carrier of the three attributes PyG's ``Explainer`` sets on every message-passing layer
    (SURVEY.md Appendix A.4): when ``_explain`` is on, messages are multiplied by
    ``sigmoid(_edge_mask)`` per original edge and the mask receives a gradient."""

    def __init__(self, aggr: Optional[str] = "add", flow: str = "source_to_target", node_dim: int = -2, **kwargs):
        # the arguments of PyG's base class are accepted (gnn/chebnet.py:52 subclasses it with aggr='add' and then
        # does its own dense algebra); there is no generic propagate() here
        super().__init__()
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim
        self._explain: bool = False
        self._edge_mask: Optional[torch.Tensor] = None
        self._apply_sigmoid: bool = True

    @property
    def explain(self) -> bool:
        return self._explain

    @explain.setter
    def explain(self, value: bool) -> None:
        self._explain = bool(value)

    def _edge_weight(self, num_edges: int) -> Optional[torch.Tensor]:
        if not self._explain or self._edge_mask is None:
            return None
        m = self._edge_mask
        if m.numel() != num_edges:
            raise ValueError(f"edge_mask has {m.numel()} entries but the graph has {num_edges} edges")
        return m.sigmoid() if self._apply_sigmoid else m


class SAGEConv(MessagePassing):
    """GraphSAGE layer, mean aggregation (reference: train.py:106,117; ablation/model1.py:58,70;
    gnn/graphsage.py:53-54).  ``out_i = W_l mean_{j->i} x_j + b_l + W_r x_i`` (Appendix A.2):
    K1 gather + one fused two-operand K4 GEMM."""

    def __init__(self, in_channels: int, out_channels: int, aggr: str = "mean", normalize: bool = False,
                 root_weight: bool = True, project: bool = False, bias: bool = True,
                 activation: Optional[str] = None, **kwargs):
        super().__init__()
        if activation not in (None, "relu", "elu"):
            raise ValueError("activation must be None, 'relu' or 'elu'")
        #: extension of PyG's signature: activation applied to the layer's output, fused into the projection's
        #: epilogue (what every reference model does next: `self.relu(self.conv2(x, edge_index))`, model1.py:70-71).
        #: With the peephole of m_gat_graphsage_b200.lazy on, the same fusion happens for unmodified model code.
        self.activation = activation
        if aggr != "mean":
            raise NotImplementedError("only aggr='mean' (the reference's setting) is implemented")
        if project:
            raise NotImplementedError("project=True is not used by the reference")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.normalize, self.root_weight = normalize, root_weight
        self.lin_l = Linear(in_channels, out_channels, bias=bias)
        if root_weight:
            self.lin_r = Linear(in_channels, out_channels, bias=False)

    def reset_parameters(self) -> None:
        self.lin_l.reset_parameters()
        if self.root_weight:
            self.lin_r.reset_parameters()

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor, size=None) -> torch.Tensor:
        require_cuda(x, "SAGEConv input x")
        x = F_.real(x)
        graph = graph_index(edge_index, x.size(0))
        ew = self._edge_weight(graph.num_edges)
        if (ew is None and self.root_weight and not self.normalize and x.dim() == 2 and x.dtype == torch.float32
                and F_.stream_width_ok(x.size(1))):
            w_l, b_l, w_r = self.lin_l.weight, self.lin_l.bias, self.lin_r.weight
            if self.activation in (None, "relu"):
                if self.activation is None and activation_fusion_enabled():
                    # promise: the aggregation runs now, the projection when the caller shows what it does with the result
                    def finish(act, x=x, graph=graph):
                        if act == "elu":
                            return torch.nn.functional.elu(F_.sage_conv(x, graph, w_l, b_l, w_r))
                        return F_.sage_conv(x, graph, w_l, b_l, w_r, activation=act)
                    needs_grad = torch.is_grad_enabled() and (x.requires_grad or w_l.requires_grad or w_r.requires_grad)
                    return PendingActivation(finish, (x.size(0), self.out_channels), x.dtype, x.device, needs_grad)
                return F_.sage_conv(x, graph, w_l, b_l, w_r, activation=self.activation)
            return _apply_activation(F_.sage_conv(x, graph, w_l, b_l, w_r), self.activation)
        agg = F_.sage_mean_aggregate(x, graph, ew)
        if self.root_weight:
            out = F_.linear(agg, self.lin_l.weight, self.lin_l.bias, x, self.lin_r.weight)
        else:
            out = F_.linear(agg, self.lin_l.weight, self.lin_l.bias)
        if self.normalize:
            out = torch.nn.functional.normalize(out, p=2.0, dim=-1)
        return _apply_activation(out, self.activation)

    def extra_repr(self) -> str:
        return f"{self.in_channels}, {self.out_channels}, aggr=mean"


class GCNConv(MessagePassing):
    """Kipf & Welling layer (reference: gnn/gcn.py:46-48, gnn/gat-gcn.py:58 and their predict twins):
    ``out = D^-1/2 (A + I) D^-1/2 (x W^T) + b`` with ``D = in-degree + 1``.  K4 projection, then ONE neighbourhood
    sum over rows pre-scaled by ``d^-1/2`` (the self loop rides along as the sum's base row) and a row scaling --
    no ``[E, F]`` message tensor, no per-edge norm vector.  ``edge_index`` must not contain self loops already
    (the reference's bond lists never do: adjacency ``nonzero`` of a zero-diagonal matrix, gnn/gcn.py:30-38)."""

    def __init__(self, in_channels: int, out_channels: int, improved: bool = False, cached: bool = False,
                 add_self_loops: bool = True, normalize: bool = True, bias: bool = True, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.cached = improved, cached
        self.add_self_loops, self.normalize = add_self_loops, normalize
        self.lin = Linear(in_channels, out_channels, bias=False, weight_initializer="glorot")
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)

    def reset_parameters(self) -> None:
        self.lin.reset_parameters()
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor, edge_weight: Optional[torch.Tensor] = None):
        require_cuda(x, "GCNConv input x")
        x = F_.real(x)
        graph = graph_index(edge_index, x.size(0))
        if self._edge_weight(graph.num_edges) is not None:
            raise NotImplementedError("explainer edge masks are implemented for GATConv / SAGEConv only")
        xw = F_.linear(x, self.lin.weight, None)
        if not self.normalize:
            out = F_.sum_aggregate(xw, graph, edge_weight, False)
        else:
            fill = (2.0 if self.improved else 1.0) if self.add_self_loops else 0.0
            if edge_weight is None:
                deg = (graph.rowptr[1:] - graph.rowptr[:-1]).to(torch.float32) + fill
            else:
                deg = torch.zeros(x.size(0), dtype=torch.float32, device=x.device)
                deg = deg.index_add_(0, edge_index[1], edge_weight.to(torch.float32)) + fill
            dinv = deg.pow(-0.5)
            dinv = torch.where(torch.isinf(dinv), torch.zeros_like(dinv), dinv).unsqueeze(1)
            z = xw * dinv
            if fill == 1.0:
                s = F_.sum_aggregate(z, graph, edge_weight, True)
            else:
                s = F_.sum_aggregate(z, graph, edge_weight, False)
                if fill != 0.0:
                    s = s + fill * z
            out = s * dinv
        return out if self.bias is None else out + self.bias

    def extra_repr(self) -> str:
        return f"{self.in_channels}, {self.out_channels}"


class GINConv(MessagePassing):
    """Graph isomorphism layer (reference: gnn/gin.py:64-77): ``out = nn((1 + eps) x_i + sum_{j->i} x_j)``; one
    neighbourhood-sum pass (with ``eps = 0`` the ``x_i`` term is the sum's base row), then the user's ``nn``."""

    def __init__(self, nn: torch.nn.Module, eps: float = 0.0, train_eps: bool = False, **kwargs):
        super().__init__()
        self.nn = nn
        self.initial_eps = float(eps)
        if train_eps:
            self.eps = torch.nn.Parameter(torch.empty(1))
        else:
            self.register_buffer("eps", torch.empty(1))
        self.train_eps = train_eps
        with torch.no_grad():
            self.eps.fill_(self.initial_eps)

    def reset_parameters(self) -> None:
        for m in self.nn.modules():
            if m is not self.nn and hasattr(m, "reset_parameters"):
                m.reset_parameters()
        with torch.no_grad():
            self.eps.fill_(self.initial_eps)

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor, size=None) -> torch.Tensor:
        require_cuda(x, "GINConv input x")
        x = F_.real(x)
        graph = graph_index(edge_index, x.size(0))
        if self._edge_weight(graph.num_edges) is not None:
            raise NotImplementedError("explainer edge masks are implemented for GATConv / SAGEConv only")
        if not self.train_eps and self.initial_eps == 0.0:
            h = F_.sum_aggregate(x, graph, None, True)
        else:
            h = F_.sum_aggregate(x, graph, None, False) + (1.0 + self.eps) * x
        return self.nn(h)

    def extra_repr(self) -> str:
        return f"nn={self.nn}"


class GATConv(MessagePassing):
    """Graph attention layer (reference: ablation/model1.py:57,68; gnn/gat.py:54-55; Appendix A.1).
    K4 projection -> K2 scores / edge softmax / aggregation, one autograd node for the message part."""

    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, concat: bool = True,
                 negative_slope: float = 0.2, dropout: float = 0.0, add_self_loops: bool = True,
                 edge_dim: Optional[int] = None, fill_value="mean", bias: bool = True,
                 activation: Optional[str] = None, **kwargs):
        super().__init__()
        if activation not in (None, "relu", "elu"):
            raise ValueError("activation must be None, 'relu' or 'elu'")
        #: extension of PyG's signature (see SAGEConv.activation): fused into the aggregation kernel's epilogue
        #: (`self.relu(self.conv1(x, edge_index))` model1.py:68-69, `F.elu(self.gcn1(x, edge_index))` gnn/gat.py:63)
        self.activation = activation
        if edge_dim is not None:
            raise NotImplementedError("edge features are not used by the reference (no edge_attr)")
        if not add_self_loops:
            raise NotImplementedError("add_self_loops=False is not used by the reference")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.negative_slope, self.dropout = concat, negative_slope, dropout
        self.add_self_loops = add_self_loops
        self.lin = Linear(in_channels, heads * out_channels, bias=False, weight_initializer="glorot")
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(heads * out_channels if concat else out_channels))
        else:
            self.register_parameter("bias", None)
        #: test hook: keep-mask [E', H] in PyG edge order (non-self-loop edges, then one self loop per
        #: node), already scaled by 1/(1-p); replaces the random attention dropout when set.
        self._injected_alpha_mask: Optional[torch.Tensor] = None
        self.reset_parameters()

    def reset_parameters(self) -> None:
        self.lin.reset_parameters()
        for p in (self.att_src, self.att_dst):
            stdv = math.sqrt(6.0 / (p.size(-2) + p.size(-1)))
            nn.init.uniform_(p, -stdv, stdv)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    # -- slot-order helpers (see include/mgs.h) -----------------------------------------------
    @staticmethod
    def _edge_slots(edge_index: torch.Tensor, graph):
        """slot of every original edge, slot of every node's self loop, and the non-self-loop mask."""
        E, N = graph.num_edges, graph.num_nodes
        dev = edge_index.device
        perm = graph.perm.long()
        slot_of_edge = torch.empty(E, dtype=torch.long, device=dev)
        slot_of_edge[perm] = torch.arange(E, device=dev) + edge_index[1][perm]
        self_slot = graph.rowptr[1:].long() + torch.arange(N, device=dev)
        keep = edge_index[0] != edge_index[1]
        return slot_of_edge, self_slot, keep

    def _mask_to_slot_order(self, mask: torch.Tensor, edge_index: torch.Tensor, graph) -> torch.Tensor:
        slot_of_edge, self_slot, keep = self._edge_slots(edge_index, graph)
        out = torch.ones(graph.num_slots, self.heads, dtype=torch.float32, device=edge_index.device)
        n_real = int(keep.sum())
        out[slot_of_edge[keep]] = mask[:n_real].to(out)
        out[self_slot] = mask[n_real:].to(out)
        return out

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor, edge_attr=None, size=None,
                return_attention_weights=None):
        require_cuda(x, "GATConv input x")
        x = F_.real(x)
        H, C, N = self.heads, self.out_channels, x.size(0)
        graph = graph_index(edge_index, N)
        # K <= 36 (the reference's GATConv(35, 35, heads=10)): projection and scores in one pass over x
        fused_scores = (x.dim() == 2 and x.dtype == torch.float32
                        and F_.gat_project_applicable(self.in_channels, H, C))
        if fused_scores:
            xh, a_src, a_dst = F_.gat_project(x, self.lin.weight, self.att_src, self.att_dst, H, C)
        else:
            xh = self.lin(x)
        alpha_mask = None
        if self._injected_alpha_mask is not None:
            alpha_mask = self._mask_to_slot_order(self._injected_alpha_mask, edge_index, graph)
        elif self.training and self.dropout > 0.0:
            keep_p = 1.0 - self.dropout
            alpha_mask = torch.empty(graph.num_slots, H, dtype=torch.float32, device=x.device)
            alpha_mask.bernoulli_(keep_p).div_(keep_p)
        fused_bias = self.bias if (self.concat and self.bias is not None) else None
        ew = self._edge_weight(graph.num_edges)
        # the activation can ride on the aggregation kernel when that kernel writes the layer's final output
        can_fuse = self.concat and not return_attention_weights and H <= 32 and F_.stream_width_ok(H * C)

        def message(act):
            if fused_scores:
                return F_.gat_message(xh, a_src, a_dst, fused_bias, graph, H, C, self.negative_slope,
                                      alpha_mask, ew, scores=True, activation=act)[0]
            return F_.gat_message(xh, self.att_src, self.att_dst, fused_bias, graph, H, C,
                                  self.negative_slope, alpha_mask, ew, activation=act)[0]

        if can_fuse and self.activation is not None:
            return message(self.activation)
        if can_fuse and activation_fusion_enabled():
            # promise: projection and edge softmax inputs are computed, the aggregation kernel is launched when the
            # caller shows what it does with the result (m_gat_graphsage_b200.lazy)
            needs_grad = torch.is_grad_enabled() and (xh.requires_grad or self.att_src.requires_grad
                                                      or (self.bias is not None and self.bias.requires_grad))
            return PendingActivation(message, (N, H * C), xh.dtype, xh.device, needs_grad)
        if fused_scores:
            out, alpha = F_.gat_message(xh, a_src, a_dst, fused_bias, graph, H, C, self.negative_slope,
                                        alpha_mask, ew, scores=True)
        else:
            out, alpha = F_.gat_message(xh, self.att_src, self.att_dst, fused_bias, graph, H, C,
                                        self.negative_slope, alpha_mask, ew)
        if not self.concat:
            out = out.reshape(N, H, C).mean(dim=1)
            if self.bias is not None:
                out = out + self.bias
        out = _apply_activation(out, self.activation)
        if return_attention_weights:
            slot_of_edge, self_slot, keep = self._edge_slots(edge_index, graph)
            loops = torch.arange(N, device=x.device).unsqueeze(0).repeat(2, 1)
            ei = torch.cat([edge_index[:, keep], loops], dim=1)
            a = alpha if alpha_mask is None else alpha * alpha_mask
            return out, (ei, torch.cat([a[slot_of_edge[keep]], a[self_slot]], dim=0))
        return out

    def extra_repr(self) -> str:
        return f"{self.in_channels}, {self.out_channels}, heads={self.heads}"


# ------------------------------------------------------------------------------------------------
# pools
# ------------------------------------------------------------------------------------------------
#: Readouts call ``gmp(x, batch)`` and ``gap(x, batch)`` on the same tensor (ablation/model1.py:72) or only ``gmp``
#: (train.py:119).  The first of the two calls computes BOTH statistics in one pass (the second one is free: x is
#: in registers anyway) and parks the ``[B, 2F]`` result here; the other call, if it comes, returns the other half
#: of the same autograd node.  Set to False to get one kernel per call.
FUSE_MAX_MEAN_POOL = True
_pool_cache = None   # (weakref(x), x._version, weakref(batch), num_graphs, grad_mode, weakref(combined))


class _PooledHalves(torch.autograd.Function):
    """``both[B, 2F] -> (both[:, :F], both[:, F:])`` as ONE autograd node.  Plain slicing gives two SliceBackward nodes:
    each allocates and zero-fills a ``[B, 2F]`` gradient, copies its half in, and autograd adds the two (5 launches).  The
    reference's readout is ``torch.cat([gmp(x, batch), gap(x, batch)], dim=1)`` (ablation/model1.py:72): the cat's backward
    hands back two ADJACENT views of one buffer, which IS the gradient of ``both`` -- no launch at all; anything else
    falls back to one ``torch.cat``."""

    @staticmethod
    def forward(ctx, both, num_feat):
        ctx.F = int(num_feat)
        ctx.shape = tuple(both.shape)
        return both[:, :ctx.F], both[:, ctx.F:]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_max, g_mean):
        B, F = ctx.shape[0], ctx.F
        if (g_max is not None and g_mean is not None and g_max.stride() == g_mean.stride() and g_max.stride(1) == 1
                and g_max.stride(0) >= 2 * F and g_max.untyped_storage().data_ptr() == g_mean.untyped_storage().data_ptr()
                and g_mean.storage_offset() == g_max.storage_offset() + F):
            return g_max.as_strided((B, 2 * F), g_max.stride(), g_max.storage_offset()), None
        if g_max is None and g_mean is None:
            return None, None
        ref = g_max if g_max is not None else g_mean
        zeros = ref.new_zeros(B, F)
        return torch.cat([g_max if g_max is not None else zeros, g_mean if g_mean is not None else zeros], dim=1), None


def _pool_maxmean(x: torch.Tensor, batch: torch.Tensor, gptr: torch.Tensor, num_graphs: int):
    """-> (max half, mean half) of the fused ``[B, 2F]`` result."""
    global _pool_cache
    grad_mode = torch.is_grad_enabled() and x.requires_grad
    c = _pool_cache
    if c is not None:
        halves = c[5]
        if (halves[0]() is not None and halves[1]() is not None and c[0]() is x and c[1] == x._version
                and c[2]() is batch and c[3] == num_graphs and c[4] == grad_mode):
            return halves[0](), halves[1]()
    both = F_.segment_pool_maxmean(x, gptr, num_graphs)
    F = x.size(1)
    if grad_mode:
        mx, mean = _PooledHalves.apply(both, F)
    else:
        mx, mean = both[:, :F], both[:, F:]
    # weak references here; `_pool` parks the half that was NOT asked for on the one it returns, so the pair lives as long
    # as the caller holds either
    _pool_cache = (weakref.ref(x), x._version, weakref.ref(batch), num_graphs, grad_mode, (weakref.ref(mx), weakref.ref(mean)))
    return mx, mean


def _forget_pooled(_ctx=None) -> None:
    """Called when a fused max+mean node runs its backward: its buffers may be freed now, so a later gmp / gap call on
    the same ``x`` must compute afresh instead of returning a slice of the spent autograd node."""
    global _pool_cache
    _pool_cache = None


F_.PoolMaxMeanFn.on_backward = staticmethod(_forget_pooled)


def _pool(x: torch.Tensor, batch: Optional[torch.Tensor], size: Optional[int], mode: str) -> torch.Tensor:
    require_cuda(x, f"global_{mode}_pool input x")
    x = F_.real(x)
    if batch is None:
        batch = torch.zeros(x.size(0), dtype=torch.long, device=x.device)
        size = 1
    if batch.numel() != x.size(0):
        raise ValueError(f"global_{mode}_pool: x has {x.size(0)} rows but batch has {batch.numel()} entries")
    num_graphs = resolve_num_graphs(batch, size)
    gptr = graph_ptr(batch, num_graphs)
    squeeze = x.dim() == 1
    if squeeze:
        x = x.unsqueeze(-1)
    if FUSE_MAX_MEAN_POOL and mode in ("max", "mean") and x.dim() == 2 and x.dtype == torch.float32:
        mx, mean = _pool_maxmean(x, batch, gptr, num_graphs)
        out, other = (mx, mean) if mode == "max" else (mean, mx)
        if getattr(other, "_mgs_pool_sibling", None) is not out:      # (no reference cycle when the second call comes)
            out._mgs_pool_sibling = other
    else:
        out = F_.segment_pool(x, gptr, num_graphs, mode)
    return out.squeeze(-1) if squeeze else out


def global_max_pool(x, batch, size: Optional[int] = None):
    """Per-molecule column-wise max (train.py:119, ablation/model1.py:72); empty molecule -> 0."""
    return _pool(x, batch, size, "max")


def global_mean_pool(x, batch, size: Optional[int] = None):
    """Per-molecule mean (ablation/model1.py:72 ``gap``)."""
    return _pool(x, batch, size, "mean")


def global_add_pool(x, batch, size: Optional[int] = None):
    return _pool(x, batch, size, "add")
