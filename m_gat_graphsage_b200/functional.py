"""``torch.autograd.Function`` wrappers around the C ABI (``include/mgs.h``).

PyTorch is plumbing here: it owns device memory, the current stream and the autograd graph; every
forward / backward body below is a handful of ``libmgs.so`` calls on raw pointers.  Each Function
cites the reference operator chain it replaces.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .graph import GraphIndex, require_cuda, stream_ptr

POOL_MODES = {"max": 0, "mean": 1, "add": 2, "sum": 2}


def _mat(t: torch.Tensor, what: str) -> torch.Tensor:
    """2-D fp32 CUDA matrix with unit inner stride (leading dimension = stride(0))."""
    require_cuda(t, what)
    if t.dtype != torch.float32:
        raise TypeError(f"{what}: expected float32, got {t.dtype} (the hot path is fp32 like the reference)")
    if t.dim() != 2:
        raise ValueError(f"{what}: expected a 2-D tensor, got shape {tuple(t.shape)}")
    if t.size(0) > 1 and (t.stride(1) != 1 or t.stride(0) < t.size(1)):
        t = t.contiguous()
    elif t.size(0) <= 1 and t.size(1) > 1 and t.stride(1) != 1:
        t = t.contiguous()
    return t


def _ld(t: torch.Tensor) -> int:
    return int(t.stride(0)) if t.size(0) > 1 else max(int(t.size(1)), 1)


def _vec(t: Optional[torch.Tensor], what: str) -> Optional[torch.Tensor]:
    if t is None:
        return None
    require_cuda(t, what)
    if t.dtype != torch.float32:
        raise TypeError(f"{what}: expected float32, got {t.dtype}")
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------------
# K4: dense projection  y = x W^T (+ x2 W2^T) + b        (ATen addmm in the reference path)
# ------------------------------------------------------------------------------------------------
def linear_forward_raw(x, w, b=None, x2=None, w2=None, relu=False):
    lib = _lib.load()
    M, K = x.shape
    Nout = w.size(0)
    K2 = x2.size(1) if x2 is not None else 0
    out = torch.empty(M, Nout, dtype=torch.float32, device=x.device)
    ws = _workspace(lib.mgs_linear_fwd_workspace_bytes(M, K, Nout, K2), x.device)
    with torch.cuda.device(x.device):
        rc = lib.mgs_linear_fwd(x.data_ptr(), _ld(x), M, K, w.data_ptr(), _ld(w), Nout, _ptr(b),
                                _ptr(x2), _ld(x2) if x2 is not None else 0, K2,
                                _ptr(w2), _ld(w2) if w2 is not None else 0,
                                out.data_ptr(), Nout, 1 if relu else 0, ws.data_ptr(), ws.numel(), stream_ptr())
    _lib.check(rc, "mgs_linear_fwd")
    return out


def linear_dgrad_raw(g, w):
    lib = _lib.load()
    M, Nout = g.shape
    K = w.size(1)
    dx = torch.empty(M, K, dtype=torch.float32, device=g.device)
    ws = _workspace(lib.mgs_linear_dgrad_workspace_bytes(M, Nout, K), g.device)
    with torch.cuda.device(g.device):
        rc = lib.mgs_linear_dgrad(g.data_ptr(), _ld(g), M, Nout, w.data_ptr(), _ld(w), K, dx.data_ptr(), K,
                                  ws.data_ptr(), ws.numel(), stream_ptr())
    _lib.check(rc, "mgs_linear_dgrad")
    return dx


def linear_wgrad_raw(g, x):
    lib = _lib.load()
    M, Nout = g.shape
    K = x.size(1)
    dw = torch.empty(Nout, K, dtype=torch.float32, device=g.device)
    ws = _workspace(lib.mgs_linear_wgrad_workspace_bytes(M, Nout, K), g.device)
    with torch.cuda.device(g.device):
        rc = lib.mgs_linear_wgrad(g.data_ptr(), _ld(g), M, Nout, x.data_ptr(), _ld(x), K, dw.data_ptr(), K,
                                  ws.data_ptr(), ws.numel(), stream_ptr())
    _lib.check(rc, "mgs_linear_wgrad")
    return dw


def colsum_raw(g):
    lib = _lib.load()
    M, Nout = g.shape
    out = torch.empty(Nout, dtype=torch.float32, device=g.device)
    ws = _workspace(lib.mgs_colsum_workspace_bytes(Nout), g.device)
    with torch.cuda.device(g.device):
        rc = lib.mgs_colsum(g.data_ptr(), _ld(g), M, Nout, out.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr())
    _lib.check(rc, "mgs_colsum")
    return out


class LinearFn(torch.autograd.Function):
    """``y = x W^T (+ x2 W2^T) + b``; the optional second pair is SAGEConv's ``lin_r(x)`` fused into
    the ``lin_l(mean)`` GEMM (one pass over the output instead of two GEMMs and an add)."""

    @staticmethod
    def forward(ctx, x, w, b, x2, w2):
        x, w = _mat(x, "x"), _mat(w, "weight")
        b = _vec(b, "bias")
        if x2 is not None:
            x2, w2 = _mat(x2, "x2"), _mat(w2, "weight2")
        ctx.save_for_backward(x, w, x2, w2)
        ctx.has_bias = b is not None
        return linear_forward_raw(x, w, b, x2, w2)

    @staticmethod
    def backward(ctx, g):
        x, w, x2, w2 = ctx.saved_tensors
        g = _mat(g, "grad_output")
        need = ctx.needs_input_grad
        dx = linear_dgrad_raw(g, w) if need[0] else None
        dw = linear_wgrad_raw(g, x) if need[1] else None
        db = colsum_raw(g) if (ctx.has_bias and need[2]) else None
        dx2 = linear_dgrad_raw(g, w2) if (x2 is not None and need[3]) else None
        dw2 = linear_wgrad_raw(g, x2) if (x2 is not None and need[4]) else None
        return dx, dw, db, dx2, dw2


def linear(x, weight, bias=None, x2=None, weight2=None):
    lead = x.shape[:-1]
    if x.dim() != 2:
        x = x.reshape(-1, x.size(-1))
        if x2 is not None:
            x2 = x2.reshape(-1, x2.size(-1))
    out = LinearFn.apply(x, weight, bias, x2, weight2)
    return out if len(lead) == 1 else out.reshape(*lead, out.size(-1))


# ------------------------------------------------------------------------------------------------
# K1: SAGEConv mean aggregation     (index_select -> scatter_add -> count -> divide, A.2)
# ------------------------------------------------------------------------------------------------
class SageAggrFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, graph: GraphIndex, edge_weight):
        x = _mat(x, "x")
        ew = _vec(edge_weight, "edge_weight")
        if x.size(0) != graph.num_nodes:
            raise ValueError(f"x has {x.size(0)} rows but the graph has {graph.num_nodes} nodes")
        if ew is not None and ew.numel() != graph.num_edges:
            raise ValueError("edge_weight must have one entry per edge")
        lib = _lib.load()
        N, F = x.shape
        out = torch.empty(N, F, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.mgs_sage_aggr_fwd(x.data_ptr(), _ld(x), N, F, graph.rowptr.data_ptr(), graph.col.data_ptr(),
                                       graph.perm.data_ptr(), _ptr(ew), out.data_ptr(), F, stream_ptr())
        _lib.check(rc, "mgs_sage_aggr_fwd")
        ctx.graph = graph
        ctx.save_for_backward(x if ew is not None else None, ew)
        return out

    @staticmethod
    def backward(ctx, g):
        x, ew = ctx.saved_tensors
        graph = ctx.graph
        g = _mat(g, "grad_output")
        lib = _lib.load()
        N, F = g.shape
        gx = gew = None
        with torch.cuda.device(g.device):
            if ctx.needs_input_grad[0]:
                gx = torch.empty(N, F, dtype=torch.float32, device=g.device)
                rc = lib.mgs_sage_aggr_bwd(g.data_ptr(), _ld(g), N, F, graph.rowptr.data_ptr(),
                                           graph.colptr.data_ptr(), graph.row.data_ptr(), graph.permt.data_ptr(),
                                           _ptr(ew), gx.data_ptr(), F, stream_ptr())
                _lib.check(rc, "mgs_sage_aggr_bwd")
            if ew is not None and ctx.needs_input_grad[2]:
                gew = torch.zeros(graph.num_edges, dtype=torch.float32, device=g.device)
                rc = lib.mgs_sage_aggr_bwd_edge_weight(g.data_ptr(), _ld(g), x.data_ptr(), _ld(x), N, F,
                                                       graph.rowptr.data_ptr(), graph.col.data_ptr(),
                                                       graph.perm.data_ptr(), gew.data_ptr(), stream_ptr())
                _lib.check(rc, "mgs_sage_aggr_bwd_edge_weight")
        return gx, None, gew


def sage_mean_aggregate(x, graph, edge_weight=None):
    return SageAggrFn.apply(x, graph, edge_weight)


# ------------------------------------------------------------------------------------------------
# K2: GATConv message passing (A.1 steps 2-8, concat layout)
# ------------------------------------------------------------------------------------------------
class GatMessageFn(torch.autograd.Function):
    """``out[i] = sum_slots alpha * w_e * xh[j] (+ bias)`` with alpha the per-destination edge softmax of
    ``leaky_relu(a_src[j] + a_dst[i])``.  One autograd node for scores, softmax, dropout mask and
    aggregation so the backward can run in the order that writes ``dxh`` exactly once."""

    @staticmethod
    def forward(ctx, xh, att_src, att_dst, bias, graph: GraphIndex, heads, channels, negative_slope,
                alpha_mask, edge_weight):
        xh = _mat(xh, "xh")
        H, C = int(heads), int(channels)
        N = xh.size(0)
        if xh.size(1) != H * C:
            raise ValueError(f"xh must be [N, {H * C}]")
        if N != graph.num_nodes:
            raise ValueError(f"xh has {N} rows but the graph has {graph.num_nodes} nodes")
        ctx.att_shape = tuple(att_src.shape)
        att_src = _vec(att_src.reshape(-1), "att_src")
        att_dst = _vec(att_dst.reshape(-1), "att_dst")
        bias = _vec(bias, "bias")
        amask = _vec(alpha_mask, "alpha_mask")
        ew = _vec(edge_weight, "edge_weight")
        S = graph.num_slots
        if amask is not None and tuple(amask.shape) != (S, H):
            raise ValueError(f"alpha_mask must be [{S}, {H}] in slot order")
        if ew is not None and ew.numel() != graph.num_edges:
            raise ValueError("edge_weight must have one entry per edge")
        lib = _lib.load()
        dev = xh.device
        f32 = dict(dtype=torch.float32, device=dev)
        a_src = torch.empty(N, H, **f32)
        a_dst = torch.empty(N, H, **f32)
        alpha = torch.empty(S, H, **f32)
        out = torch.empty(N, H * C, **f32)
        sp = stream_ptr
        with torch.cuda.device(dev):
            _lib.check(lib.mgs_gat_scores_fwd(xh.data_ptr(), _ld(xh), N, H, C, att_src.data_ptr(),
                                              att_dst.data_ptr(), a_src.data_ptr(), a_dst.data_ptr(), sp()),
                       "mgs_gat_scores_fwd")
            _lib.check(lib.mgs_gat_alpha_fwd(a_src.data_ptr(), a_dst.data_ptr(), N, H, graph.rowptr.data_ptr(),
                                             graph.col.data_ptr(), float(negative_slope), alpha.data_ptr(), sp()),
                       "mgs_gat_alpha_fwd")
            alpha_used = alpha if amask is None else alpha * amask
            _lib.check(lib.mgs_gat_aggr_fwd(xh.data_ptr(), _ld(xh), N, H, C, alpha_used.data_ptr(),
                                            graph.rowptr.data_ptr(), graph.col.data_ptr(), graph.perm.data_ptr(),
                                            _ptr(ew), _ptr(bias), out.data_ptr(), H * C, sp()),
                       "mgs_gat_aggr_fwd")
        ctx.graph, ctx.H, ctx.C, ctx.slope = graph, H, C, float(negative_slope)
        ctx.has_bias = bias is not None
        ctx.save_for_backward(xh, att_src, att_dst, a_src, a_dst, alpha, amask, ew)
        ctx.mark_non_differentiable(alpha)
        return out, alpha

    @staticmethod
    def backward(ctx, g, _g_alpha):
        xh, att_src, att_dst, a_src, a_dst, alpha, amask, ew = ctx.saved_tensors
        graph, H, C = ctx.graph, ctx.H, ctx.C
        g = _mat(g, "grad_output")
        lib = _lib.load()
        dev = g.device
        N = g.size(0)
        S = graph.num_slots
        f32 = dict(dtype=torch.float32, device=dev)
        need = ctx.needs_input_grad
        dr = torch.empty(S, H, **f32)
        da_dst = torch.empty(N, H, **f32)
        da_src = torch.empty(N, H, **f32)
        dxh = torch.empty(N, H * C, **f32)
        want_dew = ew is not None and need[9]
        dew = torch.zeros(graph.num_edges, **f32) if want_dew else None
        sp = stream_ptr
        with torch.cuda.device(dev):
            _lib.check(lib.mgs_gat_bwd_edge(g.data_ptr(), _ld(g), xh.data_ptr(), _ld(xh), N, H, C,
                                            alpha.data_ptr(), _ptr(amask), a_src.data_ptr(), a_dst.data_ptr(),
                                            ctx.slope, graph.rowptr.data_ptr(), graph.col.data_ptr(),
                                            graph.perm.data_ptr(), _ptr(ew), dr.data_ptr(), da_dst.data_ptr(),
                                            _ptr(dew), sp()), "mgs_gat_bwd_edge")
            alpha_used = alpha if amask is None else alpha * amask
            _lib.check(lib.mgs_gat_bwd_node(g.data_ptr(), _ld(g), N, H, C, alpha_used.data_ptr(), dr.data_ptr(),
                                            da_dst.data_ptr(), att_src.data_ptr(), att_dst.data_ptr(),
                                            graph.rowptr.data_ptr(), graph.colptr.data_ptr(), graph.row.data_ptr(),
                                            graph.csc_pos.data_ptr(), graph.permt.data_ptr(), _ptr(ew),
                                            dxh.data_ptr(), H * C, da_src.data_ptr(), sp()), "mgs_gat_bwd_node")
            datt_src = datt_dst = None
            if need[1] or need[2]:
                datt_src = torch.empty(H * C, **f32)
                datt_dst = torch.empty(H * C, **f32)
                ws = _workspace(lib.mgs_gat_bwd_att_workspace_bytes(H, C), dev)
                _lib.check(lib.mgs_gat_bwd_att(xh.data_ptr(), _ld(xh), N, H, C, da_src.data_ptr(),
                                               da_dst.data_ptr(), datt_src.data_ptr(), datt_dst.data_ptr(),
                                               ws.data_ptr(), ws.numel(), sp()), "mgs_gat_bwd_att")
        dbias = colsum_raw(g) if (ctx.has_bias and need[3]) else None
        if datt_src is not None:
            datt_src, datt_dst = datt_src.view(ctx.att_shape), datt_dst.view(ctx.att_shape)
        return (dxh if need[0] else None, datt_src if need[1] else None, datt_dst if need[2] else None,
                dbias, None, None, None, None, None, dew)


def gat_message(xh, att_src, att_dst, bias, graph, heads, channels, negative_slope=0.2,
                alpha_mask=None, edge_weight=None):
    """Returns ``(out [N, H*C], alpha [(E+N), H] in slot order)``."""
    out, alpha = GatMessageFn.apply(xh, att_src, att_dst, bias, graph, heads, channels, negative_slope,
                                    alpha_mask, edge_weight)
    return out, alpha


# ------------------------------------------------------------------------------------------------
# K3: segmented global pooling (A.3)
# ------------------------------------------------------------------------------------------------
class PoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gptr, num_graphs, mode):
        x = _mat(x, "x")
        lib = _lib.load()
        N, F = x.shape
        B = int(num_graphs)
        out = torch.empty(B, F, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.mgs_pool_fwd(x.data_ptr(), _ld(x), gptr.data_ptr(), B, F, mode, out.data_ptr(), F, stream_ptr())
        _lib.check(rc, "mgs_pool_fwd")
        ctx.mode, ctx.B, ctx.N = mode, B, N
        if mode == POOL_MODES["max"]:
            ctx.save_for_backward(gptr, x, out)
        else:
            ctx.save_for_backward(gptr)
        return out

    @staticmethod
    def backward(ctx, g):
        saved = ctx.saved_tensors
        gptr = saved[0]
        x = saved[1] if ctx.mode == 0 else None
        out = saved[2] if ctx.mode == 0 else None
        g = _mat(g, "grad_output")
        lib = _lib.load()
        F = g.size(1)
        gx = torch.empty(ctx.N, F, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            rc = lib.mgs_pool_bwd(g.data_ptr(), _ld(g), _ptr(x), _ld(x) if x is not None else 0,
                                  _ptr(out), F if out is not None else 0, gptr.data_ptr(), ctx.B, F, ctx.mode,
                                  gx.data_ptr(), F, stream_ptr())
        _lib.check(rc, "mgs_pool_bwd")
        return gx, None, None, None


def segment_pool(x, gptr, num_graphs, mode: str):
    return PoolFn.apply(x, gptr, num_graphs, POOL_MODES[mode])


class PoolMaxMeanFn(torch.autograd.Function):
    """``[global_max_pool(x) | global_mean_pool(x)]`` as one ``[B, 2F]`` tensor: one pass over ``x`` forward,
    one backward node writing ``gx`` once (instead of two nodes and an autograd ``add`` of two ``[N, F]``
    gradients)."""

    @staticmethod
    def forward(ctx, x, gptr, num_graphs):
        x = _mat(x, "x")
        lib = _lib.load()
        N, F = x.shape
        B = int(num_graphs)
        out = torch.empty(B, 2 * F, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.mgs_pool_maxmean_fwd(x.data_ptr(), _ld(x), gptr.data_ptr(), B, F, out.data_ptr(), 2 * F,
                                          stream_ptr())
        _lib.check(rc, "mgs_pool_maxmean_fwd")
        ctx.B, ctx.N, ctx.F = B, N, F
        ctx.save_for_backward(gptr, x, out)
        return out

    @staticmethod
    def backward(ctx, g):
        gptr, x, out = ctx.saved_tensors
        g = _mat(g, "grad_output")
        lib = _lib.load()
        gx = torch.empty(ctx.N, ctx.F, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            rc = lib.mgs_pool_maxmean_bwd(g.data_ptr(), _ld(g), x.data_ptr(), _ld(x), out.data_ptr(), 2 * ctx.F,
                                          gptr.data_ptr(), ctx.B, ctx.F, gx.data_ptr(), ctx.F, stream_ptr())
        _lib.check(rc, "mgs_pool_maxmean_bwd")
        return gx, None, None


def segment_pool_maxmean(x, gptr, num_graphs):
    return PoolMaxMeanFn.apply(x, gptr, num_graphs)
