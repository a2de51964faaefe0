"""``torch.autograd.Function`` wrappers around the C ABI (``include/mgs.h``).

PyTorch is plumbing here: it owns device memory, the current stream and the autograd graph; every
forward / backward body below is a handful of ``libmgs.so`` calls on raw pointers.  Each Function
cites the reference operator chain it replaces.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from .graph import GraphIndex, device_guard, require_cuda, stream_ptr
from .lazy import PendingActivation


def real(t):
    """A ``PendingActivation`` (lazy.py) handed to one of our own operators stands for its un-activated result."""
    return t._materialize(None) if isinstance(t, PendingActivation) else t

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


def _check_rows(x: torch.Tensor, gptr: torch.Tensor, what: str) -> None:
    """``gptr`` was built from a batch vector of ``_mgs_num_nodes`` entries (``graph.graph_ptr``): the segmented kernels
    walk ``x`` by those pointers, so a longer batch vector would read (forward) and WRITE (backward) rows past the end
    of ``x``.  PyG's scatter raises on the same mismatch.  Host-side, no sync."""
    n = getattr(gptr, "_mgs_num_nodes", None)
    if n is not None and n != x.size(0):
        raise ValueError(f"{what}: x has {x.size(0)} rows but the batch vector has {n} entries")


ACTIVATIONS = {None: 0, "relu": 1, "elu": 2}


def rows(n: int, f: int, device) -> torch.Tensor:
    """Uninitialised fp32 ``[n, f]`` activation whose rows start 16-byte aligned: a view of an ``[n, roundup(f, 4)]``
    buffer (350 floats = 1400-byte rows are only 8-byte aligned: not addressable by a TMA tensor map, 64-bit vector
    accesses at best).  Every operator takes a leading dimension, so the padding is never read as data."""
    ld = (f + 3) & ~3
    if ld == f:
        return torch.empty(n, f, dtype=torch.float32, device=device)
    return torch.empty(n, ld, dtype=torch.float32, device=device)[:, :f]


def stream_row_words(num_feat: int) -> int:
    """Words per row of the ReLU bit mask exchanged between the aggregation kernels (csrc/stream.cuh: V * iterations for
    16-byte aligned rows, i.e. ``rows()`` buffers); 0 when the width does not fit the streaming kernels."""
    v = 4 if num_feat % 4 == 0 else (2 if num_feat % 2 == 0 else 1)
    need = (num_feat // v + 31) // 32
    for it in (1, 2, 4, 6, 8):
        if need <= it:
            return v * it if v * it <= 32 else 0
    return 0


def _proj_wgrad_on_tensor_cores(n_rows: int, k: int, n_out: int, dxh: torch.Tensor) -> bool:
    """GatProjFn.backward: the weight gradient of the K = 35 projection through mgs_linear_wgrad's TMA kernel (needs
    >= 32 channels on both sides, 16-byte aligned dxh rows, enough atoms to split).  OFF unless MGS_PROJ_WGRAD_TC=1:
    measured on B200 at 130 k atoms it is 0.077 ms (tcgen05, 48-column tiles: the kernel costs ~0.45 us per 16-atom block
    whatever the tile width) + 0.037 ms for the two [N, H] score gradients on the small FFMA kernel + an 18 MB padded copy
    of x, against 0.140 ms for the one FFMA kernel -- the training step did not move (2.65 ms both ways)."""
    import os
    if os.environ.get("MGS_PROJ_WGRAD_TC", "0") != "1" or os.environ.get("MGS_WGRAD_TMA", "1") == "0":
        return False
    return n_rows >= 4096 and 32 <= k <= 48 and n_out >= 32 and _aligned_rows(dxh)


def _aligned_rows(t: torch.Tensor) -> bool:
    return t.data_ptr() % 16 == 0 and (t.size(0) <= 1 or t.stride(0) % 4 == 0)


def _mark_masked(gx: torch.Tensor, relu_out: torch.Tensor) -> None:
    """``gx`` has been multiplied by ``relu_out > 0`` by the kernel that produced it (the backward of the ReLU whose
    output ``relu_out`` is).  The version counter makes the mark void as soon as autograd accumulates another
    consumer's gradient into the same buffer in place."""
    gx._mgs_masked = (relu_out.data_ptr(), gx._version)


def _act_backward(g: torch.Tensor, out: torch.Tensor, act) -> torch.Tensor:
    """Gradient w.r.t. the pre-activation of a fused ``out = act(z)``.  ReLU: nothing to do when the consumer of
    ``out`` already masked the gradient it produced (see ``_mark_masked``); otherwise ATen's own backward kernels."""
    if act == "relu":
        if getattr(g, "_mgs_masked", None) == (out.data_ptr(), g._version):
            return g
        return torch.ops.aten.threshold_backward(g, out, 0)
    if act == "elu":
        return torch.ops.aten.elu_backward(g, 1.0, 1.0, 1.0, True, out)
    return g


STREAM_MAX_CHUNKS = 256      # csrc/common.cuh iters_for(): 8 iterations x 32 lanes of V-float chunks per row


def stream_width_ok(num_feat: int) -> bool:
    """Rows of ``num_feat`` floats fit the block-streamed aggregation kernels (``csrc/stream.cuh``): at most 256 vector
    chunks of V = 4 / 2 / 1 floats (V follows the row width's alignment)."""
    v = 4 if num_feat % 4 == 0 else (2 if num_feat % 2 == 0 else 1)
    return num_feat // v <= STREAM_MAX_CHUNKS


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
    out = rows(M, Nout, x.device)
    ws = _workspace(lib.mgs_linear_fwd_workspace_bytes(M, K, Nout, K2), x.device)
    with device_guard(x.device):
        rc = lib.mgs_linear_fwd(x.data_ptr(), _ld(x), M, K, w.data_ptr(), _ld(w), Nout, _ptr(b),
                                _ptr(x2), _ld(x2) if x2 is not None else 0, K2,
                                _ptr(w2), _ld(w2) if w2 is not None else 0,
                                out.data_ptr(), _ld(out), 1 if relu else 0, ws.data_ptr(), ws.numel(), stream_ptr())
    _lib.check(rc, "mgs_linear_fwd")
    return out


def linear_dgrad_raw(g, w):
    lib = _lib.load()
    M, Nout = g.shape
    K = w.size(1)
    dx = rows(M, K, g.device)
    ws = _workspace(lib.mgs_linear_dgrad_workspace_bytes(M, Nout, K), g.device)
    with device_guard(g.device):
        rc = lib.mgs_linear_dgrad(g.data_ptr(), _ld(g), M, Nout, w.data_ptr(), _ld(w), K, dx.data_ptr(), _ld(dx),
                                  ws.data_ptr(), ws.numel(), stream_ptr())
    _lib.check(rc, "mgs_linear_dgrad")
    return dx


def linear_wgrad_raw(g, x):
    lib = _lib.load()
    M, Nout = g.shape
    K = x.size(1)
    dw = torch.empty(Nout, K, dtype=torch.float32, device=g.device)
    ws = _workspace(lib.mgs_linear_wgrad_workspace_bytes(M, Nout, K), g.device)
    with device_guard(g.device):
        rc = lib.mgs_linear_wgrad(g.data_ptr(), _ld(g), M, Nout, x.data_ptr(), _ld(x), K, dw.data_ptr(), K,
                                  ws.data_ptr(), ws.numel(), stream_ptr())
    _lib.check(rc, "mgs_linear_wgrad")
    return dw


def colsum_raw(g):
    lib = _lib.load()
    M, Nout = g.shape
    out = torch.empty(Nout, dtype=torch.float32, device=g.device)
    ws = _workspace(lib.mgs_colsum_workspace_bytes(Nout), g.device)
    with device_guard(g.device):
        rc = lib.mgs_colsum(g.data_ptr(), _ld(g), M, Nout, out.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr())
    _lib.check(rc, "mgs_colsum")
    return out


class LinearFn(torch.autograd.Function):
    """``y = x W^T (+ x2 W2^T) + b``; the optional second pair is SAGEConv's ``lin_r(x)`` fused into
    the ``lin_l(mean)`` GEMM (one pass over the output instead of two GEMMs and an add)."""

    @staticmethod
    def forward(ctx, x, w, b, x2, w2, relu=False):
        x, w = _mat(x, "x"), _mat(w, "weight")
        b = _vec(b, "bias")
        if x2 is not None:
            x2, w2 = _mat(x2, "x2"), _mat(w2, "weight2")
        out = linear_forward_raw(x, w, b, x2, w2, relu=bool(relu))
        ctx.save_for_backward(x, w, x2, w2, out if relu else None)
        ctx.has_bias, ctx.relu = b is not None, bool(relu)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, w, x2, w2, out = ctx.saved_tensors
        if ctx.relu:
            g = _act_backward(g, out, "relu")
        g = _mat(g, "grad_output")
        need = ctx.needs_input_grad
        dx = linear_dgrad_raw(g, w) if need[0] else None
        dw = linear_wgrad_raw(g, x) if need[1] else None
        db = colsum_raw(g) if (ctx.has_bias and need[2]) else None
        dx2 = linear_dgrad_raw(g, w2) if (x2 is not None and need[3]) else None
        dw2 = linear_wgrad_raw(g, x2) if (x2 is not None and need[4]) else None
        return dx, dw, db, dx2, dw2, None


def linear(x, weight, bias=None, x2=None, weight2=None, activation=None):
    """``activation='relu'``: fused into the GEMM epilogue (the readout MLP applies it right after ``fc_g1``,
    ablation/model1.py:74)."""
    if activation not in (None, "relu"):
        raise ValueError("linear: only activation='relu' can be fused")
    x, x2 = real(x), real(x2)
    lead = x.shape[:-1]
    if x.dim() != 2:
        x = x.reshape(-1, x.size(-1))
        if x2 is not None:
            x2 = x2.reshape(-1, x2.size(-1))
    out = LinearFn.apply(x, weight, bias, x2, weight2, activation == "relu")
    if activation == "relu":
        out._mgs_act = "relu"
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
        with device_guard(x.device):
            rc = lib.mgs_sage_aggr_fwd(x.data_ptr(), _ld(x), N, F, graph.rowptr.data_ptr(), graph.col.data_ptr(),
                                       graph.perm.data_ptr(), _ptr(ew), out.data_ptr(), F, stream_ptr())
        _lib.check(rc, "mgs_sage_aggr_fwd")
        ctx.graph = graph
        ctx.save_for_backward(x if ew is not None else None, ew)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, ew = ctx.saved_tensors
        graph = ctx.graph
        g = _mat(g, "grad_output")
        lib = _lib.load()
        N, F = g.shape
        gx = gew = None
        with device_guard(g.device):
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
    return SageAggrFn.apply(real(x), graph, edge_weight)


class SageConvFn(torch.autograd.Function):
    """Whole SAGEConv (no explainer edge mask): ``out = lin_l(mean_{j->i} x_j) + lin_r(x_i)``.  One autograd node
    instead of aggregate + linear, so that the two data gradients of ``x`` meet inside the aggregation kernel
    (``gx = g W_r`` written by the GEMM, ``+= A^T (g W_l / indeg)`` by ``mgs_sage_aggr_bwd_accumulate``) and
    autograd's ``[N, F]`` add disappears."""

    @staticmethod
    def forward(ctx, x, graph: GraphIndex, w_l, b_l, w_r, relu=False, x_is_relu=False, x_bits=None):
        # relu: ReLU fused into the GEMM epilogue (model1.py:70-71 `self.relu(self.conv2(..))`)
        # x_is_relu: x is the output of a fused ReLU (model1.py:69): this node's backward applies that ReLU's mask to
        #            the gradient it produces, inside the aggregation kernel that writes it
        x, w_l, w_r = _mat(x, "x"), _mat(w_l, "lin_l.weight"), _mat(w_r, "lin_r.weight")
        b_l = _vec(b_l, "lin_l.bias")
        if x.size(0) != graph.num_nodes:
            raise ValueError(f"x has {x.size(0)} rows but the graph has {graph.num_nodes} nodes")
        lib = _lib.load()
        N, F = x.shape
        agg = rows(N, F, x.device)
        # bias gradient for free: when the rows are padded (F % 4 != 0) the first padding column of `agg` is set to 1.0,
        # and the weight-gradient GEMM g^T [agg | 1] of the backward delivers colsum(g) = d b_l as its last column
        # (the projection reads K = F columns: the padding never enters the forward)
        ctx.ones_col = bool(b_l is not None and N > 1 and agg.stride(0) > F and ctx.needs_input_grad[3])
        if ctx.ones_col:
            agg._base[:, F].fill_(1.0)
        with device_guard(x.device):
            rc = lib.mgs_sage_aggr_fwd(x.data_ptr(), _ld(x), N, F, graph.rowptr.data_ptr(), graph.col.data_ptr(),
                                       graph.perm.data_ptr(), 0, agg.data_ptr(), _ld(agg), stream_ptr())
        _lib.check(rc, "mgs_sage_aggr_fwd")
        out = linear_forward_raw(agg, w_l, b_l, x, w_r, relu=bool(relu))
        ctx.graph, ctx.has_bias, ctx.relu, ctx.x_is_relu = graph, b_l is not None, bool(relu), bool(x_is_relu)
        ctx.save_for_backward(x, agg, w_l, w_r, out if relu else None, x_bits if x_is_relu else None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, agg, w_l, w_r, out, x_bits = ctx.saved_tensors
        graph = ctx.graph
        if ctx.relu:
            g = _act_backward(g, out, "relu")
        g = _mat(g, "grad_output")
        need = ctx.needs_input_grad
        lib = _lib.load()
        N, F = x.shape
        gx = None
        if need[0]:
            gx = _sage_dx_fused(ctx, g, x, w_l, w_r, x_bits, graph)
        if need[0] and gx is None:
            # both data gradients in ONE GEMM over g: [g W_r | g W_l] (700 output columns take the 256-wide tiles the
            # tensor-core kernel is efficient with; two 350-column GEMMs are bound by the per-tile activation path)
            both = linear_dgrad_raw(g, torch.cat([w_r, w_l], dim=1))
            dx_r, d_agg = both[:, :F], both[:, F:]
            gx = rows(N, F, g.device)
            # the ReLU mask of x: the bits its producer left (48 bytes per row) when the row layouts agree, else x itself
            use_bits = (ctx.x_is_relu and x_bits is not None and x_bits.size(1) == stream_row_words(F)
                        and _aligned_rows(both) and _aligned_rows(gx))
            with device_guard(g.device):
                rc = lib.mgs_sage_aggr_bwd_accumulate(d_agg.data_ptr(), _ld(d_agg), N, F, graph.rowptr.data_ptr(),
                                                      graph.colptr.data_ptr(), graph.row.data_ptr(),
                                                      graph.permt.data_ptr(), 0, dx_r.data_ptr(), _ld(dx_r),
                                                      x.data_ptr() if (ctx.x_is_relu and not use_bits) else 0, _ld(x),
                                                      x_bits.data_ptr() if use_bits else 0,
                                                      x_bits.size(1) if use_bits else 0,
                                                      gx.data_ptr(), _ld(gx), stream_ptr())
            _lib.check(rc, "mgs_sage_aggr_bwd_accumulate")
            if ctx.x_is_relu:
                _mark_masked(gx, x)
        dw_l = db = None
        if ctx.ones_col and need[2] and need[3]:
            ext = linear_wgrad_raw(g, agg._base[:, :F + 1])          # [out, F + 1]: d W_l | d b_l
            dw_l, db = ext[:, :F], ext[:, F].contiguous()
        else:
            dw_l = linear_wgrad_raw(g, agg) if need[2] else None
            db = colsum_raw(g) if (ctx.has_bias and need[3]) else None
        dw_r = linear_wgrad_raw(g, x) if need[4] else None
        return gx, None, dw_l, db, dw_r, None, None, None


def _sage_dx_fused(ctx, g, x, w_l, w_r, x_bits, graph):
    """SAGEConv data gradient with the aggregation BEFORE the GEMM:  gx = relu'(x) (g W_r + aggT(g) W_l)  where
    aggT(v)[j] = sum_{i: j -> i} v[i] / deg(i)  (the mean aggregation's backward is linear, so it commutes with W_l).
    One K = 2 O GEMM with the ReLU mask in its epilogue (``mgs_linear_dgrad2``) instead of a [N, 2 F] GEMM output that the
    aggregation kernel reads back: 0.50 -> 0.44 ms at 130 k atoms, 350 channels, and 183 MB less traffic.  Returns None when
    the path does not apply (the caller keeps the two-column-block formulation)."""
    import os
    if os.environ.get("MGS_SAGE_DX_FUSED", "1") == "0":
        return None
    N, F = x.shape
    O = g.size(1)
    if O > F or N < 1024 or not (_aligned_rows(g) and w_l.is_contiguous() and w_r.is_contiguous()):
        return None
    bits_v = 4 if F % 4 == 0 else (2 if F % 2 == 0 else 1)
    use_bits = ctx.x_is_relu and x_bits is not None and x_bits.size(1) == stream_row_words(F) and x_bits.size(1) > 0
    if ctx.x_is_relu and not use_bits:
        return None
    lib = _lib.load()
    ghat = rows(N, O, g.device)
    gx = rows(N, F, g.device)
    # x is a fused layer's output: that layer's bias gradient is colsum(gx) -- summed in the GEMM epilogue (5.7 MB of
    # per-32-row partials) instead of a pass over the [N, F] gradient (0.041 ms)
    colsum = torch.empty(F, dtype=torch.float32, device=g.device) if ctx.x_is_relu else None
    nbytes = lib.mgs_linear_dgrad2_workspace_bytes(N, O, O, F)
    ws = _workspace(nbytes, g.device)
    with device_guard(g.device):
        rc = lib.mgs_sage_aggr_bwd(g.data_ptr(), _ld(g), N, O, graph.rowptr.data_ptr(), graph.colptr.data_ptr(),
                                   graph.row.data_ptr(), graph.permt.data_ptr(), 0, ghat.data_ptr(), _ld(ghat), stream_ptr())
        _lib.check(rc, "mgs_sage_aggr_bwd")
        rc = lib.mgs_linear_dgrad2(g.data_ptr(), _ld(g), O, w_r.data_ptr(), _ld(w_r), ghat.data_ptr(), _ld(ghat), O,
                                   w_l.data_ptr(), _ld(w_l), N, F, gx.data_ptr(), _ld(gx),
                                   x_bits.data_ptr() if use_bits else 0, x_bits.size(1) if use_bits else 0, bits_v,
                                   _ptr(colsum), ws.data_ptr(), ws.numel(), stream_ptr())
    if rc == 4:                                   # MGS_ERR_UNSUPPORTED: operands not on the TMA kernel
        return None
    _lib.check(rc, "mgs_linear_dgrad2")
    if ctx.x_is_relu:
        _mark_masked(gx, x)
        gx._mgs_colsum = (colsum, gx._version)
    return gx


def sage_conv(x, graph, w_l, b_l, w_r, activation=None):
    """``activation='relu'``: fused into the projection's epilogue.  When ``x`` itself is the output of a fused ReLU
    (``x._mgs_act``), that ReLU's backward is fused into this node's gradient kernel."""
    if activation not in (None, "relu"):
        raise ValueError("sage_conv: only activation='relu' can be fused")
    x = real(x)
    out = SageConvFn.apply(x, graph, w_l, b_l, w_r, activation == "relu", getattr(x, "_mgs_act", None) == "relu",
                           getattr(x, "_mgs_relu_bits", None))
    if activation == "relu":
        out._mgs_act = "relu"
    return out


# ------------------------------------------------------------------------------------------------
# K2: GATConv message passing (A.1 steps 2-8, concat layout)
# ------------------------------------------------------------------------------------------------
class GatMessageFn(torch.autograd.Function):
    """``out[i] = sum_slots alpha * w_e * xh[j] (+ bias)`` with alpha the per-destination edge softmax of
    ``leaky_relu(a_src[j] + a_dst[i])``.  One autograd node for scores, softmax, dropout mask and
    aggregation so the backward can run in the order that writes ``dxh`` exactly once."""

    @staticmethod
    def forward(ctx, xh, att_src, att_dst, bias, graph: GraphIndex, heads, channels, negative_slope,
                alpha_mask, edge_weight, scores=False, activation=None):
        # scores=True: `att_src` / `att_dst` ARE the scores a_src / a_dst [N, H] (computed from x by GatProjFn);
        # their gradients are da_src / da_dst and no attention-vector terms are formed here
        xh = _mat(xh, "xh")
        H, C = int(heads), int(channels)
        N = xh.size(0)
        if xh.size(1) != H * C:
            raise ValueError(f"xh must be [N, {H * C}]")
        if N != graph.num_nodes:
            raise ValueError(f"xh has {N} rows but the graph has {graph.num_nodes} nodes")
        ctx.att_shape = tuple(att_src.shape)
        ctx.scores = bool(scores)
        if scores and (tuple(att_src.shape) != (N, H) or tuple(att_dst.shape) != (N, H)):
            raise ValueError(f"a_src / a_dst must be [{N}, {H}]")
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
        alpha = torch.empty(S, H, **f32)
        out = rows(N, H * C, dev)
        sp = stream_ptr
        with device_guard(dev):
            if scores:
                a_src, a_dst = att_src.view(N, H), att_dst.view(N, H)
            else:
                a_src = torch.empty(N, H, **f32)
                a_dst = torch.empty(N, H, **f32)
                _lib.check(lib.mgs_gat_scores_fwd(xh.data_ptr(), _ld(xh), N, H, C, att_src.data_ptr(),
                                                  att_dst.data_ptr(), a_src.data_ptr(), a_dst.data_ptr(), sp()),
                           "mgs_gat_scores_fwd")
            _lib.check(lib.mgs_gat_alpha_fwd(a_src.data_ptr(), a_dst.data_ptr(), N, H, graph.rowptr.data_ptr(),
                                             graph.col.data_ptr(), float(negative_slope), alpha.data_ptr(), sp()),
                       "mgs_gat_alpha_fwd")
            alpha_used = alpha if amask is None else alpha * amask
            # ReLU epilogue: out > 0 also leaves as one bit per element for the consumer's fused ReLU backward
            words = stream_row_words(H * C) if (activation == "relu" and ctx.needs_input_grad[0] and _aligned_rows(xh)
                                                and _aligned_rows(out) and (bias is None or bias.data_ptr() % 16 == 0)) else 0
            bits = torch.empty(N, words, dtype=torch.int32, device=dev) if words else None
            _lib.check(lib.mgs_gat_aggr_fwd(xh.data_ptr(), _ld(xh), N, H, C, alpha_used.data_ptr(),
                                            graph.rowptr.data_ptr(), graph.col.data_ptr(), graph.perm.data_ptr(),
                                            _ptr(ew), _ptr(bias), out.data_ptr(), _ld(out), ACTIVATIONS[activation],
                                            _ptr(bits), words, sp()),
                       "mgs_gat_aggr_fwd")

        ctx.graph, ctx.H, ctx.C, ctx.slope = graph, H, C, float(negative_slope)
        ctx.has_bias, ctx.act = bias is not None, activation
        ctx.save_for_backward(xh, att_src, att_dst, a_src, a_dst, alpha, amask, ew, out if activation else None)
        ctx.mark_non_differentiable(alpha)
        if bits is not None:
            ctx.mark_non_differentiable(bits)
        return out, alpha, bits

    @staticmethod
    @once_differentiable
    def backward(ctx, g, _g_alpha, _g_bits=None):
        xh, att_src, att_dst, a_src, a_dst, alpha, amask, ew, out = ctx.saved_tensors
        graph, H, C = ctx.graph, ctx.H, ctx.C
        if ctx.act:
            g = _act_backward(g, out, ctx.act)
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
        dxh = rows(N, H * C, dev)
        want_dew = ew is not None and need[9]
        dew = torch.zeros(graph.num_edges, **f32) if want_dew else None
        sp = stream_ptr
        with device_guard(dev):
            _lib.check(lib.mgs_gat_bwd_edge(g.data_ptr(), _ld(g), xh.data_ptr(), _ld(xh), N, H, C,
                                            alpha.data_ptr(), _ptr(amask), a_src.data_ptr(), a_dst.data_ptr(),
                                            ctx.slope, graph.rowptr.data_ptr(), graph.col.data_ptr(),
                                            graph.perm.data_ptr(), _ptr(ew), dr.data_ptr(), da_dst.data_ptr(),
                                            _ptr(dew), sp()), "mgs_gat_bwd_edge")
            alpha_used = alpha if amask is None else alpha * amask
            _lib.check(lib.mgs_gat_bwd_node(g.data_ptr(), _ld(g), N, H, C, alpha_used.data_ptr(), dr.data_ptr(),
                                            da_dst.data_ptr(), 0 if ctx.scores else att_src.data_ptr(),
                                            0 if ctx.scores else att_dst.data_ptr(),
                                            graph.rowptr.data_ptr(), graph.colptr.data_ptr(), graph.row.data_ptr(),
                                            graph.csc_pos.data_ptr(), graph.permt.data_ptr(), _ptr(ew),
                                            dxh.data_ptr(), _ld(dxh), da_src.data_ptr(), sp()), "mgs_gat_bwd_node")
            datt_src = datt_dst = None
            if ctx.scores:
                datt_src, datt_dst = da_src, da_dst
            elif need[1] or need[2]:
                datt_src = torch.empty(H * C, **f32)
                datt_dst = torch.empty(H * C, **f32)
                ws = _workspace(lib.mgs_gat_bwd_att_workspace_bytes(H, C), dev)
                _lib.check(lib.mgs_gat_bwd_att(xh.data_ptr(), _ld(xh), N, H, C, da_src.data_ptr(),
                                               da_dst.data_ptr(), datt_src.data_ptr(), datt_dst.data_ptr(),
                                               ws.data_ptr(), ws.numel(), sp()), "mgs_gat_bwd_att")
        dbias = None
        if ctx.has_bias and need[3]:
            # the kernel that produced g may have summed its columns on the way (mgs_linear_dgrad2's epilogue)
            tag = getattr(g, "_mgs_colsum", None)
            dbias = tag[0] if (tag is not None and tag[1] == g._version and tag[0].numel() == g.size(1)) else colsum_raw(g)
        if datt_src is not None:
            datt_src, datt_dst = datt_src.view(ctx.att_shape), datt_dst.view(ctx.att_shape)
        return (dxh if need[0] else None, datt_src if need[1] else None, datt_dst if need[2] else None,
                dbias, None, None, None, None, None, dew, None, None)


def gat_message(xh, att_src, att_dst, bias, graph, heads, channels, negative_slope=0.2,
                alpha_mask=None, edge_weight=None, scores=False, activation=None):
    """Returns ``(out [N, H*C], alpha [(E+N), H] in slot order)``.  ``scores=True``: the two ``att`` arguments
    are the per-node scores ``a_src`` / ``a_dst`` ``[N, H]`` themselves (see ``gat_project``).  ``activation``
    (``'relu'`` / ``'elu'``): applied to ``out`` in the aggregation kernel's epilogue."""
    if activation not in ACTIVATIONS:
        raise ValueError(f"gat_message: unknown activation {activation!r}")
    xh = real(xh)
    out, alpha, bits = GatMessageFn.apply(xh, att_src, att_dst, bias, graph, heads, channels, negative_slope,
                                          alpha_mask, edge_weight, scores, activation)
    if activation is not None:
        out._mgs_act = activation
        out._mgs_relu_bits = bits          # `out > 0`, one bit per element (None unless ReLU + training)
    return out, alpha


# ------------------------------------------------------------------------------------------------
# K4 small-K: GATConv input projection fused with the attention scores
# ------------------------------------------------------------------------------------------------
PROJ_MAX_K_FWD, PROJ_MAX_K_WGRAD, PROJ_MAX_OUT = 64, 36, 384


def gat_project_applicable(in_channels: int, heads: int, channels: int) -> bool:
    return in_channels <= PROJ_MAX_K_WGRAD and heads * channels + 2 * heads <= PROJ_MAX_OUT


class GatProjFn(torch.autograd.Function):
    """``xh = x W^T``, ``a_src = x U_src^T``, ``a_dst = x U_dst^T`` in one pass over ``x`` (``mgs_proj_fwd``);
    backward ``[dW; dU_src; dU_dst] = [dxh | da_src | da_dst]^T x`` in one pass over the gradients."""

    @staticmethod
    def forward(ctx, x, w, u_src, u_dst):
        x, w, u_src, u_dst = _mat(x, "x"), _mat(w, "weight"), _mat(u_src, "u_src"), _mat(u_dst, "u_dst")
        lib = _lib.load()
        N, K = x.shape
        n0, H = w.size(0), u_src.size(0)
        f32 = dict(dtype=torch.float32, device=x.device)
        xh, a_src, a_dst = rows(N, n0, x.device), torch.empty(N, H, **f32), torch.empty(N, H, **f32)
        with device_guard(x.device):
            rc = lib.mgs_proj_fwd(x.data_ptr(), _ld(x), N, K, w.data_ptr(), _ld(w), n0, u_src.data_ptr(), _ld(u_src), H,
                                  u_dst.data_ptr(), _ld(u_dst), H, 0, xh.data_ptr(), _ld(xh), a_src.data_ptr(), H,
                                  a_dst.data_ptr(), H, stream_ptr())
        _lib.check(rc, "mgs_proj_fwd")
        ctx.save_for_backward(x, w, u_src, u_dst)
        return xh, a_src, a_dst

    @staticmethod
    @once_differentiable
    def backward(ctx, dxh, da_src, da_dst):
        x, w, u_src, u_dst = ctx.saved_tensors
        N, K = x.shape
        n0, H = w.size(0), u_src.size(0)
        f32 = dict(dtype=torch.float32, device=x.device)
        dxh = _mat(dxh, "dxh") if dxh is not None else torch.zeros(N, n0, **f32)
        da_src = _mat(da_src, "da_src") if da_src is not None else torch.zeros(N, H, **f32)
        da_dst = _mat(da_dst, "da_dst") if da_dst is not None else torch.zeros(N, H, **f32)
        lib = _lib.load()
        need = ctx.needs_input_grad
        dw = du_src = du_dst = dx = None
        if (need[1] or need[2] or need[3]) and _proj_wgrad_on_tensor_cores(N, K, n0, dxh):
            # d W = dxh^T x on the TMA-fed tcgen05 weight-gradient kernel (csrc/tc_wgrad.cuh): dxh goes through tensor memory,
            # x is the 48-column shared-memory operand (two 32-channel boxes) -- the FFMA kernel is FP32-issue bound at
            # 0.14 ms.  x needs 16-byte aligned rows for its tensor map: a padded copy (18 MB) when the caller's x has
            # 140-byte rows.
            xp = x
            if not _aligned_rows(x):
                xp = rows(N, K, x.device)
                xp.copy_(x)
            dw = linear_wgrad_raw(dxh, xp)
            du_src, du_dst = torch.empty(H, K, **f32), torch.empty(H, K, **f32)
            ws = _workspace(lib.mgs_proj_wgrad_workspace_bytes(K, 2 * H), x.device)
            with device_guard(x.device):      # the two [N, H] score gradients stay on the FFMA kernel
                rc = lib.mgs_proj_wgrad(da_src.data_ptr(), _ld(da_src), H, da_dst.data_ptr(), _ld(da_dst), H, 0, 0, 0,
                                        x.data_ptr(), _ld(x), N, K, du_src.data_ptr(), K, du_dst.data_ptr(), K, 0, 0,
                                        ws.data_ptr(), ws.numel(), stream_ptr())
            _lib.check(rc, "mgs_proj_wgrad")
        elif need[1] or need[2] or need[3]:
            dw, du_src, du_dst = torch.empty(n0, K, **f32), torch.empty(H, K, **f32), torch.empty(H, K, **f32)
            ws = _workspace(lib.mgs_proj_wgrad_workspace_bytes(K, n0 + 2 * H), x.device)
            with device_guard(x.device):
                rc = lib.mgs_proj_wgrad(dxh.data_ptr(), _ld(dxh), n0, da_src.data_ptr(), _ld(da_src), H,
                                        da_dst.data_ptr(), _ld(da_dst), H, x.data_ptr(), _ld(x), N, K,
                                        dw.data_ptr(), K, du_src.data_ptr(), K, du_dst.data_ptr(), K,
                                        ws.data_ptr(), ws.numel(), stream_ptr())
            _lib.check(rc, "mgs_proj_wgrad")
        if need[0]:      # atom importance / explainer node mask: d x = dxh W + da_src U_src + da_dst U_dst
            dx = linear_dgrad_raw(dxh, w)
            dx += linear_dgrad_raw(da_src, u_src)
            dx += linear_dgrad_raw(da_dst, u_dst)
        return dx, dw, du_src, du_dst


class GatScoreWeightsFn(torch.autograd.Function):
    """``U_src[h, :] = sum_c att_src[h, c] W[hC + c, :]`` (and ``U_dst``), ``[H, K]`` each: one launch forward, one
    backward (``mgs_gat_u_fwd / _bwd``) -- as a PyTorch expression (mul, sum and their autograd nodes, twice) it was ~16
    launches of 3-5 us per training step.  Autograd adds the returned ``dW`` to the one of ``GatProjFn``."""

    @staticmethod
    def forward(ctx, weight, att_src, att_dst, heads, channels):
        w = _mat(weight, "lin.weight")
        H, C, K = int(heads), int(channels), w.size(1)
        a_s, a_d = _vec(att_src.reshape(-1), "att_src"), _vec(att_dst.reshape(-1), "att_dst")
        if w.size(0) != H * C or a_s.numel() != H * C or a_d.numel() != H * C:
            raise ValueError(f"score weights: weight must be [{H * C}, K] and att_src / att_dst hold {H * C} values")
        lib = _lib.load()
        u_src = torch.empty(H, K, dtype=torch.float32, device=w.device)
        u_dst = torch.empty(H, K, dtype=torch.float32, device=w.device)
        with device_guard(w.device):
            rc = lib.mgs_gat_u_fwd(w.data_ptr(), _ld(w), a_s.data_ptr(), a_d.data_ptr(), H, C, K, u_src.data_ptr(),
                                   u_dst.data_ptr(), stream_ptr())
        _lib.check(rc, "mgs_gat_u_fwd")
        ctx.save_for_backward(w, a_s, a_d)
        ctx.dims = (H, C, K, tuple(att_src.shape), tuple(att_dst.shape))
        return u_src, u_dst

    @staticmethod
    @once_differentiable
    def backward(ctx, du_src, du_dst):
        w, a_s, a_d = ctx.saved_tensors
        H, C, K, shape_s, shape_d = ctx.dims
        f32 = dict(dtype=torch.float32, device=w.device)
        du_src = _mat(du_src, "dU_src") if du_src is not None else torch.zeros(H, K, **f32)
        du_dst = _mat(du_dst, "dU_dst") if du_dst is not None else torch.zeros(H, K, **f32)
        dw, da_s, da_d = torch.empty(H * C, K, **f32), torch.empty(H * C, **f32), torch.empty(H * C, **f32)
        lib = _lib.load()
        with device_guard(w.device):
            rc = lib.mgs_gat_u_bwd(w.data_ptr(), _ld(w), a_s.data_ptr(), a_d.data_ptr(), du_src.data_ptr(),
                                   du_dst.data_ptr(), H, C, K, dw.data_ptr(), K, da_s.data_ptr(), da_d.data_ptr(),
                                   stream_ptr())
        _lib.check(rc, "mgs_gat_u_bwd")
        return dw, da_s.view(shape_s), da_d.view(shape_d), None, None


def gat_project(x, weight, att_src, att_dst, heads: int, channels: int):
    """-> ``(xh [N, H*C], a_src [N, H], a_dst [N, H])``.  ``U[h, :] = sum_c att[h, c] W[hC + c, :]`` (``[H, K]``) is its
    own autograd node, so ``dU`` travels on to ``weight`` and the attention vectors."""
    x = real(x)
    u_src, u_dst = GatScoreWeightsFn.apply(weight, att_src, att_dst, heads, channels)
    return GatProjFn.apply(x, weight, u_src, u_dst)


# ------------------------------------------------------------------------------------------------
# K3: segmented global pooling (A.3)
# ------------------------------------------------------------------------------------------------
class PoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gptr, num_graphs, mode):
        x = _mat(x, "x")
        _check_rows(x, gptr, "global pool")
        lib = _lib.load()
        N, F = x.shape
        B = int(num_graphs)
        out = torch.empty(B, F, dtype=torch.float32, device=x.device)
        with device_guard(x.device):
            rc = lib.mgs_pool_fwd(x.data_ptr(), _ld(x), gptr.data_ptr(), B, F, mode, out.data_ptr(), F, stream_ptr())
        _lib.check(rc, "mgs_pool_fwd")
        ctx.mode, ctx.B, ctx.N = mode, B, N
        if mode == POOL_MODES["max"]:
            ctx.save_for_backward(gptr, x, out)
        else:
            ctx.save_for_backward(gptr)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        saved = ctx.saved_tensors
        gptr = saved[0]
        x = saved[1] if ctx.mode == 0 else None
        out = saved[2] if ctx.mode == 0 else None
        g = _mat(g, "grad_output")
        lib = _lib.load()
        F = g.size(1)
        gx = torch.empty(ctx.N, F, dtype=torch.float32, device=g.device)
        with device_guard(g.device):
            rc = lib.mgs_pool_bwd(g.data_ptr(), _ld(g), _ptr(x), _ld(x) if x is not None else 0,
                                  _ptr(out), F if out is not None else 0, gptr.data_ptr(), ctx.B, F, ctx.mode,
                                  gx.data_ptr(), F, stream_ptr())
        _lib.check(rc, "mgs_pool_bwd")
        return gx, None, None, None


def segment_pool(x, gptr, num_graphs, mode: str):
    return PoolFn.apply(real(x), gptr, num_graphs, POOL_MODES[mode])


class PoolMaxMeanFn(torch.autograd.Function):
    """``[global_max_pool(x) | global_mean_pool(x)]`` as one ``[B, 2F]`` tensor: one pass over ``x`` forward,
    one backward node writing ``gx`` once (instead of two nodes and an autograd ``add`` of two ``[N, F]``
    gradients)."""

    on_backward = staticmethod(lambda ctx: None)      # nn.py drops its one-entry result cache here

    @staticmethod
    def forward(ctx, x, gptr, num_graphs, x_is_relu=False):
        x = _mat(x, "x")
        _check_rows(x, gptr, "global max / mean pool")
        ctx.x_is_relu = bool(x_is_relu)
        lib = _lib.load()
        N, F = x.shape
        B = int(num_graphs)
        out = torch.empty(B, 2 * F, dtype=torch.float32, device=x.device)
        # tie counts of the max (needed by its gradient) come out of the same pass when a backward will follow
        ties = torch.empty(B, F, dtype=torch.float32, device=x.device) if x.requires_grad else None
        with device_guard(x.device):
            rc = lib.mgs_pool_maxmean_fwd(x.data_ptr(), _ld(x), gptr.data_ptr(), B, F, out.data_ptr(), 2 * F,
                                          _ptr(ties), stream_ptr())
        _lib.check(rc, "mgs_pool_maxmean_fwd")
        ctx.B, ctx.N, ctx.F = B, N, F
        ctx.save_for_backward(gptr, x, out, ties)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        gptr, x, out, ties = ctx.saved_tensors
        PoolMaxMeanFn.on_backward(ctx)
        g = _mat(g, "grad_output")
        lib = _lib.load()
        gx = rows(ctx.N, ctx.F, g.device)
        with device_guard(g.device):
            rc = lib.mgs_pool_maxmean_bwd(g.data_ptr(), _ld(g), x.data_ptr(), _ld(x), out.data_ptr(), 2 * ctx.F,
                                          gptr.data_ptr(), ctx.B, ctx.F, gx.data_ptr(), _ld(gx), _ptr(ties),
                                          1 if ctx.x_is_relu else 0, stream_ptr())
        _lib.check(rc, "mgs_pool_maxmean_bwd")
        if ctx.x_is_relu:
            _mark_masked(gx, x)
        return gx, None, None, None


def segment_pool_maxmean(x, gptr, num_graphs):
    """When ``x`` is the output of a fused ReLU (``x._mgs_act``, ablation/model1.py:71-72), that ReLU's backward rides
    on the pooling backward, which reads ``x`` anyway."""
    x = real(x)
    return PoolMaxMeanFn.apply(x, gptr, num_graphs, getattr(x, "_mgs_act", None) == "relu")


# ------------------------------------------------------------------------------------------------
# K5: streaming all-pairs attention of ModifiedGATLayer (train.py:87-99)
# ------------------------------------------------------------------------------------------------
class StreamAttnFn(torch.autograd.Function):
    """``softmax(K_new Q^T / sqrt(d)) V + V`` (train.py:96-98) on one packed projection ``y = [Q | K_new | V]``
    (``[N, 3d]``), without the ``[N, N]`` score / weight matrices the reference materialises (and keeps for the
    backward).  ``seg`` / ``gptr`` restrict every atom to its own molecule (``None``: the whole batch)."""

    @staticmethod
    def forward(ctx, y, d, scale, seg, gptr):
        y = _mat(y, "y")
        N, d = y.size(0), int(d)
        if y.size(1) != 3 * d:
            raise ValueError(f"y: expected [N, {3 * d}] = [Q | K_new | V], got {tuple(y.shape)}")
        if seg is not None and seg.numel() != N:
            raise ValueError(f"molecule attention: y has {N} rows but the batch vector has {seg.numel()} entries")
        if gptr is not None:
            _check_rows(y, gptr, "molecule attention")
        lib = _lib.load()
        ld, base = _ld(y), y.data_ptr()
        out = torch.empty(N, d, dtype=torch.float32, device=y.device)
        lse = torch.empty(N, dtype=torch.float32, device=y.device)
        with device_guard(y.device):
            rc = lib.mgs_attn_fwd(base + 4 * d, ld, base, ld, base + 8 * d, ld, N, d, float(scale), _ptr(seg),
                                  _ptr(gptr), out.data_ptr(), d, lse.data_ptr(), stream_ptr())
        _lib.check(rc, "mgs_attn_fwd")
        ctx.d, ctx.scale = d, float(scale)
        ctx.save_for_backward(y, out, lse, seg, gptr)
        return out + y[:, 2 * d:]

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        y, out, lse, seg, gptr = ctx.saved_tensors
        g = _mat(g, "grad_output")
        d, N = ctx.d, y.size(0)
        lib = _lib.load()
        delta = (g * out).sum(1)
        dy = torch.empty(N, 3 * d, dtype=torch.float32, device=y.device)
        ld, base, dbase = _ld(y), y.data_ptr(), dy.data_ptr()
        with device_guard(y.device):
            rc = lib.mgs_attn_bwd(base + 4 * d, ld, base, ld, base + 8 * d, ld, N, d, ctx.scale, _ptr(seg), _ptr(gptr),
                                  lse.data_ptr(), delta.data_ptr(), g.data_ptr(), _ld(g),
                                  dbase + 4 * d, 3 * d, dbase, 3 * d, dbase + 8 * d, 3 * d, stream_ptr())
        _lib.check(rc, "mgs_attn_bwd")
        dy[:, 2 * d:] += g                       # the residual `+ V`
        return dy, None, None, None, None


def stream_attention(y, d: int, scale: float, seg=None, gptr=None):
    return StreamAttnFn.apply(real(y), d, scale, seg, gptr)


# ------------------------------------------------------------------------------------------------
# neighbourhood SUM (GCNConv / GINConv: gnn/gcn.py:46-48, gnn/gat-gcn.py:58, gnn/gin.py:64-77)
# ------------------------------------------------------------------------------------------------
def _sum_aggr_raw(src, dst, ptr, idx, eid, ew, add_self) -> None:
    """``dst = [src] + gather-sum(src)``; rows wider than the streaming kernels take (256 vector chunks) are processed
    in column blocks of the same launch."""
    lib = _lib.load()
    N, F = src.shape
    step = F if stream_width_ok(F) else STREAM_MAX_CHUNKS
    with device_guard(src.device):
        for c0 in range(0, F, step):
            w = min(step, F - c0)
            sp, dp = src.data_ptr() + 4 * c0, dst.data_ptr() + 4 * c0
            rc = lib.mgs_sum_aggr(sp, _ld(src), N, w, ptr.data_ptr(), idx.data_ptr(), eid.data_ptr(), _ptr(ew),
                                  sp if add_self else 0, _ld(src), dp, _ld(dst), stream_ptr())
            _lib.check(rc, "mgs_sum_aggr")


class SumAggrFn(torch.autograd.Function):
    """``out_i = [x_i] + sum_{j->i} w_e x_j`` (``add_self``: the bracket).  PyG: index_select -> (* edge_weight) ->
    scatter_add; the backward is the same sum over the transposed (CSC) index.  ``edge_weight`` gets no gradient."""

    @staticmethod
    def forward(ctx, x, graph: GraphIndex, edge_weight, add_self: bool):
        x = _mat(x, "x")
        ew = _vec(edge_weight, "edge_weight")
        if x.size(0) != graph.num_nodes:
            raise ValueError(f"x has {x.size(0)} rows but the graph has {graph.num_nodes} nodes")
        if ew is not None and ew.numel() != graph.num_edges:
            raise ValueError(f"edge_weight has {ew.numel()} entries but the graph has {graph.num_edges} edges")
        N, F = x.shape
        out = torch.empty(N, F, dtype=torch.float32, device=x.device)
        _sum_aggr_raw(x, out, graph.rowptr, graph.col, graph.perm, ew, add_self)
        ctx.graph, ctx.add_self, ctx.shape = graph, bool(add_self), (N, F)
        ctx.save_for_backward(ew)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (ew,) = ctx.saved_tensors
        if ctx.needs_input_grad[2]:
            raise NotImplementedError("no gradient with respect to edge_weight of the neighbourhood sum")
        if not ctx.needs_input_grad[0]:
            return None, None, None, None
        g = _mat(g, "grad_output")
        graph = ctx.graph
        N, F = ctx.shape
        gx = torch.empty(N, F, dtype=torch.float32, device=g.device)
        _sum_aggr_raw(g, gx, graph.colptr, graph.row, graph.permt, ew, ctx.add_self)
        return gx, None, None, None


def sum_aggregate(x, graph, edge_weight=None, add_self: bool = False):
    return SumAggrFn.apply(real(x), graph, edge_weight, add_self)
