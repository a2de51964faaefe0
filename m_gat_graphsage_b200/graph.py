"""Host side of K0: device-resident sorted CSR / CSC of a batched molecule graph.

``graph_index(edge_index, num_nodes)`` runs ``mgs_csr_build`` once per ``edge_index`` tensor and
caches the result *on the tensor object* (invalidated by ``Tensor._version``), so the two conv
layers of a model, their backward passes and the explainer's 100 epochs all reuse one build.  The
reference path (PyG) instead redoes COO scatter bookkeeping inside every operator call.

``graph_ptr(batch, num_graphs)`` does the same for the molecule segment pointers used by the
pooling kernels; ``num_graphs`` comes from the hint recorded at collation time
(``data._tag_num_graphs``) so no ``batch.max().item()`` host sync is needed on the hot path.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import _lib

_DEBUG = os.environ.get("MGS_DEBUG", "0") not in ("", "0")
_CSR_ATTR = "_mgs_graph"
_PTR_ATTR = "_mgs_gptr"
_NG_ATTR = "_mgs_num_graphs"


# ------------------------------------------------------------------------------------------------
# Device status words without a hot-path sync: every K0 / segment-pointer build copies its status word to pinned host
# memory behind the kernels (non-blocking) and records an event; the words whose event has completed are examined at
# the NEXT build (one batch later), at `check_pending_status()` and at interpreter exit.  A malformed `edge_index` /
# `batch` therefore raises one step late instead of never (out-of-range ids are clamped inside the kernels, so the
# step in between computes on a well-defined, if wrong, graph and cannot fault).
# ------------------------------------------------------------------------------------------------
_pending = []            # [(event, pinned int32[1], kind)]
_MAX_PENDING = 64


def _raise_for_status(word: int, kind: str) -> None:
    if kind == "csr" and word & 1:
        raise IndexError("edge_index contains node ids outside [0, num_nodes) (detected by mgs_csr_build; reported "
                         "asynchronously, possibly one batch after the offending one)")
    if kind == "gptr" and word & 2:
        raise ValueError("batch vector is not sorted ascending (required by the segmented pooling kernels; reported "
                         "asynchronously, possibly one batch after the offending one)")
    if kind == "gptr" and word & 4:
        raise IndexError("batch vector contains graph ids outside [0, num_graphs) (reported asynchronously)")


def _watch_status(status: torch.Tensor, kind: str) -> None:
    if torch.cuda.is_current_stream_capturing():
        return                                       # no host reads of a captured region's memory
    host = torch.empty(1, dtype=torch.int32, device="cpu", pin_memory=True)   # (the launcher sets a CUDA default device)
    host.copy_(status, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    _pending.append((ev, host, kind))
    check_pending_status(block=len(_pending) > _MAX_PENDING)


def check_pending_status(block: bool = False) -> None:
    """Examine the status words of finished builds (``block=True``: wait for all of them); raises on a bad one."""
    global _pending
    keep, err = [], None
    for ev, host, kind in _pending:
        if block:
            ev.synchronize()
        if ev.query():
            try:
                _raise_for_status(int(host[0]), kind)
            except (IndexError, ValueError) as e:
                err = err or e
        else:
            keep.append((ev, host, kind))
    _pending = keep
    if err is not None:
        raise err


def _check_at_exit() -> None:      # pragma: no cover
    try:
        check_pending_status(block=True)
    except (IndexError, ValueError) as e:
        import sys
        print(f"m_gat_graphsage_b200: {e}", file=sys.stderr)
    except Exception:
        pass


import atexit  # noqa: E402

atexit.register(_check_at_exit)


def stream_ptr() -> int:
    """Raw ``cudaStream_t`` of PyTorch's current stream on the current device (no Stream object is built: this is
    called once per C-ABI call, ~50 times per training step)."""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


class _NoGuard:
    __slots__ = ()

    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def device_guard(device):
    """``torch.cuda.device(device)`` only when ``device`` is not already current (the usual one-GPU-per-process case
    pays two attribute reads instead of a context manager with four driver calls)."""
    idx = device.index
    if idx is None or idx == torch._C._cuda_getDevice():
        return _NO_GUARD
    return torch.cuda.device(idx)


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        dev = getattr(t, "device", type(t).__name__)
        raise RuntimeError(
            f"{what}: expected a CUDA tensor, got device={dev}. m_gat_graphsage_b200 runs its operators "
            "in sm_100a CUDA kernels only; there is no CPU fallback.")


class GraphIndex:
    """Output of K0 (all int32, on the device of ``edge_index``)."""
    __slots__ = ("num_nodes", "num_edges", "rowptr", "col", "perm", "colptr", "row", "permt",
                 "csc_pos", "status", "_dst_sorted")

    def check(self) -> None:
        """Host-synchronising validity check (debug / tests only)."""
        s = int(self.status.item())
        if s & 1:
            raise IndexError("edge_index contains node ids outside [0, num_nodes)")

    @property
    def num_slots(self) -> int:
        return self.num_edges + self.num_nodes


def build_graph_index(edge_index: torch.Tensor, num_nodes: int) -> GraphIndex:
    require_cuda(edge_index, "edge_index")
    if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError(f"edge_index must be int64 [2, E], got {edge_index.dtype} {tuple(edge_index.shape)}")
    if edge_index.size(1) > 0 and edge_index.stride(1) != 1:
        edge_index = edge_index.contiguous()
    lib = _lib.load()
    E, N = int(edge_index.size(1)), int(num_nodes)
    dev = edge_index.device
    gi = GraphIndex()
    gi.num_nodes, gi.num_edges = N, E
    i32 = dict(dtype=torch.int32, device=dev)
    gi.rowptr = torch.empty(N + 1, **i32)
    gi.colptr = torch.empty(N + 1, **i32)
    # one allocation for the five E-sized arrays
    flat = torch.empty(5 * E, **i32)
    gi.col, gi.perm, gi.row, gi.permt, gi.csc_pos = (flat[k * E:(k + 1) * E] for k in range(5))
    gi.status = torch.zeros(1, **i32)
    gi._dst_sorted = None
    ws_bytes = int(lib.mgs_csr_workspace_bytes(N, E))
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    with device_guard(dev):
        rc = lib.mgs_csr_build(edge_index.data_ptr(), edge_index.stride(0) if E > 0 else 0, E, N,
                               gi.rowptr.data_ptr(), gi.col.data_ptr(), gi.perm.data_ptr(),
                               gi.colptr.data_ptr(), gi.row.data_ptr(), gi.permt.data_ptr(),
                               gi.csc_pos.data_ptr(), gi.status.data_ptr(), ws.data_ptr(), ws.numel(),
                               stream_ptr())
    _lib.check(rc, "mgs_csr_build")
    if _DEBUG:
        gi.check()
    else:
        _watch_status(gi.status, "csr")
    return gi


def graph_index(edge_index: torch.Tensor, num_nodes: int) -> GraphIndex:
    """Cached K0 build for this ``edge_index`` tensor object."""
    cached = getattr(edge_index, _CSR_ATTR, None)
    if cached is not None:
        version, n, gi = cached
        if version == edge_index._version and n == num_nodes:
            return gi
    gi = build_graph_index(edge_index, num_nodes)
    try:
        setattr(edge_index, _CSR_ATTR, (edge_index._version, num_nodes, gi))
    except (AttributeError, RuntimeError):  # pragma: no cover - exotic tensor subclasses
        pass
    return gi


def resolve_num_graphs(batch: torch.Tensor, size: Optional[int]) -> int:
    if size is not None:
        return int(size)
    hinted = getattr(batch, _NG_ATTR, None)
    if hinted is not None:
        return int(hinted)
    # Same as PyG: one host sync (test.py:185 / gnnexplainer.py:645 hand-made `batch` vectors).  The same
    # round trip also verifies the ascending order the segmented pooling kernels rely on.
    if batch.numel() == 0:
        n = 0
    else:
        mx, unsorted = torch.stack([batch.max(), (batch[1:] < batch[:-1]).any().to(batch.dtype)]).tolist()
        if unsorted:
            raise ValueError("batch vector must be sorted ascending (as Batch collation produces it)")
        n = int(mx) + 1
    try:
        setattr(batch, _NG_ATTR, n)
    except (AttributeError, RuntimeError):  # pragma: no cover
        pass
    return n


def graph_ptr(batch: torch.Tensor, num_graphs: int) -> torch.Tensor:
    """``gptr[B+1]`` int32: first atom of every molecule (cached on the ``batch`` tensor)."""
    require_cuda(batch, "batch")
    if batch.dtype != torch.int64 or batch.dim() != 1:
        raise ValueError(f"batch must be int64 [N], got {batch.dtype} {tuple(batch.shape)}")
    cached = getattr(batch, _PTR_ATTR, None)
    if cached is not None:
        version, b, gptr = cached
        if version == batch._version and b == num_graphs:
            return gptr
    if batch.numel() > 0 and batch.stride(0) != 1:
        batch = batch.contiguous()
    lib = _lib.load()
    gptr = torch.empty(num_graphs + 1, dtype=torch.int32, device=batch.device)
    status = torch.zeros(1, dtype=torch.int32, device=batch.device)
    with device_guard(batch.device):
        rc = lib.mgs_graph_ptr(batch.data_ptr(), batch.numel(), num_graphs, gptr.data_ptr(),
                               status.data_ptr(), stream_ptr())
    _lib.check(rc, "mgs_graph_ptr")
    if _DEBUG:
        s = int(status.item())
        if s & 2:
            raise ValueError("batch vector is not sorted ascending (required by the segmented pooling kernels)")
        if s & 4:
            raise IndexError("batch vector contains graph ids outside [0, num_graphs)")
    else:
        _watch_status(status, "gptr")
    gptr._mgs_num_nodes = int(batch.numel())      # lets the pooling operators check x against the batch vector
    try:
        setattr(batch, _PTR_ATTR, (batch._version, num_graphs, gptr))
    except (AttributeError, RuntimeError):  # pragma: no cover
        pass
    return gptr


def drop_cached_index(*tensors: torch.Tensor) -> None:
    """Forget the CSR / segment pointers cached on ``edge_index`` / ``batch`` tensors (they are keyed on the
    tensor's version counter; CUDA-graph capture must see the builder kernels, so it drops them first)."""
    for t in tensors:
        for a in (_CSR_ATTR, _PTR_ATTR):
            if hasattr(t, a):
                delattr(t, a)
