"""ctypes binding of ``libmgs.so`` (the C ABI of ``include/mgs.h``).

There is no CPU fallback and no alternative backend: if the shared library is missing the first
operator call raises ``MgsLibraryError`` (build it with ``python __graft_entry__.py`` or
``python -m m_gat_graphsage_b200._build``)."""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_double, c_float, c_int32, c_int64, c_size_t, c_uint64, c_void_p
from pathlib import Path
from typing import Optional

LIB_PATH = Path(__file__).resolve().parent / "libmgs.so"


class MgsLibraryError(RuntimeError):
    pass


class MgsError(RuntimeError):
    pass


P = c_void_p
I32, I64, F32, SZ = c_int32, c_int64, c_float, c_size_t
F64 = c_double

# name -> (restype, argtypes); mirrors include/mgs.h one to one
SIGNATURES = {
    "mgs_version": (c_int32, []),
    "mgs_last_error_string": (c_char_p, []),
    "mgs_launch_count": (c_uint64, []),
    "mgs_csr_workspace_bytes": (SZ, [I64, I64]),
    "mgs_csr_build": (I32, [P, I64, I64, I64, P, P, P, P, P, P, P, P, P, SZ, P]),
    "mgs_graph_ptr": (I32, [P, I64, I64, P, P, P]),
    "mgs_wire_expand": (I32, [P, I64, I32, P, I64, P, I64, P, P, I64, P, P]),
    "mgs_sage_aggr_fwd": (I32, [P, I64, I64, I32, P, P, P, P, P, I64, P]),
    "mgs_sage_aggr_bwd": (I32, [P, I64, I64, I32, P, P, P, P, P, P, I64, P]),
    "mgs_sage_aggr_bwd_accumulate": (I32, [P, I64, I64, I32, P, P, P, P, P, P, I64, P, I64, P, I32, P, I64, P]),
    "mgs_sage_aggr_bwd_edge_weight": (I32, [P, I64, P, I64, I64, I32, P, P, P, P, P]),
    "mgs_selftest_div": (I32, [I32, I32, c_uint64, P, P]),
    "mgs_gat_scores_fwd": (I32, [P, I64, I64, I32, I32, P, P, P, P, P]),
    "mgs_gat_alpha_fwd": (I32, [P, P, I64, I32, P, P, F32, P, P]),
    "mgs_gat_aggr_fwd": (I32, [P, I64, I64, I32, I32, P, P, P, P, P, P, P, I64, I32, P, I32, P]),
    "mgs_gat_bwd_edge": (I32, [P, I64, P, I64, I64, I32, I32, P, P, P, P, F32, P, P, P, P, P, P, P, P]),
    "mgs_gat_bwd_node": (I32, [P, I64, I64, I32, I32, P, P, P, P, P, P, P, P, P, P, P, P, I64, P, P]),
    "mgs_gat_bwd_att_workspace_bytes": (SZ, [I32, I32]),
    "mgs_gat_bwd_att": (I32, [P, I64, I64, I32, I32, P, P, P, P, P, SZ, P]),
    "mgs_proj_fwd": (I32, [P, I64, I64, I32, P, I64, I32, P, I64, I32, P, I64, I32, P, P, I64, P, I64, P, I64, P]),
    "mgs_proj_wgrad_workspace_bytes": (SZ, [I32, I32]),
    "mgs_proj_wgrad": (I32, [P, I64, I32, P, I64, I32, P, I64, I32, P, I64, I64, I32, P, I64, P, I64, P, I64, P, SZ, P]),
    "mgs_pool_fwd": (I32, [P, I64, P, I64, I32, I32, P, I64, P]),
    "mgs_pool_bwd": (I32, [P, I64, P, I64, P, I64, P, I64, I32, I32, P, I64, P]),
    "mgs_pool_maxmean_fwd": (I32, [P, I64, P, I64, I32, P, I64, P, P]),
    "mgs_pool_maxmean_bwd": (I32, [P, I64, P, I64, P, I64, P, I64, I32, P, I64, P, I32, P]),
    "mgs_sum_aggr": (I32, [P, I64, I64, I32, P, P, P, P, P, I64, P, I64, P]),
    "mgs_attn_fwd": (I32, [P, I64, P, I64, P, I64, I64, I32, F32, P, P, P, I64, P, P]),
    "mgs_attn_bwd": (I32, [P, I64, P, I64, P, I64, I64, I32, F32, P, P, P, P, P, I64, P, I64, P, I64, P, I64, P]),
    "mgs_linear_fwd_workspace_bytes": (SZ, [I64, I32, I32, I32]),
    "mgs_linear_fwd": (I32, [P, I64, I64, I32, P, I64, I32, P, P, I64, I32, P, I64, P, I64, I32, P, SZ, P]),
    "mgs_linear_dgrad_workspace_bytes": (SZ, [I64, I32, I32]),
    "mgs_linear_dgrad": (I32, [P, I64, I64, I32, P, I64, I32, P, I64, P, SZ, P]),
    "mgs_linear_wgrad_workspace_bytes": (SZ, [I64, I32, I32]),
    "mgs_linear_wgrad": (I32, [P, I64, I64, I32, P, I64, I32, P, I64, P, SZ, P]),
    "mgs_colsum_workspace_bytes": (SZ, [I32]),
    "mgs_colsum": (I32, [P, I64, I64, I32, P, P, SZ, P]),
    "mgs_linear_dgrad2_workspace_bytes": (SZ, [I64, I32, I32, I32]),
    "mgs_linear_dgrad2": (I32, [P, I64, I32, P, I64, P, I64, I32, P, I64, I64, I32, P, I64, P, I32, I32, P, P, SZ, P]),
    "mgs_gat_u_fwd": (I32, [P, I64, P, P, I32, I32, I32, P, P, P]),
    "mgs_gat_u_bwd": (I32, [P, I64, P, P, P, P, I32, I32, I32, P, I64, P, P, P]),
    "mgs_adam_step": (I32, [I32, P, P, P, P, P, F64, F64, F64, F64, F64, I64, P]),
}

_lib: Optional["_Library"] = None


class _Library:
    """Thin proxy over the ctypes handle.  ``profile()`` brackets every C-ABI call with CUDA events on
    the current stream (used by bench.py to attribute step time to kernels); otherwise calls go
    straight through."""

    def __init__(self, cdll: ctypes.CDLL):
        self._cdll = cdll
        self._records = None
        self._call_log = None
        for name in SIGNATURES:
            setattr(self, name, self._make(name, getattr(cdll, name)))

    def _make(self, name, fn):
        cdll_count = self._cdll.mgs_launch_count
        timed = name not in ("mgs_version", "mgs_last_error_string", "mgs_launch_count") and \
            not name.endswith("_workspace_bytes")

        def call(*args):
            if self._call_log is not None and timed:
                n0 = int(cdll_count())
                rc = fn(*args)
                self._call_log.append((name, args, int(cdll_count()) - n0))
                return rc
            if self._records is None or not timed:
                return fn(*args)
            import torch
            start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record()
            rc = fn(*args)
            end.record()
            self._records.append((name, args, start, end))
            return rc

        call.__name__ = name
        return call

    def start_call_log(self) -> None:
        """Record ``(entry point, args, kernels launched)`` for every C-ABI call in order (no events, no syncs): the
        key that joins an ncu launch list of the same program back to C-ABI calls (tools/ncu_join.py)."""
        self._call_log = []

    def stop_call_log(self):
        log, self._call_log = self._call_log or [], None
        return log

    def start_profile(self) -> None:
        self._records = []

    def stop_profile(self):
        """-> list of (name, args, milliseconds); synchronises the device."""
        import torch
        torch.cuda.synchronize()
        recs, self._records = self._records or [], None
        return [(n, a, s.elapsed_time(e)) for n, a, s, e in recs]


def load() -> _Library:
    """Load ``libmgs.so`` once; raise loudly if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise MgsLibraryError(
            f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built. "
            "Run `python __graft_entry__.py` (or `python -m m_gat_graphsage_b200._build`). "
            "There is no CPU or PyTorch fallback for these operators.")
    try:
        from . import _build
        if (_build.CSRC / "common.cu").exists() and not _build.is_current():
            import warnings
            warnings.warn(f"{LIB_PATH} is older than the sources under {_build.CSRC}: rebuild it with "
                          "`python __graft_entry__.py` (the stale binary is being used)", RuntimeWarning, stacklevel=2)
    except Exception:  # pragma: no cover - never let the freshness check break loading
        pass
    try:
        cdll = ctypes.CDLL(str(LIB_PATH))
    except OSError as e:  # pragma: no cover
        raise MgsLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(cdll, name)
        except AttributeError as e:
            raise MgsLibraryError(f"{LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    _lib = _Library(cdll)
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().mgs_last_error_string()
        raise MgsError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load().mgs_launch_count())
