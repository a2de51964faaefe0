"""ctypes binding of ``libmgs.so`` (the C ABI of ``include/mgs.h``).

There is no CPU fallback and no alternative backend: if the shared library is missing the first
operator call raises ``MgsLibraryError`` (build it with ``python __graft_entry__.py`` or
``python -m m_gat_graphsage_b200._build``)."""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_float, c_int32, c_int64, c_size_t, c_uint64, c_void_p
from pathlib import Path
from typing import Optional

LIB_PATH = Path(__file__).resolve().parent / "libmgs.so"


class MgsLibraryError(RuntimeError):
    pass


class MgsError(RuntimeError):
    pass


P = c_void_p
I32, I64, F32, SZ = c_int32, c_int64, c_float, c_size_t

# name -> (restype, argtypes); mirrors include/mgs.h one to one
SIGNATURES = {
    "mgs_version": (c_int32, []),
    "mgs_last_error_string": (c_char_p, []),
    "mgs_launch_count": (c_uint64, []),
    "mgs_csr_workspace_bytes": (SZ, [I64, I64]),
    "mgs_csr_build": (I32, [P, I64, I64, I64, P, P, P, P, P, P, P, P, P, SZ, P]),
    "mgs_graph_ptr": (I32, [P, I64, I64, P, P, P]),
    "mgs_sage_aggr_fwd": (I32, [P, I64, I64, I32, P, P, P, P, P, I64, P]),
    "mgs_sage_aggr_bwd": (I32, [P, I64, I64, I32, P, P, P, P, P, P, I64, P]),
    "mgs_sage_aggr_bwd_edge_weight": (I32, [P, I64, P, I64, I64, I32, P, P, P, P, P]),
    "mgs_gat_scores_fwd": (I32, [P, I64, I64, I32, I32, P, P, P, P, P]),
    "mgs_gat_alpha_fwd": (I32, [P, P, I64, I32, P, P, F32, P, P]),
    "mgs_gat_aggr_fwd": (I32, [P, I64, I64, I32, I32, P, P, P, P, P, P, P, I64, P]),
    "mgs_gat_bwd_edge": (I32, [P, I64, P, I64, I64, I32, I32, P, P, P, P, F32, P, P, P, P, P, P, P, P]),
    "mgs_gat_bwd_node": (I32, [P, I64, I64, I32, I32, P, P, P, P, P, P, P, P, P, P, P, P, I64, P, P]),
    "mgs_gat_bwd_att_workspace_bytes": (SZ, [I32, I32]),
    "mgs_gat_bwd_att": (I32, [P, I64, I64, I32, I32, P, P, P, P, P, SZ, P]),
    "mgs_pool_fwd": (I32, [P, I64, P, I64, I32, I32, P, I64, P]),
    "mgs_pool_bwd": (I32, [P, I64, P, I64, P, I64, P, I64, I32, I32, P, I64, P]),
    "mgs_linear_fwd": (I32, [P, I64, I64, I32, P, I64, I32, P, P, I64, I32, P, I64, P, I64, I32, P]),
    "mgs_linear_dgrad": (I32, [P, I64, I64, I32, P, I64, I32, P, I64, P]),
    "mgs_linear_wgrad_workspace_bytes": (SZ, [I64, I32, I32]),
    "mgs_linear_wgrad": (I32, [P, I64, I64, I32, P, I64, I32, P, I64, P, SZ, P]),
    "mgs_colsum_workspace_bytes": (SZ, [I32]),
    "mgs_colsum": (I32, [P, I64, I64, I32, P, P, SZ, P]),
}

_lib: Optional[ctypes.CDLL] = None


def load() -> ctypes.CDLL:
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
        lib = ctypes.CDLL(str(LIB_PATH))
    except OSError as e:  # pragma: no cover
        raise MgsLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise MgsLibraryError(f"{LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().mgs_last_error_string()
        raise MgsError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load().mgs_launch_count())
