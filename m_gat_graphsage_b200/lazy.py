"""Activation peephole across the reference's module boundary.

Every reference model applies an element-wise activation to a conv layer's output in its OWN ``forward``:
``self.relu(self.conv1(x, edge_index))`` (ablation/model1.py:68-71, train.py:115-118), ``F.elu(self.gcn1(..))``
(gnn/gat.py:63).  The layer cannot know that when it is called -- unless it answers with a promise.

When the peephole is on (``set_activation_fusion(True)``; the ``torch_geometric`` import shim and the launcher turn it
on), an eligible conv layer returns a ``PendingActivation``: a storage-less ``torch.Tensor`` subclass that carries the
shape / dtype / device of the result and a thunk that launches the layer's LAST kernel.  The first torch function
applied to it decides:

* ``relu`` / ``elu(alpha=1)`` (function, method, ``nn.ReLU`` module, in-place spellings)  ->  the thunk runs with the
  activation fused into the kernel's epilogue, and the ReLU's backward is fused into whichever of our kernels produces
  the gradient (``functional._act_backward``); the result is an ordinary tensor;
* anything else  ->  the thunk runs without activation and the function is applied to the ordinary result.

Either way the numbers are the ones eager PyTorch would produce (same comparisons as ATen's threshold / elu kernels);
nothing is speculated.  Metadata reads (``.shape``, ``.size()``, ``.dtype`` ...) do not trigger the thunk.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.nn.functional as F
from torch.utils._pytree import tree_map

_ENABLED = False


def set_activation_fusion(on: bool = True) -> bool:
    """Switch the peephole on / off process-wide; returns the previous setting."""
    global _ENABLED
    prev, _ENABLED = _ENABLED, bool(on)
    return prev


def activation_fusion_enabled() -> bool:
    return _ENABLED


_RELU = {torch.relu, F.relu, torch.Tensor.relu, torch.relu_, torch.Tensor.relu_, F.relu_}
_ELU = {F.elu, F.elu_}
_T = torch.Tensor
_METADATA = {_T.shape.__get__, _T.size, _T.dim, _T.ndim.__get__, _T.dtype.__get__, _T.device.__get__,
             _T.is_cuda.__get__, _T.requires_grad.__get__, _T.numel, _T.nelement, _T.ndimension, _T.layout.__get__,
             _T.is_floating_point, _T.is_complex, _T.is_sparse.__get__, _T.element_size, _T.__len__, _T.get_device,
             _T.is_contiguous, _T.is_leaf.__get__}


class PendingActivation(torch.Tensor):
    """See the module docstring.  ``thunk(activation)`` -> the layer's real output (``activation`` is ``None``,
    ``'relu'`` or ``'elu'``)."""

    @staticmethod
    def __new__(cls, thunk: Callable[[Optional[str]], torch.Tensor], shape, dtype, device, requires_grad: bool):
        t = torch.Tensor._make_wrapper_subclass(cls, tuple(shape), dtype=dtype, device=device,
                                                requires_grad=bool(requires_grad))
        t._thunk = thunk
        t._plain = None        # the un-activated result, once somebody needed it
        t._alias = None        # after an IN-PLACE activation the promise stands for the activated tensor
        return t

    def _materialize(self, activation: Optional[str] = None) -> torch.Tensor:
        if self._alias is not None:
            if activation is None:
                return self._alias
            return F.relu(self._alias) if activation == "relu" else F.elu(self._alias)
        if activation is None:
            if self._plain is None:
                self._plain = self._thunk(None)
            return self._plain
        return self._thunk(activation)

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):  # pragma: no cover
        # never reached through the Python API (``__torch_function__`` below sees every call first and hands ordinary
        # tensors on); C++ callers that bypass it get the plain result
        def real(a):
            return a._materialize(None) if isinstance(a, PendingActivation) else a
        return func(*tree_map(real, args), **tree_map(real, kwargs or {}))

    def __repr__(self):  # pragma: no cover
        return f"PendingActivation(shape={tuple(self.shape)}, device={self.device})"

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        first = args[0] if args else None
        if func in _METADATA and isinstance(first, PendingActivation):
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        if isinstance(first, PendingActivation) and first._alias is None and first._plain is None:
            act = None
            if func in _RELU and len(args) == 1:
                act = "relu"
            elif func in _ELU and len(args) == 1 and float(kwargs.get("alpha", 1.0)) == 1.0:
                act = "elu"
            if act is not None:
                out = first._materialize(act)
                if kwargs.get("inplace", False) or func in (torch.relu_, torch.Tensor.relu_, F.relu_, F.elu_):
                    first._alias = out
                return out

        def real(a):
            return a._materialize(None) if isinstance(a, PendingActivation) else a

        return func(*tree_map(real, args), **tree_map(real, kwargs))
