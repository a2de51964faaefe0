"""Route the dense layers of the reference's model classes (``fc_g1`` / ``fc_g2`` / ``out`` of the readout MLP,
train.py:107-111,120-123; ablation/model1.py:59-64,73-76) through the K4 kernels.

The layers are plain ``torch.nn.Linear`` modules declared inside the reference scripts, so they cannot be
replaced by name.  ``use_mgs_linear(model)`` rebinds ``forward`` on every ``nn.Linear`` instance of a model;
``patch_torch_linear()`` does the same process-wide (what ``python -m m_gat_graphsage_b200.run`` does).  Only
CUDA fp32 inputs are rerouted; parameters, ``state_dict`` keys and autograd semantics are untouched."""
from __future__ import annotations

import types

import torch
import torch.nn as nn

from . import functional as F_
from .lazy import PendingActivation, activation_fusion_enabled

_ORIG_FORWARD = nn.Linear.forward


def _mgs_forward(self: nn.Linear, x: torch.Tensor) -> torch.Tensor:
    x = F_.real(x)
    if x.is_cuda and x.dtype == torch.float32 and self.weight.dtype == torch.float32 and x.dim() >= 2:
        if activation_fusion_enabled() and x.dim() == 2:
            # `self.relu(self.fc_g1(x))` (ablation/model1.py:74): the ReLU rides on the GEMM's epilogue (lazy.py)
            def finish(act, x=x):
                if act == "relu":
                    return F_.linear(x, self.weight, self.bias, activation="relu")
                out = F_.linear(x, self.weight, self.bias)
                return out if act is None else torch.nn.functional.elu(out)
            needs_grad = torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad)
            return PendingActivation(finish, (x.size(0), self.weight.size(0)), x.dtype, x.device, needs_grad)
        return F_.linear(x, self.weight, self.bias)
    return _ORIG_FORWARD(self, x)


def use_mgs_linear(model: nn.Module) -> int:
    """Rebind every ``nn.Linear`` of ``model`` to the sm_100a projection kernels; returns how many."""
    n = 0
    for m in model.modules():
        if type(m) is nn.Linear:
            m.forward = types.MethodType(_mgs_forward, m)
            n += 1
    return n


def patch_torch_linear() -> None:
    nn.Linear.forward = _mgs_forward


def unpatch_torch_linear() -> None:
    nn.Linear.forward = _ORIG_FORWARD
