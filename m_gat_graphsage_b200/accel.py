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


class FusedAdam(torch.optim.Optimizer):
    """``torch.optim.Adam`` (train.py:216-222, ablation/model1.py:113: L2 ``weight_decay``, no amsgrad) with the whole
    update of a parameter group as ONE launch of ``mgs_adam_step`` (csrc/adam.cu): PyTorch's fused implementation deals
    64 Ki-element chunks to CTAs -- 26 CTAs for the 1.6 M parameters of the model1 trunk, 80 us per step on a B200 --
    this one 1024-element chunks to 4 CTAs per SM.  Same constructor arguments and ``state_dict`` layout
    (``step`` / ``exp_avg`` / ``exp_avg_sq`` per parameter) as ``torch.optim.Adam``; fp32 CUDA parameters only."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("FusedAdam: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self._tables = {}

    @torch.no_grad()
    def step(self, closure=None):
        import ctypes

        from . import _lib
        from .functional import device_guard, stream_ptr
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            steps = set()
            for p in ps:
                if not (p.is_cuda and p.dtype == torch.float32 and p.grad.dtype == torch.float32 and not p.grad.is_sparse):
                    raise RuntimeError("FusedAdam: fp32 CUDA parameters with dense gradients only (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] = int(st["step"]) + 1
                steps.add(st["step"])
            # one launch needs one step number; parameters that joined later (rare) go in their own launch
            for step_no in sorted(steps):
                sel = [p for p in ps if self.state[p]["step"] == step_no]
                grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in sel]
                if any(not p.is_contiguous() for p in sel):
                    raise RuntimeError("FusedAdam: parameters must be contiguous")
                key = (gi, step_no == max(steps), tuple(p.data_ptr() for p in sel), tuple(g.data_ptr() for g in grads))
                tab = self._tables.get(key[:2])
                if tab is None or tab[0] != key:
                    n = len(sel)
                    arr = lambda vals: (ctypes.c_void_p * n)(*vals)       # noqa: E731
                    tab = (key, arr([p.data_ptr() for p in sel]), arr([g.data_ptr() for g in grads]),
                           arr([self.state[p]["exp_avg"].data_ptr() for p in sel]),
                           arr([self.state[p]["exp_avg_sq"].data_ptr() for p in sel]),
                           (ctypes.c_int64 * n)(*[p.numel() for p in sel]), n)
                    self._tables[key[:2]] = tab
                b1, b2 = group["betas"]
                with device_guard(sel[0].device):
                    rc = lib.mgs_adam_step(tab[6], tab[1], tab[2], tab[3], tab[4], tab[5], float(group["lr"]), float(b1),
                                           float(b2), float(group["eps"]), float(group["weight_decay"]), step_no,
                                           stream_ptr())
                _lib.check(rc, "mgs_adam_step")
        return loss
