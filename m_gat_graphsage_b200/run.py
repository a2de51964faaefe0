"""Run an UNMODIFIED reference script on the B200 path:

    python -m m_gat_graphsage_b200.run /path/to/train.py [script args...]

* puts the ``torch_geometric`` import shim first on ``sys.path`` so that
  ``from torch_geometric.nn import GATConv, SAGEConv, global_max_pool`` etc. resolve to the sm_100a kernels;
* makes CUDA the default device -- the reference scripts never call ``.to(device)`` (train.py is CPU-only as
  written) and the operators have no CPU fallback;
* keeps fp32 semantics: cuDNN's TF32 convolutions (PyTorch's default) are switched off, they would put 1e-3
  errors into ``ModifiedGATLayer``'s Conv1d (train.py:83-84);
* routes ``nn.Linear`` (readout MLP) through the K4 projection kernels (``--no-mgs-linear`` keeps cuBLAS);
* routes the ``ModifiedGATLayer`` the script declares (train.py:77-99) through the K5 streaming attention
  (``--no-mgs-attention`` keeps the script's dense ``[N, N]`` code);
* ``--max-steps N`` ends the run cleanly after N optimiser steps (the scripts hard-code 1000 epochs) and prints the time
  per step; ``torch.load`` gets its pre-2.6 default back (the scripts' checkpoints hold a pickled sklearn scaler);
* turns on the activation peephole (``m_gat_graphsage_b200.lazy``): the ``relu`` / ``elu`` the script applies to a conv
  layer's output is fused into that layer's last kernel (``--no-activation-fusion`` keeps them separate launches).
"""
from __future__ import annotations

import runpy
import sys
from pathlib import Path


class StepBudgetReached(SystemExit):
    """Raised (exit code 0) from ``optimizer.step()`` when ``--max-steps`` optimisation steps have been taken."""


def install_step_budget(max_steps: int, report=None):
    """The reference scripts hard-code 1000 epochs (train.py:229, ablation/model1.py:120): ``--max-steps N`` ends the run
    cleanly after N calls of ``Optimizer.step`` (smoke runs, benchmarks).  ``report``: optional callable invoked with the
    list of host time stamps (seconds) taken at every step."""
    import time

    import torch
    stamps = []
    original = torch.optim.Optimizer.step

    def patch(cls):
        inner = cls.step

        def step(self, *a, **kw):
            out = inner(self, *a, **kw)
            stamps.append(time.perf_counter())
            if len(stamps) >= max_steps:
                if report is not None:
                    report(stamps)
                raise StepBudgetReached(0)
            return out
        cls.step = step

    seen = set()
    for name in dir(torch.optim):
        cls = getattr(torch.optim, name)
        if isinstance(cls, type) and issubclass(cls, torch.optim.Optimizer) and cls.step is not original \
                and cls not in seen:
            seen.add(cls)
            patch(cls)
    return stamps


def legacy_torch_load():
    """The reference targets torch 2.4 (README.md:22-32), where ``torch.load`` unpickles arbitrary objects by default;
    train.py:284-295 stores a fitted sklearn ``StandardScaler`` in its checkpoint and test.py:160-164 reads it back.
    torch >= 2.6 defaults to ``weights_only=True`` and refuses: restore the old default for the script's own files."""
    import functools

    import torch
    if getattr(torch.load, "_mgs_legacy", False):
        return
    original = torch.load

    @functools.wraps(original)
    def load(*a, **kw):
        kw.setdefault("weights_only", False)
        return original(*a, **kw)

    load._mgs_legacy = True
    torch.load = load


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    use_linear = True
    if "--no-mgs-linear" in argv:
        argv.remove("--no-mgs-linear")
        use_linear = False
    use_attention = True
    if "--no-mgs-attention" in argv:
        argv.remove("--no-mgs-attention")
        use_attention = False
    max_steps = None
    if "--max-steps" in argv:
        i = argv.index("--max-steps")
        max_steps = int(argv[i + 1])
        del argv[i:i + 2]
    fuse_act = True
    if "--no-activation-fusion" in argv:
        argv.remove("--no-activation-fusion")
        fuse_act = False
    if not argv:
        raise SystemExit(__doc__)
    shim = str(Path(__file__).resolve().parent / "shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    import torch

    from . import _lib
    _lib.load()                                    # fail loudly before the script starts if libmgs.so is missing
    if not torch.cuda.is_available():
        raise RuntimeError("m_gat_graphsage_b200.run needs a CUDA device: the operators have no CPU fallback")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_default_device("cuda")
    if use_linear:
        from .accel import patch_torch_linear
        patch_torch_linear()
    if use_attention:
        from .attention import patch_layer_classes
        patch_layer_classes()
    from .lazy import set_activation_fusion
    set_activation_fusion(fuse_act)
    legacy_torch_load()
    if max_steps is not None:
        def report(stamps):
            torch.cuda.synchronize()
            if len(stamps) > 2:
                ms = 1e3 * (stamps[-1] - stamps[1]) / (len(stamps) - 2)
                print(f"[m_gat_graphsage_b200.run] {len(stamps)} optimisation steps, {ms:.3f} ms per step (host clock, "
                      "first step excluded)", file=sys.stderr)
        install_step_budget(max_steps, report)
    script = argv[0]
    sys.argv = argv
    try:
        runpy.run_path(script, run_name="__main__")
    except StepBudgetReached:
        pass


if __name__ == "__main__":
    main()
