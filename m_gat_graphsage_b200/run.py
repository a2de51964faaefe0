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
* turns on the activation peephole (``m_gat_graphsage_b200.lazy``): the ``relu`` / ``elu`` the script applies to a conv
  layer's output is fused into that layer's last kernel (``--no-activation-fusion`` keeps them separate launches).
"""
from __future__ import annotations

import runpy
import sys
from pathlib import Path


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
    script = argv[0]
    sys.argv = argv
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
