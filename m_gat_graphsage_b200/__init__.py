"""B200-native (sm_100a) message-passing hot path for M-GAT-GraphSAGE.

Public surface = the PyG-shaped operator API the reference scripts import
(SURVEY.md section 8b): ``nn.GATConv``, ``nn.SAGEConv``, ``nn.global_*_pool``,
``data.Data/Batch/DataLoader``, ``explain.*``.  All arithmetic runs in the
hand-written CUDA kernels of ``csrc/`` reached through the C ABI declared in
``include/mgs.h`` (``libmgs.so``).  There is no CPU fallback: operators raise
on CPU tensors or when the library is missing.
"""
__version__ = "0.1.0"

from . import data  # noqa: F401  (host-only; importable without a GPU)
