#!/usr/bin/env python
"""Pin the oracle to the REAL PyG.  ``torch_geometric`` cannot be installed in the build
container or on the GPU box (no network), so this script is for a maintainer's machine that has it:

    pip install torch torch_geometric        # any 2.3 <= version <= 2.6, CPU is enough
    python tools/dump_pyg_golden.py           # writes tests/golden/pyg/*.pt

It runs the genuine ``torch_geometric.nn.{GATConv, SAGEConv, global_*_pool}`` on the same seeded synthetic molecules
and weights the tests use and stores inputs, weights, outputs and gradients.  Once the files exist,
``tests/test_oracle.py::test_oracle_matches_real_pyg_dump`` compares ``oracle/pyg_oracle.py`` with them on every run
(and the CUDA operators are compared with the oracle everywhere else), which turns "parity unpinned" into "pinned
against torch_geometric <version>".  Nothing from PyG is copied: only tensors are stored.
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
OUT = ROOT / "tests" / "golden" / "pyg"

CASES = [  # (name, layer, ctor kwargs, in_channels)
    ("gat_35_35_h10", "GATConv", dict(in_channels=35, out_channels=35, heads=10), 35),
    ("gat_350_128_h1", "GATConv", dict(in_channels=350, out_channels=128, heads=1), 350),
    ("gat_35_32_h8_mean", "GATConv", dict(in_channels=35, out_channels=32, heads=8, concat=False), 35),
    ("sage_35_35", "SAGEConv", dict(in_channels=35, out_channels=35), 35),
    ("sage_350_350", "SAGEConv", dict(in_channels=350, out_channels=350), 350),
]


def main():
    try:
        import torch_geometric
        from torch_geometric import nn as gnn
    except ImportError:
        sys.exit("torch_geometric is not installed here: run this on a machine that has it (see the docstring)")
    from m_gat_graphsage_b200.synth import synth_batch
    OUT.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(1)
    b = synth_batch(16, 4242)
    for name, layer, kw, fin in CASES:
        torch.manual_seed(7)
        conv = getattr(gnn, layer)(**kw)
        x = (b.x if fin == 35 else torch.randn(b.x.size(0), fin, generator=torch.Generator().manual_seed(3)))
        x = x.clone().requires_grad_(True)
        out = conv(x, b.edge_index)
        w = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
        grads = torch.autograd.grad((out * w).sum(), [x] + list(conv.parameters()))
        torch.save({"pyg_version": torch_geometric.__version__, "torch_version": torch.__version__, "layer": layer,
                    "kwargs": kw, "x": x.detach(), "edge_index": b.edge_index, "state_dict": conv.state_dict(),
                    "out": out.detach(), "cotangent": w, "x_grad": grads[0],
                    "param_grads": {k: g for (k, _), g in zip(conv.named_parameters(), grads[1:])}}, OUT / f"{name}.pt")
        print("wrote", OUT / f"{name}.pt")
    x = torch.randn(b.x.size(0), 7, generator=torch.Generator().manual_seed(9)).requires_grad_(True)
    pools = {}
    for pname in ("global_max_pool", "global_mean_pool", "global_add_pool"):
        out = getattr(gnn, pname)(x, b.batch)
        (gx,) = torch.autograd.grad(out.sum(), x)
        pools[pname] = {"out": out.detach(), "x_grad": gx}
    torch.save({"pyg_version": torch_geometric.__version__, "x": x.detach(), "batch": b.batch, "pools": pools},
               OUT / "pools.pt")
    print("wrote", OUT / "pools.pt")


if __name__ == "__main__":
    main()
