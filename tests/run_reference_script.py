"""TEST INFRASTRUCTURE: run an UNMODIFIED reference script on the CPU, for the host-side plumbing only.

    python tests/run_reference_script.py [--max-steps N] /root/reference/train.py

What is real: the script's own text (featuriser, model classes, loaders, training / evaluation loop, checkpointing),
this repository's ``Data`` / ``Batch`` / ``DataLoader`` (host-only product code) behind the ``torch_geometric.data`` name,
the launcher's step budget and ``torch.load`` compatibility (``m_gat_graphsage_b200.run``).
What is substituted, because the build container has neither a GPU nor RDKit: the operators behind
``torch_geometric.nn`` are the CPU oracle's (``oracle/pyg_oracle.py`` -- the product operators refuse CPU tensors by
design), and ``rdkit`` is the synthetic stand-in of ``tests/stubs``.  Never part of the product path."""
import runpy
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "stubs"))


def install_cpu_torch_geometric():
    from m_gat_graphsage_b200 import data as mdata
    from oracle import pyg_oracle as O
    tg = types.ModuleType("torch_geometric")
    tg.__version__ = "0.0+oracle-cpu"
    nn_mod = types.ModuleType("torch_geometric.nn")
    for name in ("GATConv", "SAGEConv", "GCNConv", "GINConv", "global_max_pool", "global_mean_pool", "global_add_pool"):
        setattr(nn_mod, name, getattr(O, name))
    data_mod = types.ModuleType("torch_geometric.data")
    data_mod.Data, data_mod.Batch, data_mod.DataLoader = mdata.Data, mdata.Batch, mdata.DataLoader
    loader_mod = types.ModuleType("torch_geometric.loader")
    loader_mod.DataLoader = mdata.DataLoader
    tg.nn, tg.data, tg.loader = nn_mod, data_mod, loader_mod
    sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": nn_mod, "torch_geometric.data": data_mod,
                        "torch_geometric.loader": loader_mod})


def main():
    argv = sys.argv[1:]
    max_steps = None
    if argv and argv[0] == "--max-steps":
        max_steps = int(argv[1])
        argv = argv[2:]
    from m_gat_graphsage_b200.run import StepBudgetReached, install_step_budget, legacy_torch_load
    install_cpu_torch_geometric()
    legacy_torch_load()
    if max_steps is not None:
        install_step_budget(max_steps)
    sys.argv = argv
    try:
        runpy.run_path(argv[0], run_name="__main__")
    except StepBudgetReached:
        print("STEP_BUDGET_REACHED")


if __name__ == "__main__":
    main()
