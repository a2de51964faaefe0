"""Generate the golden fixtures in this directory.  Run ONLY in the build container, where
``/root/reference`` is mounted (it does not exist on the GPU box):

    python tests/golden/make_golden.py

What is pinned, and against what:

* The reference's own model classes (``GAT_GraphSAGE`` of ablation/model1.py and train.py incl.
  ``ModifiedGATLayer``, ``GATNet`` of gnn/gat.py, ``SAGENet`` of gnn/graphsage.py) are extracted from
  the reference SOURCE with ``ast`` (the scripts cannot be imported: module-level code needs RDKit
  and CSV files) and executed with the CPU oracle's operators bound to the ``torch_geometric`` names.
* ``ref_trunks.py`` (our re-declaration used on the GPU box) must reproduce them BIT-EXACTLY from the
  same ``state_dict`` -- this pins model wiring, layer names and ``ModifiedGATLayer``.
* Outputs (logits, parameter-gradient checksums, per-atom gradient-L2 importances) are stored so that
  the oracle itself cannot drift silently and the CUDA path is compared on identical inputs.

The PyG operator arithmetic itself stays "parity unpinned" (PyG is not installable here, see
``oracle/pyg_oracle.py``).  No reference source is copied into the repository: classes are compiled
in memory from where they lie.
"""
from __future__ import annotations

import ast
import os
import sys
from pathlib import Path

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import ref_trunks  # noqa: E402
from m_gat_graphsage_b200.data import Data  # noqa: E402
from m_gat_graphsage_b200.synth import synth_batch  # noqa: E402
from oracle import pyg_oracle as O  # noqa: E402

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent

CASES = {
    # name: (reference file, [classes to extract], model class, mirror name, batch seed, num molecules)
    "model1": ("ablation/model1.py", ["GAT_GraphSAGE"], "GAT_GraphSAGE", "model1", 1001, 12),
    "gat": ("gnn/gat.py", ["GATNet"], "GATNet", "gat", 1002, 12),
    "graphsage": ("gnn/graphsage.py", ["SAGENet"], "SAGENet", "graphsage", 1003, 12),
    "train": ("train.py", ["ModifiedGATLayer", "GAT_GraphSAGE"], "GAT_GraphSAGE", "train", 1004, 12),
}


def extract_classes(path: Path, names):
    tree = ast.parse(path.read_text(encoding="utf-8"))
    body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in names]
    assert len(body) == len(names), f"{path}: expected classes {names}"
    ns = {"torch": torch, "nn": nn, "F": F, "GATConv": O.GATConv, "SAGEConv": O.SAGEConv,
          "global_max_pool": O.global_max_pool, "gap": O.global_mean_pool,
          "global_mean_pool": O.global_mean_pool, "Data": Data}
    exec(compile(ast.Module(body=body, type_ignores=[]), str(path), "exec"), ns)
    return ns


def checksum(sd):
    return {k: float(v.double().abs().sum()) for k, v in sd.items()}


def full_train_model_fixture():
    """train.py's three networks (GNN trunk + CNNNet + CombinedNet) and its loss `mse + 0.001 * kl` (train.py:212-214,
    236-249), compiled from the reference source and run on the oracle operators -> tests/golden/train_full.pt."""
    names = ["ModifiedGATLayer", "GAT_GraphSAGE", "CNNNet", "CombinedNet"]
    ns = extract_classes(REF / "train.py", names)
    tree = ast.parse((REF / "train.py").read_text(encoding="utf-8"))
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "kl_loss"]
    exec(compile(ast.Module(body=fn, type_ignores=[]), "train.py", "exec"), ns)
    torch.manual_seed(42)
    gnn, cnn, comb = ns["GAT_GraphSAGE"](n_output=1, num_features_xd=35), ns["CNNNet"](input_dim=1024, output_dim=1024), \
        ns["CombinedNet"](input_dim=1025, hidden_dim=512, output_dim=1)
    for m in (gnn, cnn, comb):
        m.eval()                                                    # dropout off: deterministic fixture
    mine = ref_trunks.TrainPyModel(O).eval()
    mine.gat_graphsage_model.load_state_dict(gnn.state_dict(), strict=True)
    mine.cnn_model.load_state_dict(cnn.state_dict(), strict=True)
    mine.combined_model.load_state_dict(comb.state_dict(), strict=True)
    nmol = 8
    batch = synth_batch(nmol, 1005)
    gen = torch.Generator().manual_seed(7)
    ecfp = (torch.rand(nmol, 1, 1024, generator=gen) < 0.05).float()
    d = Data(x=batch.x, edge_index=batch.edge_index, batch=batch.batch)
    g_out = gnn(d)
    combined = torch.cat((g_out, cnn(ecfp)), dim=1)
    final = comb(combined)
    loss = F.mse_loss(final, batch.y.view(-1, 1)) + 0.001 * ns["kl_loss"](combined)
    d2 = Data(x=batch.x, edge_index=batch.edge_index, batch=batch.batch)
    d2.y = batch.y
    final_m, combined_m = mine(d2, ecfp)
    assert torch.equal(final, final_m) and torch.equal(combined, combined_m), "TrainPyModel mirror differs"
    assert torch.equal(loss, mine.loss(d2, ecfp)), "TrainPyModel loss differs from train.py's"
    params = list(gnn.parameters()) + list(cnn.parameters()) + list(comb.parameters())
    names_p = ["gat_graphsage_model." + k for k, _ in gnn.named_parameters()] + \
              ["cnn_model." + k for k, _ in cnn.named_parameters()] + ["combined_model." + k for k, _ in comb.named_parameters()]
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    # the 33.5 M-element fc1 gradient is stored as 4 of its 256 rows, other matrices above 64 k elements as every
    # `stride`-th row; everything else in full
    store, strides = {}, {}
    for k, g in zip(names_p, grads):
        strides[k] = 1
        if g is None:
            store[k] = None
            continue
        if g.numel() > (1 << 20):
            strides[k] = 64
        elif g.numel() > 65536:
            strides[k] = -(-g.numel() // 65536)
        store[k] = g.detach()[::strides[k]].clone()
    fixture = {"reference_file": "train.py", "num_molecules": nmol, "x": batch.x, "edge_index": batch.edge_index,
               "batch": batch.batch, "y": batch.y, "ecfp": ecfp, "final": final.detach(), "combined": combined.detach(),
               "loss": loss.detach(), "param_grads": store, "param_grad_row_stride": strides, "weights_seed": 42,
               "state_checksum": {k: float(v.double().abs().sum()) for k, v in mine.state_dict().items()},
               "torch_version": torch.__version__}
    torch.save(fixture, OUT / "train_full.pt")
    print(f"train_full: loss={float(loss):.6f} -> {OUT / 'train_full.pt'} ({os.path.getsize(OUT / 'train_full.pt') / 1024:.0f} KiB)")


def main():
    assert REF.exists(), "/root/reference is not mounted; golden fixtures can only be generated in the build container"
    torch.set_num_threads(1)
    full_train_model_fixture()
    for name, (rel, classes, cls_name, mirror, seed, nmol) in CASES.items():
        ns = extract_classes(REF / rel, classes)
        torch.manual_seed(42)
        ref_model = ns[cls_name]()
        ref_model.eval()
        mine = ref_trunks.TRUNKS[mirror](O)
        mine.load_state_dict(ref_model.state_dict(), strict=True)
        mine.eval()

        batch = synth_batch(nmol, seed)
        x = batch.x.clone().requires_grad_(True)
        d = Data(x=x, edge_index=batch.edge_index, batch=batch.batch)
        out_ref = ref_model(d)
        out_mine = mine(Data(x=batch.x, edge_index=batch.edge_index, batch=batch.batch))
        assert torch.equal(out_ref, out_mine), f"{name}: ref_trunks mirror differs from the reference class"

        loss = F.mse_loss(out_ref.view(-1), batch.y)
        params = [p for p in ref_model.parameters()]
        grads = torch.autograd.grad(loss, params, retain_graph=True, allow_unused=True)
        gx, = torch.autograd.grad(out_ref.sum(), x)
        grad_sums = {k: (float(g.double().abs().sum()) if g is not None else None)
                     for (k, _), g in zip(ref_model.named_parameters(), grads)}
        # every gradient ELEMENT is pinned; matrices above 64 k elements store every `stride`-th row (the fixture
        # stays small: the fc_g1 / fc_g2 matrices of model1 alone are 5 MB)
        strides = {k: (1 if g is None or g.numel() <= 65536 else -(-g.numel() // 65536))
                   for (k, _), g in zip(ref_model.named_parameters(), grads)}
        param_grads = {k: (g.detach()[:: strides[k]].clone() if g is not None else None)
                       for (k, _), g in zip(ref_model.named_parameters(), grads)}
        fixture = {
            "reference_file": rel, "seed": seed, "num_molecules": nmol, "weights_seed": 42,
            "x": batch.x, "edge_index": batch.edge_index, "batch": batch.batch, "y": batch.y,
            "state_checksum": checksum(ref_model.state_dict()),
            "logits": out_ref.detach(), "loss": loss.detach(),
            "param_grad_abs_sums": grad_sums,
            "param_grads": param_grads, "param_grad_row_stride": strides,
            "grad_conv_first": next(g for g in grads if g is not None).detach(),
            "atom_importance": torch.norm(gx, dim=1).detach(),
            "x_grad": gx.detach(),
            "torch_version": torch.__version__,
        }
        torch.save(fixture, OUT / f"{name}.pt")
        print(f"{name}: logits[:3]={out_ref.view(-1)[:3].tolist()} -> {OUT / (name + '.pt')} "
              f"({os.path.getsize(OUT / (name + '.pt')) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
