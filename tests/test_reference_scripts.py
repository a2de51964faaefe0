"""The REAL reference scripts, unmodified, on the CPU (build container only: /root/reference does not exist on the
GPU box).  BASELINE.json configs[0] -- "inference (test.py) on the independent test-set CSV ... on CPU (reference
plumbing)" -- with a synthetic 961-molecule CSV (gnnexplainer.py:1439: the test set has 961 molecules; the CSV files
themselves are not shipped).  See tests/run_reference_script.py for what is real and what is substituted."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

REF = Path("/root/reference")
ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.skipif(not REF.exists(), reason="/root/reference is only mounted in the build container")


def write_csv(path, n, first_seed):
    rng = np.random.default_rng(first_seed)
    pd.DataFrame({"Smiles": [f"SYN{first_seed + k}" for k in range(n)],
                  "pchembl": np.round(rng.normal(6.5, 1.2, n), 3)}).to_csv(path, index=False)


def run(script, cwd, max_steps=None, timeout=900):
    cmd = [sys.executable, str(ROOT / "tests" / "run_reference_script.py")]
    if max_steps is not None:
        cmd += ["--max-steps", str(max_steps)]
    cmd.append(str(script))
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-5000:]
    return r.stdout


def test_rdkit_stand_in_reproduces_the_generator_through_the_scripts_own_featuriser():
    """train.py's `smiles_to_graph` (compiled from the reference source) on the stand-in molecules gives back exactly the
    synthetic generator's features and edge order -- the featuriser facts of SURVEY.md Appendix C."""
    import ast
    import random as _random

    import torch
    sys.path.insert(0, str(ROOT / "tests" / "stubs"))
    from rdkit import Chem
    from m_gat_graphsage_b200.synth import synth_batch
    tree = ast.parse((REF / "train.py").read_text(encoding="utf-8"))
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("one_of_k_encoding_unk", "smiles_to_graph")]
    ns = {"Chem": Chem, "np": np, "torch": torch, "random": _random}
    exec(compile(ast.Module(body=body, type_ignores=[]), "train.py", "exec"), ns)
    for seed in (1, 7, 12345):
        x, ei = ns["smiles_to_graph"](f"SYN{seed}")
        b = synth_batch(1, seed)
        assert torch.equal(x, b.x) and torch.equal(ei, b.edge_index)
    assert Chem.MolFromSmiles("c1ccccc1") is None


def test_train_py_then_test_py_run_unchanged_on_the_cpu(tmp_path):
    """train.py (one epoch + one step of the next, so its best-model checkpoint incl. the pickled scaler is written) and
    then test.py on a 961-molecule CSV, one molecule per forward as the script does (test.py:175-208)."""
    write_csv(tmp_path / "train_data.csv", 96, 10_000)
    write_csv(tmp_path / "validation_data.csv", 32, 20_000)
    out = run(REF / "train.py", tmp_path, max_steps=2)
    assert "Epoch    1 | Train Loss" in out and "New best model saved at epoch 1" in out and "STEP_BUDGET_REACHED" in out
    assert (tmp_path / "best_model.pth").exists()
    write_csv(tmp_path / "test_data.csv", 961, 30_000)
    out = run(REF / "test.py", tmp_path)
    assert "Number of test samples: 961" in out and "Pearson correlation" in out
    res = pd.read_csv(tmp_path / "model_prediction_results.csv")
    assert len(res) == 961 and np.isfinite(res["Predicted_Value"]).all()


def _csv_names(script):
    """The string constants a script assigns to train_csv_file / test_csv_file (Windows absolute paths in several
    scripts: on Linux those are plain file names containing backslashes, relative to the working directory)."""
    import ast
    out = {}
    for node in ast.parse(script.read_text(encoding="utf-8")).body:
        if isinstance(node, ast.Assign) and isinstance(node.value, ast.Constant) and isinstance(node.value.value, str):
            for t in node.targets:
                if isinstance(t, ast.Name) and t.id in ("train_csv_file", "test_csv_file"):
                    out[t.id] = node.value.value
    return out


def test_model1_script_runs_unchanged_on_the_cpu(tmp_path):
    """ablation/model1.py -- the north-star GATConv + SAGEConv model: loaders of batch 64 / 32, Adam 1e-4, MSE."""
    script = REF / "ablation" / "model1.py"
    names = _csv_names(script)
    assert set(names) == {"train_csv_file", "test_csv_file"}, names
    write_csv(tmp_path / names["train_csv_file"], 130, 40_000)
    write_csv(tmp_path / names["test_csv_file"], 40, 50_000)
    out = run(script, tmp_path, max_steps=4)
    assert "Epoch 1, Loss" in out and "Improved model found at epoch 1" in out
