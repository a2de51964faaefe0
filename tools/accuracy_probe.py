"""Gradient accuracy of the model1 trunk vs the CPU oracle for the three K4 configurations."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch, torch.nn.functional as F
import ref_trunks
from m_gat_graphsage_b200 import nn as mnn
from m_gat_graphsage_b200.accel import use_mgs_linear
from m_gat_graphsage_b200.data import Data
from m_gat_graphsage_b200.synth import synth_batch
from oracle import pyg_oracle as O

torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 300
def rel(a, c):
    a, c = a.detach().cpu().double(), c.detach().double()
    return float((a - c).abs().max()) / max(float(c.abs().max()), 1e-30)
ref = ref_trunks.build_trunk("model1", O, seed=42).eval()
b = synth_batch(nmol, 99)
x = b.x + 0.05 * torch.randn(b.x.shape, generator=torch.Generator().manual_seed(3))
d_ref = Data(x=x.clone().requires_grad_(True), edge_index=b.edge_index, batch=b.batch)
out_r = ref(d_ref)
gr = torch.autograd.grad(F.mse_loss(out_r.view(-1), b.y), list(ref.parameters()) + [d_ref.x])
# fp64 oracle = ground truth for conditioning
ref64 = ref_trunks.build_trunk("model1", O, seed=42).double().eval()
d64 = Data(x=x.double().requires_grad_(True), edge_index=b.edge_index, batch=b.batch)
g64 = torch.autograd.grad(F.mse_loss(ref64(d64).view(-1), b.y.double()), list(ref64.parameters()) + [d64.x])
names = [k for k, _ in ref.named_parameters()] + ["x"]
print(f"{'tensor':22s} {'oracle32 vs 64':>15s} " + " ".join(f"{c:>15s}" for c in ("all FFMA", "conv TC", "all TC")))
cols = []
for cfg in ("ffma", "conv", "all"):
    if cfg == "ffma": os.environ["MGS_DISABLE_TC"] = "1"
    else: os.environ.pop("MGS_DISABLE_TC", None)
    mine = ref_trunks.build_trunk("model1", mnn, seed=1)
    mine.load_state_dict(ref.state_dict()); mine = mine.to(dev).eval()
    if cfg in ("all", "ffma"): use_mgs_linear(mine)
    d = Data(x=x.to(dev).requires_grad_(True), edge_index=b.edge_index.to(dev), batch=b.batch.to(dev))
    out = mine(d)
    gg = torch.autograd.grad(F.mse_loss(out.view(-1), b.y.to(dev)), list(mine.parameters()) + [d.x])
    cols.append([rel(out, ref64(d64))] + [rel(a, c) for a, c in zip(gg, g64)])
o32 = [rel(out_r, ref64(d64))] + [rel(a, c) for a, c in zip(gr, g64)]
for i, n in enumerate(["logits"] + names):
    print(f"{n:22s} {o32[i]:15.2e} " + " ".join(f"{c[i]:15.2e}" for c in cols))
