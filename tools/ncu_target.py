"""ncu target: ONE SAGE aggregate fwd+bwd and ONE GAT message fwd+bwd at the model1 shapes on post-ReLU-like data,
inside cudaProfilerStart/Stop (use `ncu --profile-from-start off`).  Profiling aid, never a bench value."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from m_gat_graphsage_b200 import functional as Fm
from m_gat_graphsage_b200.graph import build_graph_index
from m_gat_graphsage_b200.synth import synth_batch

dev = torch.device("cuda:0")
b = synth_batch(4096, 42, device=dev)
N = b.x.size(0)
gi = build_graph_index(b.edge_index, N)
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.relu(torch.randn(N, 350, device=dev, generator=gen))
go = torch.randn(N, 350, device=dev, generator=gen) * (torch.rand(N, 350, device=dev, generator=gen) < 0.3)
att = torch.randn(2, 350, device=dev, generator=gen)


def once():
    xr = x.detach().requires_grad_(True)
    Fm.sage_mean_aggregate(xr, gi).backward(go)
    xr = x.detach().requires_grad_(True)
    out, _ = Fm.gat_message(xr, att[0].view(1, 10, 35), att[1].view(1, 10, 35), None, gi, 10, 35)
    out.backward(go)


for _ in range(2):
    once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
