"""ncu target: GATConv(35, 35, heads=10) projection forward + wgrad at the model1 batch (profiling aid)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from m_gat_graphsage_b200 import functional as Fm
from m_gat_graphsage_b200.synth import synth_batch

dev = torch.device("cuda:0")
b = synth_batch(4096, 42, device=dev)
N = b.x.size(0)
gen = torch.Generator(device=dev).manual_seed(0)
w = (torch.randn(350, 35, device=dev, generator=gen) * 0.1).requires_grad_(True)
att = (torch.randn(2, 10, 35, device=dev, generator=gen)).requires_grad_(True)
go = torch.randn(N, 350, device=dev, generator=gen)
ga = torch.randn(N, 10, device=dev, generator=gen)


def once():
    xh, a_s, a_d = Fm.gat_project(b.x, w, att[0], att[1], 10, 35)
    torch.autograd.backward([xh, a_s, a_d], [go, ga, ga])


for _ in range(3):
    once()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); once(); e.record(); torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
print("fwd+bwd ms:", sorted(ts)[5])
torch.cuda.cudart().cudaProfilerStart()
once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
