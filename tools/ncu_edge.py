"""ncu target: one GAT message backward at the model1 shape (profiling aid)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from m_gat_graphsage_b200 import functional as Fm
from m_gat_graphsage_b200.graph import build_graph_index
from m_gat_graphsage_b200.synth import synth_batch
dev = torch.device("cuda:0")
H, C = 10, 35
b = synth_batch(4096, 42, device=dev)
N = b.x.size(0)
gi = build_graph_index(b.edge_index, N)
gen = torch.Generator(device=dev).manual_seed(0)
xh = Fm.rows(N, H * C, dev); xh.normal_(generator=gen)
go = Fm.rows(N, H * C, dev); go.normal_(generator=gen)
a_s, a_d = torch.randn(N, H, device=dev, generator=gen), torch.randn(N, H, device=dev, generator=gen)
x = xh.detach().requires_grad_(True)
out, _ = Fm.gat_message(x, a_s.requires_grad_(True), a_d.requires_grad_(True), None, gi, H, C, scores=True)
out.backward(go, retain_graph=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out.backward(go, retain_graph=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
