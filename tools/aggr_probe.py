"""Micro-benchmark of the aggregation kernels at the model1 shapes (profiling aid, not a bench value)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from m_gat_graphsage_b200 import functional as Fm
from m_gat_graphsage_b200.graph import build_graph_index
from m_gat_graphsage_b200.synth import synth_batch

which = sys.argv[1] if len(sys.argv) > 1 else "sage_fwd"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
b = synth_batch(4096, 42, device=dev)
N, E = b.x.size(0), b.edge_index.size(1)
gi = build_graph_index(b.edge_index, N)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(N, 350, device=dev, generator=g)
go = torch.randn(N, 350, device=dev, generator=g)
flush = torch.empty(64 * 1024 * 1024, device=dev)
att = torch.randn(2, 350, device=dev, generator=g)

def run():
    if which == "sage_fwd":
        return Fm.sage_mean_aggregate(x, gi)
    if which == "gat":
        xr = x.clone().requires_grad_(True)
        out, _ = Fm.gat_message(xr, att[0].view(1, 10, 35), att[1].view(1, 10, 35), None, gi, 10, 35)
        out.backward(go)
        return out

for _ in range(3):
    run()
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); run(); e.record(); torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
print(f"{which}: flushed-L2 median {sorted(ts)[len(ts)//2]:.4f} ms (N={N}, E={E})")
ts = []
for _ in range(reps):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); run(); e.record(); torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
print(f"{which}: back-to-back median {sorted(ts)[len(ts)//2]:.4f} ms")
if which == "sage_fwd":
    pre = torch.randn(N, 350, device=dev, generator=g)
    ts = []
    for _ in range(reps):
        xin = torch.relu(pre)                      # producer kernel right before, like in the model
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); Fm.sage_mean_aggregate(xin, gi); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    print(f"{which}: after-relu median {sorted(ts)[len(ts)//2]:.4f} ms")
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); y = x.clone(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    print(f"clone [N,350] (183 MB read + 183 MB write): median {sorted(ts)[len(ts)//2]:.4f} ms")
