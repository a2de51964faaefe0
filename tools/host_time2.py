"""Host-side time of one training step (is the step launch bound? what do the round-2 host features cost?):
fusion off / on, and a cProfile of the fused step."""
import cProfile, pstats, sys, time, io
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
import ref_trunks
from m_gat_graphsage_b200 import nn as mnn
from m_gat_graphsage_b200.accel import use_mgs_linear

dev = torch.device("cuda:0")
torch.manual_seed(42)
model = ref_trunks.Model1Trunk(mnn).to(dev).train()
use_mgs_linear(model)
opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
batches = bench.make_batches(dev, 0, 6)


def measure(label):
    for i in range(12):
        bench.drop_index_cache(batches[i % 6]); bench.train_step(model, opt, batches[i % 6])
    torch.cuda.synchronize()
    host = []
    for i in range(30):
        b = batches[i % 6]
        bench.drop_index_cache(b)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bench.train_step(model, opt, b)
        host.append((time.perf_counter() - t0) * 1e3)
    host.sort()
    print(f"{label}: host ms per step (empty queue): median {host[15]:.3f}, min {host[0]:.3f}", flush=True)


mnn.set_activation_fusion(False)
measure("fusion off")
mnn.set_activation_fusion(True)
measure("fusion on ")
pr = cProfile.Profile()
torch.cuda.synchronize()
pr.enable()
for i in range(10):
    bench.drop_index_cache(batches[i % 6]); bench.train_step(model, opt, batches[i % 6])
pr.disable()
torch.cuda.synchronize()
st = io.StringIO()
pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(28)
print(st.getvalue()[:6000])
