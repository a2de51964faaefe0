"""Host-side (launch) time of one training step vs its GPU time (is the step launch bound?)."""
import sys, time
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
opt = bench.make_adam(model.parameters(), lr=1e-4)
mnn.set_activation_fusion(True)
batches = bench.make_batches(dev, 0, 6)
for i in range(12):
    bench.drop_index_cache(batches[i % 6]); bench.train_step(model, opt, batches[i % 6])
torch.cuda.synchronize()
host = []
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(30):
    b = batches[i % 6]
    bench.drop_index_cache(b)
    t0 = time.perf_counter()
    bench.train_step(model, opt, b)
    host.append((time.perf_counter() - t0) * 1e3)
    if i % 3 == 2:
        torch.cuda.synchronize()          # let the host run ahead at most 3 steps, then measure unqueued launches
e.record(); torch.cuda.synchronize()
host.sort()
print(f"host ms per step: median {host[15]:.3f}, min {host[0]:.3f}; GPU ms per step {s.elapsed_time(e) / 30:.3f}")
