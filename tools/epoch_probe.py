"""Steady-state throughput over MANY distinct batch shapes (a slice of BASELINE configs[2]: every batch of an epoch has
its own atom / edge count, so the caching allocator sees new sizes all the time)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench, ref_trunks
from m_gat_graphsage_b200 import nn as mnn
from m_gat_graphsage_b200.accel import use_mgs_linear
from m_gat_graphsage_b200.synth import batch_seed, synth_batch

dev = torch.device("cuda:0")
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 80
torch.manual_seed(42)
model = ref_trunks.Model1Trunk(mnn).to(dev).train()
use_mgs_linear(model)
opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
batches = [synth_batch(4096, batch_seed(42, 0, 1000 + i), device=dev) for i in range(nb)]
print("atoms per batch: min", min(b.x.size(0) for b in batches), "max", max(b.x.size(0) for b in batches))
for b in batches[:8]:
    bench.train_step(model, opt, b)
torch.cuda.synchronize()
marks = [torch.cuda.Event(enable_timing=True) for _ in range(nb + 1)]
marks[0].record()
for i, b in enumerate(batches):
    bench.train_step(model, opt, b)
    marks[i + 1].record()
torch.cuda.synchronize()
ts = [marks[i].elapsed_time(marks[i + 1]) for i in range(nb)]
tot = marks[0].elapsed_time(marks[nb])
print(f"{nb} distinct batches: {4096 * nb / tot * 1e3:.0f} molecules/s, ms/step median {sorted(ts)[nb // 2]:.3f} max {max(ts):.3f}; "
      f"reserved {torch.cuda.memory_reserved() / 2**30:.2f} GB, allocated peak {torch.cuda.max_memory_allocated() / 2**30:.2f} GB")
