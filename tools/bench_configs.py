"""Throughput of the other BASELINE.json configs on one B200 (informational; bench.py prints the contract line).

    configs[1] inference        model1 trunk, batch 4096, eval / no_grad
    configs[3] atom importance  model1 trunk, batch 4096, d pred / d x only, per-atom L2 norm
    configs[4] stress           GATConv(256 -> 8 x 32) + SAGEConv(256, 256), 94-atom molecules, batch 16384, training step
"""
import json
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.nn.functional as F
import ref_trunks
from m_gat_graphsage_b200 import nn as mnn
from m_gat_graphsage_b200.accel import use_mgs_linear
from m_gat_graphsage_b200.synth import batch_seed, synth_batch

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10


def drop(b):
    for t in (b.edge_index, b.batch):
        for a in ("_mgs_graph", "_mgs_gptr"):
            if hasattr(t, a):
                delattr(t, a)


def timed(fn, batches, warm=4):
    for i in range(warm):
        drop(batches[i % len(batches)])
        fn(batches[i % len(batches)])
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(reps):
        drop(batches[i % len(batches)])
        fn(batches[i % len(batches)])
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


out = {}
model = ref_trunks.build_trunk("model1", mnn).to(dev)
use_mgs_linear(model)
batches = [synth_batch(4096, batch_seed(42, 0, i), device=dev) for i in range(4)]
model.eval()
with torch.no_grad():
    ms = timed(lambda b: model(b), batches)
out["configs[1] inference (model1, B=4096)"] = {"ms_per_batch": round(ms, 4), "molecules_per_s": round(4096 / ms * 1e3)}
ms = timed(lambda b: ref_trunks.atom_importance(model, b), batches)
out["configs[3] atom importance (model1, B=4096)"] = {"ms_per_batch": round(ms, 4), "molecules_per_s": round(4096 / ms * 1e3),
                                                       "atoms_per_s": round(batches[0].x.size(0) / ms * 1e3)}
del model, batches
torch.cuda.empty_cache()
stress = ref_trunks.build_trunk("stress", mnn).to(dev).train()
use_mgs_linear(stress)
opt = torch.optim.Adam(stress.parameters(), lr=1e-4, fused=True)
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
sb = []
for i in range(2):
    b = synth_batch(B, batch_seed(42, 0, 100 + i), device=dev, fixed_atoms=94)
    fin = stress.conv1.in_channels if hasattr(stress, "conv1") else 256
    if b.x.size(1) != fin:                      # the stress trunk takes 256 input features
        g = torch.Generator(device=dev).manual_seed(i)
        b.x = torch.randn(b.x.size(0), fin, device=dev, generator=g)
    sb.append(b)


def step(b):
    opt.zero_grad(set_to_none=True)
    loss = F.mse_loss(stress(b).view(-1), b.y)
    loss.backward()
    opt.step()


ms = timed(step, sb, warm=3)
out[f"configs[4] stress training step (B={B}, {sb[0].x.size(0)} atoms, {sb[0].edge_index.size(1)} edges)"] = {
    "ms_per_step": round(ms, 3), "molecules_per_s": round(B / ms * 1e3), "atoms_per_s": round(sb[0].x.size(0) / ms * 1e3),
    "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
print(json.dumps(out, indent=1))
