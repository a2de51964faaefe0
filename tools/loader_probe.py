"""Row a1: what DataLoader collation costs next to the train step.  Builds a dataset of `nb` x 4096 synthetic
molecules as a Python list of Data (what train.py:169-193 builds), then times (1) the per-batch Python collation,
(2) the flat gather on the host, (3) the flat gather on the GPU, and (4) a whole reference-style training epoch
`for batch in loader: batch.to(device); step` through the device-resident loader."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench, ref_trunks
from m_gat_graphsage_b200 import nn as mnn
from m_gat_graphsage_b200.accel import use_mgs_linear
from m_gat_graphsage_b200.data import DataLoader
from m_gat_graphsage_b200.synth import batch_seed, synth_batch

dev = torch.device("cuda:0")
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
B = 4096
mols = []
for i in range(nb):
    b = synth_batch(B, batch_seed(42, 0, 500 + i))
    ms = b.to_data_list()
    for k, m in enumerate(ms):
        m.y = b.y[k]
    mols += ms
print(f"dataset: {len(mols)} molecules")


def rate(loader, batches, sync):
    it = iter(loader)
    next(it)
    if sync:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    for _ in range(batches):
        try:
            next(it)
        except StopIteration:
            break
        n += 1
    if sync:
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return dt / n * 1e3, B * n / dt


ms, r = rate(DataLoader(mols, batch_size=B, shuffle=True, fast=False), 3, False)
print(f"python collation      : {ms:8.2f} ms per batch, {r:10.0f} molecules/s")
t0 = time.perf_counter()
host = DataLoader(mols, batch_size=B, shuffle=True)
iter(host).__next__()
print(f"one-time flat collation (host): {time.perf_counter() - t0:.2f} s")
ms, r = rate(host, nb - 1, False)
print(f"flat gather, host     : {ms:8.2f} ms per batch, {r:10.0f} molecules/s")
t0 = time.perf_counter()
gpu = DataLoader(mols, batch_size=B, shuffle=True, device=dev)
iter(gpu).__next__()
torch.cuda.synchronize()
print(f"one-time flat collation (+upload): {time.perf_counter() - t0:.2f} s")
for _ in range(3):
    ms, r = rate(gpu, nb - 1, True)
print(f"flat gather, GPU      : {ms:8.2f} ms per batch, {r:10.0f} molecules/s")

torch.manual_seed(42)
model = ref_trunks.Model1Trunk(mnn).to(dev).train()
use_mgs_linear(model)
opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
for loader, name in ((gpu, "device-resident loader"), (host, "host flat loader + .to(device)")):
    for epoch in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 0
        for batch in loader:
            batch = batch.to(dev)
            bench.train_step(model, opt, batch)
            n += batch.num_graphs
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"training epoch, {name}: {n / dt:10.0f} molecules/s ({dt / (n / B) * 1e3:.2f} ms per step)")
