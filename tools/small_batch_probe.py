"""The reference's own batch sizes (64: ablation/model1.py:109, 128: train.py:209): eager steps are bound by the
host's launch rate; graphed.GraphedStep replays one CUDA graph per step on padded buffers."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.nn.functional as F
import ref_trunks
from m_gat_graphsage_b200 import nn as mnn
from m_gat_graphsage_b200.accel import use_mgs_linear
from m_gat_graphsage_b200.graphed import GraphedStep
from m_gat_graphsage_b200.synth import batch_seed, synth_batch

dev = torch.device("cuda:0")
loss_fn = lambda out, y: F.mse_loss(out.view(-1), y)
for B in (64, 128, 512):
    batches = [synth_batch(B, batch_seed(42, 0, 7000 + i), device=dev) for i in range(40)]
    n_max = max(b.x.size(0) for b in batches)
    e_max = max(b.edge_index.size(1) for b in batches)
    res = []
    for kind in ("eager", "graphed"):
        torch.manual_seed(42)
        model = ref_trunks.build_trunk("model1", mnn).to(dev).train()
        use_mgs_linear(model)
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=True, fused=True)
        if kind == "eager":
            def step(b):
                opt.zero_grad(set_to_none=True)
                loss = loss_fn(model(b), b.y)
                loss.backward()
                opt.step()
                return loss
        else:
            step = GraphedStep(model, B, int(n_max * 1.08) + 8, int(e_max * 1.08) + 8, optimizer=opt, loss_fn=loss_fn)
        for b in batches[:5]:
            step(b)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            for b in batches:
                step(b)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / (3 * len(batches))
        extra = f" (replays {step.replays}, eager fallbacks {step.eager})" if kind == "graphed" else ""
        res.append(f"{kind}: {dt * 1e3:.3f} ms/step {B / dt:9.0f} mol/s{extra}")
    print(f"B={B:4d} (<= {n_max} atoms): " + " | ".join(res))

# a reference-style epoch at the script's batch size: list of Data -> DataLoader (collated once, gathered per batch on
# the GPU) -> graphed step
from m_gat_graphsage_b200.data import DataLoader
B = 128
mols = []
for i in range(16):
    b = synth_batch(512, batch_seed(42, 0, 9000 + i))
    ms_ = b.to_data_list()
    for k, m in enumerate(ms_):
        m.y = b.y[k]
    mols += ms_
loader = DataLoader(mols, batch_size=B, shuffle=True, device=dev, drop_last=True)
torch.manual_seed(42)
model = ref_trunks.build_trunk("model1", mnn).to(dev).train()
use_mgs_linear(model)
opt = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=True, fused=True)
step = GraphedStep(model, B, 5200, 11200, optimizer=opt, loss_fn=loss_fn)
for mode in ("eager", "graphed"):
    for epoch in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 0
        for batch in loader:
            if mode == "graphed":
                step(batch)
            else:
                opt.zero_grad(set_to_none=True)
                loss_fn(model(batch), batch.y).backward()
                opt.step()
            n += batch.num_graphs
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"epoch of {n} molecules, batch {B}, device-resident DataLoader, {mode} step: {n / dt:9.0f} molecules/s "
          f"({dt / (n / B) * 1e3:.3f} ms per step)" + (f" replays {step.replays} eager {step.eager}" if mode == "graphed" else ""))
