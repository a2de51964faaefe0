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
