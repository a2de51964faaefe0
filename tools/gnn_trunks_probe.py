"""Training-step throughput of the reference's other GNN trunks (gnn/gcn.py, gnn/gat-gcn.py, gnn/gin.py,
gnn/gat.py, gnn/graphsage.py) at 4096 molecules per batch, next to the north-star model1 trunk."""
import sys
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
B = 4096
batches = [synth_batch(B, batch_seed(42, 0, 900 + i), device=dev) for i in range(4)]
for name in ("model1", "gat-gcn", "gcn", "gin", "gat", "graphsage"):
    model = ref_trunks.build_trunk(name, mnn).to(dev).train()
    use_mgs_linear(model)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)

    def step(b):
        opt.zero_grad(set_to_none=True)
        F.mse_loss(model(b).view(-1), b.y).backward()
        opt.step()

    for i in range(6):
        step(batches[i % 4])
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(20):
        step(batches[i % 4])
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    line = f"{name:10s} eager {ms:7.3f} ms/step {B / ms * 1e3:10.0f} molecules/s"
    if name != "gin":          # BatchNorm statistics would see the padding atoms: not valid under padding
        model = ref_trunks.build_trunk(name, mnn).to(dev).train()
        use_mgs_linear(model)
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True, capturable=True)
        n_cap = int(max(b.x.size(0) for b in batches) * 1.02) + 8
        e_cap = int(max(b.edge_index.size(1) for b in batches) * 1.02) + 8
        gs = GraphedStep(model, B, n_cap, e_cap, optimizer=opt, loss_fn=lambda o, y: F.mse_loss(o.view(-1), y))
        for i in range(4):
            gs(batches[i % 4])
        torch.cuda.synchronize()
        s.record()
        for i in range(20):
            gs(batches[i % 4])
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        line += f" | CUDA graph {ms:7.3f} ms/step {B / ms * 1e3:10.0f} molecules/s (replays {gs.replays}, eager {gs.eager})"
    print(line)
