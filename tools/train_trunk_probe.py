"""The trunk the three reference scripts actually run (train.py:102-124: ModifiedGATLayer -> SAGEConv -> max pool
-> MLP), training step and inference, per batch size:
  stock   the script's own dense [N, N] attention (PyTorch ops on the GPU) + our SAGE / pool / MLP kernels
  k5      use_mgs_attention: streaming attention over the whole batch (the reference's semantics)
  k5-mol  the same under molecule_attention(batch): softmax restricted to each molecule"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.nn.functional as F
import ref_trunks
from m_gat_graphsage_b200 import nn as mnn
from m_gat_graphsage_b200.accel import use_mgs_linear
from m_gat_graphsage_b200.attention import molecule_attention, use_mgs_attention
from m_gat_graphsage_b200.synth import batch_seed, synth_batch

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


base = ref_trunks.build_trunk("train", mnn).to(dev)
use_mgs_linear(base)
for B in (128, 512, 2048, 4096):
    graphed_line = None
    b = synth_batch(B, batch_seed(42, 0, B), device=dev)
    n = b.x.size(0)
    row = [f"B={B:5d} N={n:7d}"]
    for kind in ("stock", "k5", "k5-mol"):
        if kind == "stock" and n > 40000:
            row.append("stock: [N,N] fp32 x3 does not fit / not attempted")
            continue
        model = ref_trunks.build_trunk("train", mnn).to(dev).train()
        model.load_state_dict(base.state_dict())
        use_mgs_linear(model)
        if kind != "stock":
            use_mgs_attention(model)
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)

        def step():
            opt.zero_grad(set_to_none=True)
            if kind == "k5-mol":
                with molecule_attention(b.batch):
                    out = model(b)
            else:
                out = model(b)
            F.mse_loss(out.view(-1), b.y).backward()
            opt.step()

        reps = 3 if (kind != "k5-mol" and n > 40000) else 20
        try:
            ms = timed(step, reps)
            torch.cuda.synchronize()
            peak = torch.cuda.max_memory_allocated() / 2**30
            row.append(f"{kind}: {ms:8.3f} ms/step {B / ms * 1e3:9.0f} mol/s (peak {peak:.1f} GB)")
        except torch.OutOfMemoryError:
            row.append(f"{kind}: out of memory")
        del model, opt
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
    # per-molecule attention under a CUDA graph (graphed.GraphedStep)
    from m_gat_graphsage_b200.graphed import GraphedStep
    model = ref_trunks.build_trunk("train", mnn).to(dev).train()
    model.load_state_dict(base.state_dict())
    use_mgs_linear(model); use_mgs_attention(model)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True, capturable=True)

    def fwd(m, d):
        with molecule_attention(d.batch):
            return m(d)

    gs = GraphedStep(model, B, int(n * 1.05) + 8, int(b.edge_index.size(1) * 1.05) + 8, optimizer=opt,
                     loss_fn=lambda o, y: F.mse_loss(o.view(-1), y), forward=fwd)
    ms = timed(lambda: gs(b), 20)
    row.append(f"k5-mol graphed: {ms:8.3f} ms/step {B / ms * 1e3:9.0f} mol/s")
    if n <= 20000:          # the reference's whole-batch softmax on the padded batch
        from m_gat_graphsage_b200.attention import padded_batch_attention
        model = ref_trunks.build_trunk("train", mnn).to(dev).train()
        model.load_state_dict(base.state_dict())
        use_mgs_linear(model); use_mgs_attention(model)
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True, capturable=True)

        def fwd_all(m, d):
            with padded_batch_attention(d.batch, B):
                return m(d)

        gs = GraphedStep(model, B, int(n * 1.05) + 8, int(b.edge_index.size(1) * 1.05) + 8, optimizer=opt,
                         loss_fn=lambda o, y: F.mse_loss(o.view(-1), y), forward=fwd_all)
        ms = timed(lambda: gs(b), 20)
        row.append(f"k5 whole-batch graphed: {ms:8.3f} ms/step {B / ms * 1e3:9.0f} mol/s")
    print(" | ".join(row))
