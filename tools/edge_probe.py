"""Time mgs_gat_bwd_edge (model1 shape, batch 4096): FMA path (edge_fma.cuh) vs tensor-core path (edge_mma.cuh) vs the staged kernel (edge.cuh)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from m_gat_graphsage_b200 import functional as Fm
from m_gat_graphsage_b200.graph import build_graph_index
from m_gat_graphsage_b200.synth import synth_batch
from torch.profiler import ProfilerActivity, profile

dev = torch.device("cuda:0")
for (H, C, fixed) in ((10, 35, None), (8, 32, 94)):
    b = synth_batch(4096, 42, device=dev, fixed_atoms=fixed)
    N = b.x.size(0)
    gi = build_graph_index(b.edge_index, N)
    gen = torch.Generator(device=dev).manual_seed(0)
    xh = Fm.rows(N, H * C, dev); xh.normal_(generator=gen)
    go = Fm.rows(N, H * C, dev); go.normal_(generator=gen)
    a_s, a_d = torch.randn(N, H, device=dev, generator=gen), torch.randn(N, H, device=dev, generator=gen)
    res = {}
    for label, env, fma in (("fma d3", "1", "3"), ("fma d2", "1", "2"), ("fma d4", "1", "4"), ("mma", "1", "0"), ("staged", "0", "0")):
        os.environ["MGS_EDGE_MMA"] = env
        os.environ["MGS_EDGE_FMA"] = fma
        x = xh.detach().requires_grad_(True)
        out, _ = Fm.gat_message(x, a_s.requires_grad_(True), a_d.requires_grad_(True), None, gi, H, C, scores=True)
        for _ in range(3):
            out.backward(go, retain_graph=True)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(10):
                out.backward(go, retain_graph=True)
            torch.cuda.synchronize()
        ks = {e.key[:70]: round(e.device_time_total / 10, 1) for e in prof.key_averages() if "edge" in e.key or "softmax" in e.key}
        res[label] = (ks, x.grad.clone())
        print(f"H={H} C={C} N={N} {label}: {ks}", flush=True)
    for k in ("mma", "fma d3", "fma d2", "fma d4"):
        err = float((res[k][1] - res['staged'][1]).abs().max()) / float(res['staged'][1].abs().max())
        print(f"   d xh agreement {k} vs staged: {err:.2e}")
