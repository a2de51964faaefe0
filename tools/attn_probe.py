"""K5 timing: streaming attention fwd / bwd at the batch sizes of the reference (128 molecules) up to 4096, global
and per-molecule; FFMA-issue roofline (83 FFMA per score fwd; 70 + 83 (bwd_q) + 70 + 96 (bwd_kv) backward)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from m_gat_graphsage_b200 import functional as Fm
from m_gat_graphsage_b200.graph import graph_ptr
from m_gat_graphsage_b200.synth import synth_batch

dev = torch.device("cuda:0")
d = 35
peak = 148 * 128 * 1.9e9        # FFMA lanes/s at ~1.9 GHz


def timeit(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for B in (128, 512, 2048, 4096):
    b = synth_batch(B, 7, device=dev)
    n = b.x.size(0)
    y = torch.randn(n, 3 * d, device=dev, requires_grad=True)
    g = torch.randn(n, d, device=dev)
    seg, gptr = b.batch.to(torch.int32), graph_ptr(b.batch, B)
    for name, args, scores in (("global", (None, None), float(n) * n),
                               ("per-molecule", (seg, gptr), float(((b.ptr[1:] - b.ptr[:-1]) ** 2).sum()))):
        reps = 3 if scores > 1e9 else 20
        with torch.no_grad():
            tf = timeit(lambda: Fm.stream_attention(y, d, d ** -0.5, *args), reps)
        out = Fm.stream_attention(y, d, d ** -0.5, *args)
        tb = timeit(lambda: torch.autograd.grad(out, y, g, retain_graph=True), reps)
        stock = ""
        if name == "global" and n <= 70000:
            from oracle import pyg_oracle as O
            yd = y.detach().clone().requires_grad_(True)
            fs = lambda: O.modified_gat_attention(yd[:, :d], yd[:, d:2 * d], yd[:, 2 * d:])
            with torch.no_grad():
                ts = timeit(fs, reps)
            tsb = float("nan")
            if n <= 40000:
                o2 = fs()
                tsb = timeit(lambda: torch.autograd.grad(o2, yd, g, retain_graph=True), reps)
                del o2
            stock = f"   stock PyTorch dense: fwd {ts:.3f} ms bwd {tsb:.3f} ms"
            torch.cuda.empty_cache()
        print(f"B={B:5d} N={n:7d} {name:12s} scores {scores:.3g}: fwd {tf:9.3f} ms ({scores * 83 / tf / 1e-3 / peak * 100:5.1f}% of FFMA issue)"
              f"  bwd {tb:9.3f} ms ({scores * 319 / tb / 1e-3 / peak * 100:5.1f}%){stock}")
