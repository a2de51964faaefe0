"""Ablation timing of the CTA-pair TMA GEMM (MGS_TMA_DEBUG bits: 4 no MMA, 8 no conversion, 32 one commit per K block)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from m_gat_graphsage_b200 import functional as Fm
dev = torch.device("cuda:0")
M = 130512
def timed(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps
os.environ["MGS_TC_TMA"] = "2"
for (k, k2, n) in ((350, 350, 350), (350, 0, 128)):
    x = Fm.rows(M, k, dev); x.normal_()
    w = torch.randn(n, k, device=dev)
    x2 = w2 = None
    if k2:
        x2 = Fm.rows(M, k2, dev); x2.normal_(); w2 = torch.randn(n, k2, device=dev)
    for pair in ("0", "1"):
        os.environ["MGS_TMA_2CTA"] = pair
        print(f"pair={pair} [{M},{k}+{k2}]->{n}: ", end="", flush=True)
        for dbg in (0, 32, 4, 8, 12, 36, 44):
            os.environ["MGS_TMA_DEBUG"] = str(dbg)
            print(f"dbg{dbg}: {timed(lambda: Fm.linear_forward_raw(x, w, None, x2, w2)):.3f} | ", end="", flush=True)
        os.environ.pop("MGS_TMA_DEBUG")
        print(flush=True)
    if n == 350:
        os.environ["MGS_TMA_2CTA"] = "1"
        a = Fm.linear_forward_raw(x, w, None, x2, w2).clone()
        os.environ["MGS_TMA_DEBUG"] = "32"
        b = Fm.linear_forward_raw(x, w, None, x2, w2).clone()
        os.environ.pop("MGS_TMA_DEBUG")
        print("one-commit variant bit-equal:", torch.equal(a, b), flush=True)
