"""Time the K4 kernels on the model1 shapes: TMA-fed kernel at BN = 128 / 176 / 256 against the cp.async kernel.
    python tools/gemm_probe2.py           (GPU box)"""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from m_gat_graphsage_b200 import functional as Fm

dev = torch.device("cuda:0")
M = 130512


def timed(fn, reps=20):
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


def case(name, m, k, k2, n):
    x = Fm.rows(m, k, dev); x.normal_()
    w = torch.randn(n, k, device=dev)
    x2 = w2 = None
    if k2:
        x2 = Fm.rows(m, k2, dev); x2.normal_()
        w2 = torch.randn(n, k2, device=dev)
    flops = 2.0 * m * (k + k2) * n
    res = []
    for label, env in [("cp.async", {"MGS_TC_TMA": "0"}), ("tma bn128", {"MGS_TMA_BN": "128"}), ("tma bn176", {"MGS_TMA_BN": "176"}),
                       ("tma bn256", {"MGS_TMA_BN": "256"})] + [(f"tma bn256 split{s}", {"MGS_TMA_BN": "256", "MGS_TMA_SPLITS": str(s)})
                                                                 for s in ((2, 4) if m <= 8192 else ())]:
        for k_, v in env.items():
            os.environ[k_] = v
        try:
            ms = timed(lambda: Fm.linear_forward_raw(x, w, None, x2, w2))
            res.append(f"{label}: {ms:.4f} ms {flops / ms / 1e9:.0f} TF/s")
        except Exception as ex:  # noqa: BLE001
            res.append(f"{label}: {type(ex).__name__}")
        for k_ in env:
            os.environ.pop(k_)
    print(f"{name:34s} " + " | ".join(res), flush=True)


case("SAGE fwd  [130k,350+350]->350", M, 350, 350, 350)
case("SAGE dgrad [130k,350]->700", M, 350, 0, 700)
case("single   [130k,700]->350", M, 700, 0, 350)
case("wide     [130k,700]->1500", M, 700, 0, 1500)
case("stress   [130k,256+256]->256", M, 256, 256, 256)
case("fc_g1    [4096,700]->1500", 4096, 700, 0, 1500)
case("fc_g2    [4096,1500]->128", 4096, 1500, 0, 128)
case("fc_g2 dg [4096,128]->1500", 4096, 128, 0, 1500)
case("fc_g1 dg [4096,1500]->700", 4096, 1500, 0, 700)
