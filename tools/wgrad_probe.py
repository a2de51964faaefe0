"""Weight-gradient GEMM: TMA-fed kernel (tc_wgrad.cuh) against the cp.async kernel (tc_linear.cuh) -- error vs fp64 and time.
    python tools/wgrad_probe.py           (GPU box)"""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from m_gat_graphsage_b200 import functional as Fm

dev = torch.device("cuda:0")


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


def case(name, m, nout, k, check=True):
    torch.manual_seed(0)
    g = Fm.rows(m, nout, dev); g.normal_()
    x = Fm.rows(m, k, dev); x.normal_()
    ref = None
    if check:
        ref = (g.double().t() @ x.double())
    flops = 2.0 * m * nout * k
    res = []
    variants = [("cp.async", {"MGS_WGRAD_TMA": "0"}), ("tma", {}), ("tma 1 wave", {"MGS_WGRAD_WAVES": "1"}),
                ("tma 2 waves", {"MGS_WGRAD_WAVES": "2"}), ("tma 3 waves", {"MGS_WGRAD_WAVES": "3"})]
    for label, env in variants:
        for k_, v in env.items():
            os.environ[k_] = v
        try:
            out = Fm.linear_wgrad_raw(g, x)
            torch.cuda.synchronize()
            err = float((out.double() - ref).abs().max() / ref.abs().max()) if check else float("nan")
            ms = timed(lambda: Fm.linear_wgrad_raw(g, x))
            res.append(f"{label}: {ms:.4f} ms {flops / ms / 1e9:.0f} TF/s err {err:.1e}")
        except Exception as ex:  # noqa: BLE001
            res.append(f"{label}: {type(ex).__name__} {ex}")
        for k_ in env:
            os.environ.pop(k_)
    print(f"{name:30s} " + " | ".join(res), flush=True)


if __name__ == "__main__":
    case("tiny   [2048] 64 x 64", 2048, 64, 64)
    case("small  [4096] 128 x 176", 4096, 128, 176)
    case("SAGE   [130k] 350 x 350", 130512, 350, 350)
    case("SAGE+b [130k] 350 x 351", 130512, 350, 351)
    case("stress [130k] 256 x 256", 130512, 256, 256)
    case("fc_g1  [4096] 1500 x 700", 4096, 1500, 700)
    case("fc_g2  [4096] 128 x 1500", 4096, 128, 1500)
    case("ragged [5000] 100 x 90", 5000, 100, 90)
    case("proj   [130k] 350 x 35", 130512, 350, 35)
    case("proj^T [130k] 35 x 350", 130512, 35, 350)
