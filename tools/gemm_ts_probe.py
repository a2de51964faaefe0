"""K4 TMA kernel, activation operand from tensor memory (TS) against shared memory (SS): bit identity, then time.
    timeout -s KILL 120 python tools/gemm_ts_probe.py           (GPU box)"""
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


def run(env, fn):
    for k_, v in env.items():
        os.environ[k_] = v
    try:
        return fn()
    finally:
        for k_ in env:
            os.environ.pop(k_)


def case(name, m, k, k2, n, bn, time_it=True):
    torch.manual_seed(0)
    x = Fm.rows(m, k, dev); x.normal_()
    w = torch.randn(n, k, device=dev)
    b = torch.randn(n, device=dev)
    x2 = w2 = None
    if k2:
        x2 = Fm.rows(m, k2, dev); x2.normal_()
        w2 = torch.randn(n, k2, device=dev)
    f = lambda: Fm.linear_forward_raw(x, w, b, x2, w2)
    base = {"MGS_TC_TMA": "2", "MGS_TMA_BN": str(bn)}
    y_ss = run({**base, "MGS_TMA_TS": "0"}, f).clone()
    print(f"{name:32s} bn{bn} ss ok", end=" ", flush=True)
    y_ts = run({**base, "MGS_TMA_TS": "1"}, f).clone()
    torch.cuda.synchronize()
    ref = (x[:, :k].double() @ w.double().t() + (x2[:, :k2].double() @ w2.double().t() if k2 else 0) + b.double())
    print(f"| ts==ss {bool(torch.equal(y_ts, y_ss))} maxdiff {float((y_ts - y_ss).abs().max()):.3e} "
          f"err_ts {float((y_ts.double() - ref).abs().max() / ref.abs().max()):.2e} "
          f"err_ss {float((y_ss.double() - ref).abs().max() / ref.abs().max()):.2e}", end=" ", flush=True)
    if time_it:
        flops = 2.0 * m * (k + k2) * n
        t_ss = run({**base, "MGS_TMA_TS": "0"}, lambda: timed(f))
        t_ts = run({**base, "MGS_TMA_TS": "1"}, lambda: timed(f))
        print(f"| ss {t_ss:.4f} ms {flops / t_ss / 1e9:.0f} TF/s | ts {t_ts:.4f} ms {flops / t_ts / 1e9:.0f} TF/s", end="")
    print(flush=True)


case("tiny [300,40]->24", 300, 40, 0, 24, 176, False)
case("small [1000,350+350]->350", 1000, 350, 350, 350, 176, False)
case("small [1000,350]->350 bn128", 1000, 350, 0, 350, 128, False)
M = 130512
case("SAGE fwd [130k,350+350]->350", M, 350, 350, 350, 176)
case("SAGE fwd [130k,350+350]->350", M, 350, 350, 350, 128)
case("SAGE dgrad [130k,350]->700", M, 350, 0, 700, 176)
case("single [130k,700]->350", M, 700, 0, 350, 176)
case("stress [130k,256+256]->256", M, 256, 256, 256, 128)
case("fc_g1 [4096,700]->1500", 4096, 700, 0, 1500, 176)
case("fc_g2 [4096,1500]->128", 4096, 1500, 0, 128, 128)
