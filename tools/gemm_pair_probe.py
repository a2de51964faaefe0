"""CTA-pair TMA GEMM (tcgen05 cta_group::2, csrc/tc_tma.cuh gemm_tma2_kernel) against the one-CTA TMA kernel: bit equality
and time on the model1 / stress shapes.     python tools/gemm_pair_probe.py   (GPU box; wrap in `timeout`)"""
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


def case(name, m, k, k2, n, bias=False, relu=False):
    g = torch.Generator(device=dev).manual_seed(m + k + n)
    x = Fm.rows(m, k, dev); x.normal_(generator=g)
    w = torch.randn(n, k, device=dev, generator=g)
    b = torch.randn(n, device=dev, generator=g) if bias else None
    x2 = w2 = None
    if k2:
        x2 = Fm.rows(m, k2, dev); x2.normal_(generator=g)
        w2 = torch.randn(n, k2, device=dev, generator=g)
    flops = 2.0 * m * (k + k2) * n
    out = {}
    line = []
    for label, env in (("one-cta", {"MGS_TMA_2CTA": "0", "MGS_TC_TMA": "2"}), ("pair", {"MGS_TMA_2CTA": "1", "MGS_TC_TMA": "2"})):
        os.environ.update(env)
        fn = lambda: Fm.linear_forward_raw(x, w, b, x2, w2, relu)
        out[label] = fn().clone()
        torch.cuda.synchronize()
        ms = timed(fn)
        line.append(f"{label}: {ms:.4f} ms {flops / ms / 1e9:.0f} TF/s")
    ref = x[:, :k].double() @ w.double().t()
    if k2:
        ref = ref + x2[:, :k2].double() @ w2.double().t()
    if b is not None:
        ref = ref + b.double()
    if relu:
        ref = ref.clamp_min(0)
    same = torch.equal(out["one-cta"], out["pair"])
    err = float((out["pair"].double() - ref).abs().max() / ref.abs().max())
    print(f"{name:34s} " + " | ".join(line) + f" | bit-equal {same} | vs fp64 {err:.2e}", flush=True)


M = 130512
case("tiny     [300,64]->176", 300, 64, 0, 176)
case("odd rows [1000,350]->350 bias", 1000, 350, 0, 350, bias=True, relu=True)
case("SAGE fwd  [130k,350+350]->350", M, 350, 350, 350, bias=True)
case("SAGE dgrad [130k,350]->700", M, 350, 0, 700)
case("single   [130k,700]->350", M, 700, 0, 350)
case("stress   [130k,256+256]->256", M, 256, 256, 256)
case("fc_g1 dg [4096,1500]->700", 4096, 1500, 0, 700)
case("n=128    [130k,350]->128", M, 350, 0, 128)

# the converter-bound BN = 128 kernels, many launches (a rare hang of the one-CTA kernel was found here: tc_tma.cuh)
x = Fm.rows(M, 350, dev); x.normal_()
w = torch.randn(128, 350, device=dev)
for label, env in (("one-cta", "0"), ("pair", "1")):
    os.environ["MGS_TMA_2CTA"] = env
    for i in range(150):
        Fm.linear_forward_raw(x, w)
    torch.cuda.synchronize()
    print(f"BN=128 {label}: 150 launches done", flush=True)
