"""Readout-MLP GEMMs (M = 4096 rows): kernel / tile / split variants.   python tools/mlp_gemm_probe.py   (GPU box)"""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from m_gat_graphsage_b200 import functional as Fm
sys.path.insert(0, str(Path(__file__).resolve().parent))
from wgrad_probe import timed

dev = torch.device("cuda:0")
VARIANTS = [("default", {})]
for bn in ("128", "176"):
    for sp in ("1", "2", "3", "4"):
        for pair in ("1", "0"):
            VARIANTS.append((f"tma bn{bn} s{sp} pair{pair}", {"MGS_TC_TMA": "2", "MGS_TMA_BN": bn, "MGS_TMA_SPLITS": sp, "MGS_TMA_2CTA": pair}))


def case(name, m, k, n, dgrad):
    x = Fm.rows(m, k, dev); x.normal_()
    w = torch.randn(n, k, device=dev) if not dgrad else torch.randn(k, n, device=dev)
    fn = (lambda: Fm.linear_dgrad_raw(x, w)) if dgrad else (lambda: Fm.linear_forward_raw(x, w))
    res = []
    for label, env in VARIANTS:
        os.environ.update(env)
        try:
            res.append((timed(fn, 30), label))
        except Exception as ex:  # noqa: BLE001
            res.append((9e9, label + " " + type(ex).__name__))
        for k_ in env:
            os.environ.pop(k_)
    base = res[0][0]
    res.sort()
    print(f"{name:28s} default {base:.4f} | best: " + " | ".join(f"{l} {t:.4f}" for t, l in res[:4]), flush=True)


case("fc_g1 fwd [4096,700]->1500", 4096, 700, 1500, False)
case("fc_g1 dgrad [4096,1500]->700", 4096, 1500, 700, True)
case("fc_g2 fwd [4096,1500]->128", 4096, 1500, 128, False)
case("fc_g2 dgrad [4096,128]->1500", 4096, 128, 1500, True)
