import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from m_gat_graphsage_b200 import functional as Fm
dev = torch.device("cuda:0")
def rel(a, c): return float((a.double() - c.double()).abs().max()) / max(float(c.double().abs().max()), 1e-30)
g = torch.Generator(device=dev).manual_seed(0)
for (M, K, N) in [(9500, 350, 350), (3000, 350, 350), (9472, 350, 350), (9500, 35, 350), (300, 700, 1500), (130512, 350, 350)]:
    x = torch.randn(M, K, device=dev, generator=g); w = torch.randn(N, K, device=dev, generator=g) / K ** 0.5
    b = torch.randn(N, device=dev, generator=g); go = torch.randn(M, N, device=dev, generator=g)
    res = {}
    for mode in ("tc", "ffma"):
        if mode == "ffma": os.environ["MGS_DISABLE_TC"] = "1"
        else: os.environ.pop("MGS_DISABLE_TC", None)
        res[mode] = (Fm.linear_forward_raw(x, w, b), Fm.linear_dgrad_raw(go, w), Fm.linear_wgrad_raw(go, x))
    ref = (torch.nn.functional.linear(x.double(), w.double(), b.double()), go.double() @ w.double(), go.double().t() @ x.double())
    print(f"M={M} K={K} N={N}: " + "  ".join(f"{nm}: tc {rel(res['tc'][i], ref[i]):.1e} ffma {rel(res['ffma'][i], ref[i]):.1e}" for i, nm in enumerate(("fwd", "dgrad", "wgrad"))))
    # where is the wgrad error?
    d = (res['tc'][2].double() - ref[2]).abs()
    if rel(res['tc'][2], ref[2]) > 1e-4:
        idx = torch.nonzero(d > 0.5 * d.max())
        print("   worst wgrad entries (row, col):", idx[:8].tolist(), "max", float(d.max()))
