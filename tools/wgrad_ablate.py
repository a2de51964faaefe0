import os, sys
sys.path.insert(0, "/root/repo")
import torch
from m_gat_graphsage_b200 import functional as Fm
sys.path.insert(0, "/root/repo/tools")
from wgrad_probe import timed
dev = torch.device("cuda:0")
for name, m, nout, k in [("SAGE 350x350", 130512, 350, 350), ("proj 350x35", 130512, 350, 35)]:
    g = Fm.rows(m, nout, dev); g.normal_()
    x = Fm.rows(m, k, dev); x.normal_()
    res = []
    for label, dbg in [("full", "0"), ("no MMA", "4"), ("no conversion", "8"), ("no MMA, no conversion", "12")]:
        os.environ["MGS_TMA_DEBUG"] = dbg
        os.environ["MGS_WGRAD_WAVES"] = "1"
        res.append(f"{label}: {timed(lambda: Fm.linear_wgrad_raw(g, x)):.4f}")
    print(name, " | ".join(res), flush=True)
