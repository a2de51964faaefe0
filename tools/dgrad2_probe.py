"""mgs_linear_dgrad2 (two data gradients as one GEMM, ReLU-mask epilogue) with and without the mask, against the forward
GEMM of the same shape.    python tools/dgrad2_probe.py    (GPU box)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from m_gat_graphsage_b200 import functional as Fm, _lib
from m_gat_graphsage_b200.functional import _ld, _workspace, device_guard, stream_ptr, rows, stream_row_words
sys.path.insert(0, str(Path(__file__).resolve().parent))
from wgrad_probe import timed

dev = torch.device("cuda:0")
M, O, F = 130512, 350, 350
lib = _lib.load()
g = rows(M, O, dev); g.normal_()
gh = rows(M, O, dev); gh.normal_()
wr = torch.randn(O, F, device=dev) / 18
wl = torch.randn(O, F, device=dev) / 18
gx = rows(M, F, dev)
words = stream_row_words(F)
bits = torch.randint(0, 2 ** 31, (M, words), device=dev, dtype=torch.int32)
ws = _workspace(lib.mgs_linear_dgrad2_workspace_bytes(M, O, O, F), dev)


cs = torch.empty(F, device=dev)


def run(with_bits, with_colsum=False):
    with device_guard(dev):
        rc = lib.mgs_linear_dgrad2(g.data_ptr(), _ld(g), O, wr.data_ptr(), _ld(wr), gh.data_ptr(), _ld(gh), O, wl.data_ptr(),
                                   _ld(wl), M, F, gx.data_ptr(), _ld(gx), bits.data_ptr() if with_bits else 0,
                                   words if with_bits else 0, 2, cs.data_ptr() if with_colsum else 0, ws.data_ptr(),
                                   ws.numel(), stream_ptr())
    _lib.check(rc, "mgs_linear_dgrad2")


run(False)
ref = g.double() @ wr.double() + gh.double() @ wl.double()
print("err without mask", float((gx.double() - ref).abs().max() / ref.abs().max()))
run(False, True)
torch.cuda.synchronize()
print("colsum err", float((cs.double() - gx.double().sum(0)).abs().max() / gx.double().sum(0).abs().max()))
print(f"dgrad2 with mask + column sums {timed(lambda: run(True, True)):.4f} ms")
print(f"dgrad2 without mask {timed(lambda: run(False)):.4f} ms, with mask {timed(lambda: run(True)):.4f} ms")
w2l, w2r = wl.t().contiguous(), wr.t().contiguous()
print(f"forward form, same shape {timed(lambda: Fm.linear_forward_raw(g, w2r, None, gh, w2l)):.4f} ms")
