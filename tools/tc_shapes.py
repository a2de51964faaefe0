"""Run the packed tcgen05 GEMM on a few shapes, one process each (debug aid)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from m_gat_graphsage_b200 import functional as Fm
M, K, N = (int(v) for v in sys.argv[1:4])
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(M, K, device=dev, generator=g)
w = torch.randn(N, K, device=dev, generator=g) / K ** 0.5
out = Fm.linear_forward_raw(x, w, None)
torch.cuda.synchronize()
ref = (x.double() @ w.double().t())
print(M, K, N, "max rel err", float((out.double() - ref).abs().max() / ref.abs().max()))
