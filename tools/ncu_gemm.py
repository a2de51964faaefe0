"""ncu target: one SAGE-projection-shaped GEMM on the TMA-fed kernel and one on the cp.async kernel (profiling aid).
   ncu --set full --profile-from-start off -o out python tools/ncu_gemm.py [bn]"""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from m_gat_graphsage_b200 import functional as Fm

dev = torch.device("cuda:0")
M = 130512
if len(sys.argv) > 1:
    os.environ["MGS_TMA_BN"] = sys.argv[1]
x = Fm.rows(M, 350, dev); x.normal_().relu_()
x2 = Fm.rows(M, 350, dev); x2.normal_().relu_()
w, w2 = torch.randn(350, 350, device=dev), torch.randn(350, 350, device=dev)
for _ in range(2):
    Fm.linear_forward_raw(x, w, None, x2, w2)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
Fm.linear_forward_raw(x, w, None, x2, w2)
os.environ["MGS_TC_TMA"] = "0"
Fm.linear_forward_raw(x, w, None, x2, w2)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
