"""ncu target: one SAGE-projection-shaped GEMM on the CTA-pair TMA kernel and one on the one-CTA TMA kernel.
   ncu --set full --profile-from-start off -o out python tools/ncu_gemm_pair.py"""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from m_gat_graphsage_b200 import functional as Fm

dev = torch.device("cuda:0")
M = 130512
x = Fm.rows(M, 350, dev); x.normal_().relu_()
x2 = Fm.rows(M, 350, dev); x2.normal_().relu_()
w, w2 = torch.randn(350, 350, device=dev), torch.randn(350, 350, device=dev)
for _ in range(2):
    Fm.linear_forward_raw(x, w, None, x2, w2)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
os.environ["MGS_TMA_2CTA"] = "1"
Fm.linear_forward_raw(x, w, None, x2, w2)
os.environ["MGS_TMA_2CTA"] = "0"
Fm.linear_forward_raw(x, w, None, x2, w2)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
