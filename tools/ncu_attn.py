"""ncu target: one K5 forward + backward (whole batch) at B = 1024 molecules."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from m_gat_graphsage_b200 import functional as Fm
from m_gat_graphsage_b200.synth import synth_batch
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
d = 35
n = synth_batch(B, 7).x.size(0)
y = torch.randn(n, 3 * d, device=dev, requires_grad=True)
out = Fm.stream_attention(y, d, d ** -0.5)
out.backward(torch.randn(n, d, device=dev))
torch.cuda.synchronize()
print("atoms", n)
