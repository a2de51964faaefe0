"""ncu target: training steps of the model1 trunk at the reference's batch size (64 molecules)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch, torch.nn.functional as F
import ref_trunks
from m_gat_graphsage_b200 import nn as mnn
from m_gat_graphsage_b200.accel import use_mgs_linear
from m_gat_graphsage_b200.synth import synth_batch
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model = ref_trunks.build_trunk("model1", mnn).to(dev).train()
use_mgs_linear(model)
opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
bs = [synth_batch(B, 50 + i, device=dev) for i in range(3)]
for i in range(4):
    if i == 3:
        torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
    b = bs[i % 3]
    opt.zero_grad(set_to_none=True)
    F.mse_loss(model(b).view(-1), b.y).backward()
    opt.step()
torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
