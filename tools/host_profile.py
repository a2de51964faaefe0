"""cProfile of the host side of the training step (where do the ~1.9 ms of launch time go?)."""
import cProfile, pstats, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench, ref_trunks
from m_gat_graphsage_b200 import nn as mnn
from m_gat_graphsage_b200.accel import use_mgs_linear

dev = torch.device("cuda:0")
torch.manual_seed(42)
model = ref_trunks.Model1Trunk(mnn).to(dev).train()
use_mgs_linear(model)
opt = bench.make_adam(model.parameters(), lr=1e-4)
mnn.set_activation_fusion(True)
batches = bench.make_batches(dev, 0, 6)
for i in range(12):
    bench.drop_index_cache(batches[i % 6]); bench.train_step(model, opt, batches[i % 6])
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(30):
    b = batches[i % 6]
    bench.drop_index_cache(b)
    bench.train_step(model, opt, b)
    if i % 3 == 2:
        torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
