"""Batched vs one-by-one GNNExplainer: molecules explained per second (profiling aid)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import ref_trunks
from m_gat_graphsage_b200 import nn as mnn
from m_gat_graphsage_b200.data import Data
from m_gat_graphsage_b200.explain import BatchedGNNExplainer, Explainer, GNNExplainer, ModelConfig
from m_gat_graphsage_b200.synth import synth_batch

dev = torch.device("cuda:0")
epochs = 100
trunk = ref_trunks.build_trunk("model1", mnn).to(dev).eval()
model = ref_trunks.ExplainableWrapper(trunk, Data).eval()
cfg = dict(explanation_type="model", node_mask_type="attributes", edge_mask_type="object",
           model_config=ModelConfig(mode="regression", task_level="graph", return_type="raw"))
b = synth_batch(1024, 77, device=dev)
one = Explainer(model=model, algorithm=GNNExplainer(epochs=epochs, lr=0.01), **cfg)
n_seq = 8
torch.cuda.synchronize(); t0 = time.perf_counter()
for g in range(n_seq):
    lo, hi = int(b.ptr[g]), int(b.ptr[g + 1])
    em = (b.edge_index[0] >= lo) & (b.edge_index[0] < hi)
    one(x=b.x[lo:hi].contiguous(), edge_index=(b.edge_index[:, em] - lo).contiguous(),
        batch=torch.zeros(hi - lo, dtype=torch.long, device=dev))
torch.cuda.synchronize(); t_seq = (time.perf_counter() - t0) / n_seq
many = Explainer(model=model, algorithm=BatchedGNNExplainer(epochs=epochs, lr=0.01), **cfg)
many(x=b.x, edge_index=b.edge_index, batch=b.batch)
torch.cuda.synchronize(); t0 = time.perf_counter()
many(x=b.x, edge_index=b.edge_index, batch=b.batch)
torch.cuda.synchronize(); t_bat = time.perf_counter() - t0
print(f"GNNExplainer, {epochs} epochs: one by one {t_seq * 1e3:.1f} ms / molecule ({1 / t_seq:.1f} molecules/s); "
      f"batched 1024 molecules {t_bat * 1e3:.1f} ms ({1024 / t_bat:.0f} molecules/s, {t_seq * 1024 / t_bat:.0f}x)")
