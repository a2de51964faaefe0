"""Per-call timing of the atom-importance pass (BASELINE configs[3]) at batch 4096 (profiling aid)."""
import sys
from collections import defaultdict
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import ref_trunks
from m_gat_graphsage_b200 import _lib, nn as mnn
from m_gat_graphsage_b200.accel import use_mgs_linear
from m_gat_graphsage_b200.synth import synth_batch

dev = torch.device("cuda:0")
model = ref_trunks.build_trunk("model1", mnn).to(dev).eval()
use_mgs_linear(model)
b = synth_batch(4096, 42, device=dev)
lib = _lib.load()
for _ in range(3):
    ref_trunks.atom_importance(model, b)
acc = defaultdict(list)
tot = []
for _ in range(5):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib.start_profile()
    s.record()
    ref_trunks.atom_importance(model, b)
    e.record()
    recs = lib.stop_profile()
    tot.append(s.elapsed_time(e))
    seen = defaultdict(int)
    for name, a, ms in recs:
        seen[name] += 1
        acc[f"{name}#{seen[name]}"].append(ms)
print("importance pass ms:", sorted(tot)[2])
rows = sorted(((sum(v) / len(v), k) for k, v in acc.items()), reverse=True)
print("sum of libmgs calls:", round(sum(r[0] for r in rows), 3))
for ms, k in rows[:24]:
    print(f"{k:40s} {ms:.4f}")
