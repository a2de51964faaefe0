"""Why are the aggregation kernels 2x slower inside a step than alone?  (profiling aid, not a bench value)

Separates three candidate causes for every C-ABI call of SAGE / GAT message passing at the model1 shapes:
  data     : dense random rows vs post-ReLU rows (50 % exact zeros: IEEE division slow path, ...)
  producer : what ran right before (memset flush / a copy kernel leaving dirty lines / the real producer)
  idle     : same, but with the device left idle for a while after the producer
"""
import sys
from collections import defaultdict
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from m_gat_graphsage_b200 import _lib
from m_gat_graphsage_b200 import functional as Fm
from m_gat_graphsage_b200.graph import build_graph_index
from m_gat_graphsage_b200.synth import synth_batch

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 7
dev = torch.device("cuda:0")
b = synth_batch(4096, 42, device=dev)
N, E = b.x.size(0), b.edge_index.size(1)
gi = build_graph_index(b.edge_index, N)
gen = torch.Generator(device=dev).manual_seed(0)
dense = torch.randn(N, 350, device=dev, generator=gen)
sparse = torch.relu(dense)
go_dense = torch.randn(N, 350, device=dev, generator=gen)
go_sparse = go_dense * (torch.rand(N, 350, device=dev, generator=gen) < 0.1)
other = torch.randn(N, 350, device=dev, generator=gen)
flush = torch.empty(128 * 1024 * 1024, device=dev)
att = torch.randn(2, 350, device=dev, generator=gen)
lib = _lib.load()


def producer(kind):
    if kind == "memset":
        flush.zero_()
    elif kind == "copy":
        flush[: N * 350].view(N, 350).copy_(other)
        flush[N * 350: 2 * N * 350].view(N, 350).copy_(other)
    elif kind == "copy+idle":
        producer("copy")
        torch.cuda._sleep(2_000_000)  # ~1 ms of idle SMs
    elif kind == "none":
        pass


def sage(x, go):
    xr = x.detach().requires_grad_(True)
    out = Fm.sage_mean_aggregate(xr, gi)
    out.backward(go)


def gat(x, go):
    xr = x.detach().requires_grad_(True)
    out, _ = Fm.gat_message(xr, att[0].view(1, 10, 35), att[1].view(1, 10, 35), None, gi, 10, 35)
    out.backward(go)


def measure(fn, x, go, kind):
    acc = defaultdict(list)
    for _ in range(reps):
        producer(kind)
        lib.start_profile()
        fn(x, go)
        for name, _a, ms in lib.stop_profile():
            acc[name].append(ms)
    return {k: sorted(v)[len(v) // 2] for k, v in acc.items()}


for fn in (sage, gat):
    for _ in range(2):
        fn(dense, go_dense)
    torch.cuda.synchronize()
    rows = {}
    for data, (x, go) in {"dense": (dense, go_dense), "relu/sparse": (sparse, go_sparse)}.items():
        for kind in ("memset", "copy", "copy+idle", "none"):
            rows[(data, kind)] = measure(fn, x, go, kind)
    names = list(next(iter(rows.values())).keys())
    print(f"== {fn.__name__}: per-call median ms (N={N}, E={E}, F=350) ==")
    print(f"{'data':12s} {'producer':10s} " + " ".join(f"{n.replace('mgs_', ''):>16s}" for n in names))
    for (data, kind), r in rows.items():
        print(f"{data:12s} {kind:10s} " + " ".join(f"{r[n]:16.4f}" for n in names))

# the real in-model neighbours: x = relu(gat_out) produced right before the SAGE aggregation
for _ in range(3):
    xin = torch.relu(other)
    lib.start_profile()
    Fm.sage_mean_aggregate(xin, gi)
    print("sage_fwd right after torch.relu:", [round(ms, 4) for _n, _a, ms in lib.stop_profile()])
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(reps):
    flush.zero_()
    s.record(); y = dense.clone(); e.record(); torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
print(f"clone [N,350] (183 MB read + 183 MB write) after memset: median {sorted(ts)[len(ts)//2]:.4f} ms")
