"""Micro-benchmark of the K4 projections at the model1 shapes (profiling aid, not a bench value)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from m_gat_graphsage_b200 import functional as Fm

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
M, K, N = 130512, int(sys.argv[3]) if len(sys.argv) > 3 else 350, int(sys.argv[4]) if len(sys.argv) > 4 else 350
g = torch.Generator(device=dev).manual_seed(0)
a = torch.randn(M, K, device=dev, generator=g)
x = torch.randn(M, K, device=dev, generator=g)
wl = torch.randn(N, K, device=dev, generator=g) / 18
wr = torch.randn(N, K, device=dev, generator=g) / 18
b = torch.randn(N, device=dev, generator=g)
go = torch.randn(M, N, device=dev, generator=g)
flush = torch.empty(64 * 1024 * 1024, device=dev)

def run():
    if which == "fwd":
        return Fm.linear_forward_raw(a, wl, b, x, wr)
    if which == "dgrad":
        return Fm.linear_dgrad_raw(go, wl)
    if which == "wgrad":
        return Fm.linear_wgrad_raw(go, a)

for _ in range(3):
    run()
torch.cuda.synchronize()
times = []
for _ in range(reps):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); run(); e.record(); torch.cuda.synchronize()
    times.append(s.elapsed_time(e))
flops = 2.0 * M * N * (2 * K if which == "fwd" else K)
t = sorted(times)[len(times) // 2]
print(f"{which}: median {t:.4f} ms  {flops / t / 1e9:.1f} TFLOP/s fp32-equivalent")
