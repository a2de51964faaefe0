#!/usr/bin/env python
"""Benchmark of the M-GAT-GraphSAGE message-passing hot path on B200 (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of 4096 synthetic molecules per GPU:
K0 CSR build -> GATConv -> SAGEConv -> max||mean pooling -> readout MLP -> MSE -> backward -> Adam
(BASELINE.json configs[2], the configuration the metric "molecules/sec fwd+bwd at 1/2/4/8 B200" is
quoted on), model = the GATConv+SAGEConv trunk of /root/reference/ablation/model1.py:53-77, optimiser =
Adam(lr=1e-4) (model1.py:113), data-parallel over molecules (DDP / NCCL) for N > 1, weak scaling.
Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

METRIC = "molecules/sec fwd+bwd"
UNIT = "molecules/s"
BATCH = 4096
BASE_SEED = 42
N_DISTINCT_BATCHES = 6


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm": float(d["hbm_gbs"]), "tensor_burst": float(d["bf16_tflops"]),
                "tensor": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "source": "fallback"}


# --------------------------------------------------------------------------------------------------
# clocks (pynvml) sampled DURING the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.01)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def reset(self):
        """Forget what was sampled so far (the thread is started BEFORE the barrier that opens a timed region: its
        start-up and the first, cold NVML queries cost the launching thread up to 5 ms -- seen as a 9 ms first step
        at 2 GPUs -- and must not sit inside the region; samples are taken during it all the same)."""
        self.samples.clear()
        self.reasons.clear()

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "nvml unavailable"}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
# algorithmic bytes / flops per C-ABI call (DESIGN.md "Kernels")
# --------------------------------------------------------------------------------------------------
def call_cost(name, a, ctx):
    """-> (bound, algorithmic bytes, flops) for one libmgs call with ctypes args `a`."""
    N, E, B = ctx["N"], ctx["E"], ctx["B"]
    S = E + N
    if name == "mgs_csr_build":
        return "hbm", 16 * E + 8 * (N + 1) + 20 * E, 0
    if name == "mgs_graph_ptr":
        return "hbm", 8 * N + 4 * (B + 1), 0
    if name in ("mgs_sage_aggr_fwd", "mgs_sage_aggr_bwd"):
        n, f = a[2], a[3]
        return "hbm", 8 * n * f + 4 * (n + 1) + 4 * E, 0
    if name == "mgs_gat_scores_fwd":
        n, h, c = a[2], a[3], a[4]
        return "hbm", 4 * n * h * c + 8 * n * h + 8 * h * c, 0
    if name == "mgs_gat_alpha_fwd":
        n, h = a[2], a[3]
        return "hbm", 8 * n * h + 4 * (n + 1) + 4 * E + 4 * S * h, 0
    if name == "mgs_gat_aggr_fwd":
        n, h, c = a[2], a[3], a[4]
        return "hbm", 8 * n * h * c + 4 * S * h + 4 * (n + 1) + 4 * E, 0
    if name == "mgs_gat_bwd_edge":
        n, h, c = a[4], a[5], a[6]
        return "hbm", 8 * n * h * c + 8 * S * h + 12 * n * h + 4 * (n + 1) + 4 * E, 0
    if name == "mgs_gat_bwd_node":
        n, h, c = a[2], a[3], a[4]
        return "hbm", 8 * n * h * c + 8 * S * h + 8 * n * h + 8 * (n + 1) + 12 * E, 0
    if name == "mgs_gat_bwd_att":
        n, h, c = a[2], a[3], a[4]
        return "hbm", 4 * n * h * c + 8 * n * h, 0
    if name == "mgs_pool_fwd":
        b, f = a[3], a[4]
        return "hbm", 4 * N * f + 4 * b * f + 4 * (b + 1), 0
    if name == "mgs_pool_bwd":
        b, f, mode = a[7], a[8], a[9]
        return "hbm", (8 * N * f + 8 * b * f) if mode == 0 else (4 * N * f + 4 * b * f), 0
    if name == "mgs_sage_aggr_bwd_accumulate":      # reads g, base (+ the ReLU mask of the fused backward); writes gx
        n, f = a[2], a[3]
        return "hbm", (16 if a[11] else 12) * n * f + 4 * (n + 1) + 4 * E, 0
    if name == "mgs_proj_fwd":       # x[N,K] read, [N, n0+n1+n2] written, weights once
        n, k, nt = a[2], a[3], a[6] + a[9] + a[12]
        return "hbm", 4 * n * (k + nt) + 4 * k * nt, 2 * n * k * nt
    if name == "mgs_proj_wgrad":     # gradients [N, n0+n1+n2] and x[N,K] read once
        n, k, nt = a[11], a[12], a[2] + a[5] + a[8]
        return "hbm", 4 * n * (k + nt) + 4 * k * nt, 2 * n * k * nt
    if name == "mgs_pool_maxmean_fwd":
        b, f = a[3], a[4]
        return "hbm", 4 * N * f + 8 * b * f + 4 * (b + 1), 0
    if name == "mgs_pool_maxmean_bwd":
        b, f = a[7], a[8]
        return "hbm", 8 * N * f + 16 * b * f + 4 * (b + 1), 0
    if name == "mgs_linear_fwd":
        m, k, nout, k2 = a[2], a[3], a[6], a[10]
        kt = k + k2
        return _gemm_bound(kt, nout), 4 * (m * kt + nout * kt + m * nout), 2 * m * kt * nout
    if name == "mgs_linear_dgrad":
        m, nout, k = a[2], a[3], a[6]
        return _gemm_bound(nout, k), 4 * (m * nout + nout * k + m * k), 2 * m * nout * k
    if name == "mgs_linear_dgrad2":   # gx = g W_r + aggT(g) W_l as one GEMM (contraction N0 + N1), mask bits in the epilogue
        n0, n1, m, k = a[2], a[7], a[10], a[11]
        return _gemm_bound(n0 + n1, k), 4 * (m * (n0 + n1) + m * k + (n0 + n1) * k) + 4 * m * int(a[15]), 2 * m * (n0 + n1) * k
    if name == "mgs_linear_wgrad":
        m, nout, k = a[2], a[3], a[6]
        return _gemm_bound(nout, k), 4 * (m * nout + m * k + nout * k), 2 * m * nout * k
    if name == "mgs_colsum":
        m, nout = a[2], a[3]
        return "hbm", 4 * m * nout, 0
    if name == "mgs_adam_step":          # reads param, grad, exp_avg, exp_avg_sq; writes param and both moments
        return "hbm", 28 * sum(int(v) for v in a[5][:int(a[0])]), 0
    return "hbm", 0, 0


# DRAM bytes per C-ABI call measured with `ncu --set full` over ONE training step of this very command
# (`tools/profile_step.sh`: bench.py --ncu-steps 1 under ncu, joined with the ordered call log by tools/ncu_join.py).
# bench.py cannot run under a profiler, so the per-call table is read from the committed capture.
NCU_CALLS_FILE = ROOT / "profiles" / "round2i_ncu_calls.json"


def load_ncu_calls():
    try:
        d = json.loads(NCU_CALLS_FILE.read_text())
        return d.get("calls", {}), d.get("captured_on", None)
    except Exception:
        return {}, None


def call_key(name, a):
    """Key of one C-ABI call in the per-call tables: the entry point plus the shape arguments that select a different
    kernel (M varies with the batch and is not part of the key)."""
    if name == "mgs_linear_fwd":
        return f"{name}[K={a[3]}+{a[10]},N={a[6]}]"
    if name in ("mgs_linear_dgrad", "mgs_linear_wgrad"):
        return f"{name}[N={a[3]},K={a[6]}]"
    if name == "mgs_linear_dgrad2":
        return f"{name}[N={a[2]}+{a[7]},K={a[11]}]"
    if name in ("mgs_pool_fwd", "mgs_pool_bwd"):
        return f"{name}[mode={a[5] if name == 'mgs_pool_fwd' else a[9]}]"
    if name == "mgs_colsum":
        return f"{name}[N={a[3]}]"
    return name


def _gemm_bound(d1, d2):
    # fp32-accurate tensor-core GEMM is 3 TF32 passes; below ~128 in either dimension the projection is HBM-bound
    return "tensor" if min(d1, d2) >= 128 else "hbm"


# --------------------------------------------------------------------------------------------------
def make_batches(device, rank, n, batch_size=BATCH):
    from m_gat_graphsage_b200.synth import batch_seed, synth_batch
    return [synth_batch(batch_size, batch_seed(BASE_SEED, rank, i), device=device) for i in range(n)]


def drop_index_cache(batch):
    """K0 is part of every step: forget the CSR / segment pointers cached on the index tensors."""
    for t in (batch.edge_index, batch.batch):
        for attr in ("_mgs_graph", "_mgs_gptr"):
            if hasattr(t, attr):
                delattr(t, attr)


def make_adam(params, **kw):
    if os.environ.get("MGS_BENCH_TORCH_ADAM", "0") == "1":
        return torch.optim.Adam(params, fused=True, **kw)
    from m_gat_graphsage_b200.accel import FusedAdam
    return FusedAdam(params, **kw)


def train_step(model, opt, batch):
    opt.zero_grad(set_to_none=True)
    loss = F.mse_loss(model(batch).view(-1), batch.y)
    loss.backward()
    opt.step()
    return loss


def run_ours(args):
    import torch.distributed as dist

    import ref_trunks
    from m_gat_graphsage_b200 import _lib
    from m_gat_graphsage_b200 import nn as mnn
    from m_gat_graphsage_b200.data import Batch, _tag_num_graphs
    from m_gat_graphsage_b200.synth import batch_seed, synth_batch
    from torch.nn.parallel import DistributedDataParallel as DDP

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.backends.cudnn.allow_tf32 = False          # the reference is fp32 end to end (stock layers included)
    torch.backends.cuda.matmul.allow_tf32 = False
    if world > 1:
        # NCCL writes its banner / debug lines ("NCCL version ...") to stdout unless told otherwise; stdout carries
        # exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)                     # (the banner ignores NCCL_DEBUG_FILE on this build: point fd 1 at stderr)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()                # communicator is created here at the latest
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    lib = _lib.load()
    # the `self.relu(conv(..))` / `self.relu(self.fc_g1(..))` of the reference's own forward (model1.py:68-74) ride on
    # the producing kernels' epilogues, their backward on the kernels that produce the gradients (lazy.py)
    mnn.set_activation_fusion(not args.no_activation_fusion)

    torch.manual_seed(BASE_SEED)
    model = ref_trunks.Model1Trunk(mnn).to(dev).train()
    from m_gat_graphsage_b200.accel import use_mgs_linear
    use_mgs_linear(model)                       # readout MLP (fc_g1 / fc_g2 / out) on the K4 kernels too
    n_params = sum(p.numel() for p in model.parameters())
    step_model = model
    if world > 1:
        # 2 MB buckets: the readout MLP's gradients (fc_g1 = 4.2 of the 6 MB) are complete right after the MLP backward,
        # so their all-reduce overlaps the whole message-passing backward; only the 1 MB conv bucket is exposed
        step_model = DDP(model, device_ids=[local_rank], bucket_cap_mb=float(os.environ.get("MGS_BENCH_BUCKET_MB", "2")),
                         gradient_as_bucket_view=True,
                         # the autograd graph of the step never changes: DDP skips its per-iteration bookkeeping (measured at
                         # N = 2: 3.18 -> 3.08 ms per step; bucket sizes between 0.5 and 25 MB made no difference)
                         static_graph=os.environ.get("MGS_BENCH_STATIC_GRAPH", "1") == "1")
    # model1.py:113 optimiser and hyper-parameters.  accel.FusedAdam = the same update as ONE mgs_adam_step launch over
    # every parameter tensor (csrc/adam.cu; PyTorch's `fused=True` deals 64 Ki-element chunks: 26 CTAs, 80 us per step;
    # its default foreach path is ~12 latency-bound launches).  MGS_BENCH_TORCH_ADAM=1 times torch.optim.Adam(fused=True).
    opt = make_adam(model.parameters(), lr=1e-4)

    batches = make_batches(dev, rank, N_DISTINCT_BATCHES)
    ctx0 = {"N": batches[0].x.size(0), "E": batches[0].edge_index.size(1), "B": BATCH}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident throughput (`value`) ----------------
    # Pre-conditioning (untimed, reported as `config.preconditioning_steps`): every distinct batch shape is seen twice
    # (allocator growth, lazy module loading, clocks: on a fresh box the first ~10 steps run ~10 % slower), then the
    # settle loop below.  The W warm-up steps the command line asks for are run AFTER it, right before the timed region.
    precond = 0
    for i in range(2 * N_DISTINCT_BATCHES):
        b = batches[i % len(batches)]
        drop_index_cache(b)
        train_step(step_model, opt, b)
        precond += 1
    barrier()
    # Keep warming up (untimed, at most ~3 s) until the step time is steady: on a fresh box the image is still
    # paging in and the host can be too slow to keep the GPU fed for the first seconds (seen: 4.5 instead of 3.4 ms).
    t_settle, prev = time.perf_counter(), None
    while args.ncu_steps == 0:
        s0_, s1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0_.record()
        for i in range(len(batches)):
            drop_index_cache(batches[i])
            train_step(step_model, opt, batches[i])
        s1_.record()
        torch.cuda.synchronize()
        precond += len(batches)
        cur = s0_.elapsed_time(s1_)
        stop = (prev is not None and abs(cur - prev) <= 0.03 * prev) or time.perf_counter() - t_settle >= 3.0
        prev = cur
        if world > 1:                                   # every rank takes the same number of extra steps
            flag = torch.tensor([1.0 if stop else 0.0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            stop = bool(flag.item() > 0.5)
        if stop:
            break
    barrier()
    if args.ncu_steps > 0:
        # profiling aid (never a bench value): `ncu --profile-from-start off ... bench.py --ncu-steps 2`
        # captures exactly these steps (cudaProfilerStart/Stop), not data generation or warm-up
        lib.start_call_log()
        torch.cuda.cudart().cudaProfilerStart()
        for i in range(args.ncu_steps):
            b = batches[i % len(batches)]
            drop_index_cache(b)
            train_step(step_model, opt, b)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        calls = [{"call": call_key(n, a), "kernels": k} for n, a, k in lib.stop_call_log()]
        if rank == 0:
            if args.call_log:
                Path(args.call_log).write_text(json.dumps({"steps": args.ncu_steps, "batch": BATCH, "calls": calls}))
            print(json.dumps({"ncu_steps": args.ncu_steps, "libmgs_calls": len(calls),
                              "note": "profiling run, not a benchmark value"}))
        return
    launches0 = _lib.launch_count()
    gc.collect()
    gc.disable()          # no cyclic-GC pause inside a timed region (re-enabled after the end-to-end leg)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(local_rank) as clocks:
        # sampler thread up and its first (cold) NVML queries done while the GPU keeps working, so that the timed
        # region neither contains them nor starts on an idle, down-clocked GPU
        for i in range(args.warmup):                       # the W warm-up steps of the contract
            drop_index_cache(batches[i % len(batches)])
            train_step(step_model, opt, batches[i % len(batches)])
        barrier()
        clocks.reset()
        launches0 = _lib.launch_count()
        ev0.record()
        marks[0].record()
        for i in range(args.steps):
            b = batches[i % len(batches)]
            drop_index_cache(b)
            train_step(step_model, opt, b)
            marks[i + 1].record()
        ev1.record()
        barrier()
    total_ms = max_over_ranks(ev0.elapsed_time(ev1))
    per_step = [marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps)]
    if rank == 0:
        print("per-step ms: " + " ".join(f"{t:.2f}" for t in per_step), file=sys.stderr)
    launches = _lib.launch_count() - launches0
    ms_per_step = total_ms / args.steps
    value = world * BATCH * args.steps / (total_ms / 1e3)

    # ---------------- end to end: pinned host buffers -> H2D -> step -> loss D2H ----------------
    # Wire format (data.WireBatch): the atom features are exactly 0 / 1 (train.py:33-43) and cross PCIe as one bit each,
    # edge_index as int32, the batch vector as B + 1 segment pointers: 3.2 instead of 23.6 MB per 4096 molecules; ONE
    # launch (mgs_wire_expand) rebuilds x fp32 / edge_index int64 / batch int64 on the device, bit-identical.
    from m_gat_graphsage_b200.data import WireBatch
    host = [WireBatch.from_batch(b) for b in batches]
    h2d = max(w.nbytes for w in host)
    h2d_fp32 = sum(getattr(batches[0], k).numel() * getattr(batches[0], k).element_size() for k in ("x", "edge_index", "batch", "y"))

    # Input pipeline as a training loop with a pinned-memory loader runs it: the H2D copies of step i+1 are issued
    # on a copy stream while step i computes (every step's inputs still cross PCIe inside the timed region, and
    # every step ends with a D2H read of its loss).
    copy_stream = torch.cuda.Stream(device=dev)
    # two sets of device input buffers (largest batch), filled alternately: no allocator traffic across streams
    n_max = max(w.xbits.numel() for w in host)
    e_max = max(w.edge_index.size(1) for w in host)

    def make_bufs():
        i64, i32, f32 = torch.int64, torch.int32, torch.float32
        return {"xbits": torch.empty(n_max, dtype=i64, device=dev), "edge_index32": torch.empty(2 * e_max, dtype=i32, device=dev),
                "ptr32": torch.empty(BATCH + 1, dtype=i32, device=dev), "ptr": torch.empty(BATCH + 1, dtype=i64, device=dev),
                "y": torch.empty(BATCH, dtype=f32, device=dev), "x": torch.empty(n_max * 35, dtype=f32, device=dev),
                "edge_index": torch.empty(2 * e_max, dtype=i64, device=dev), "batch": torch.empty(n_max, dtype=i64, device=dev)}

    dev_buf = [make_bufs() for _ in range(2)]
    free_ev = [None, None]          # recorded on the compute stream when a step has consumed its buffer set

    h2d_marks = []

    def upload(h, slot):
        with torch.cuda.stream(copy_stream):
            if free_ev[slot] is not None:
                copy_stream.wait_event(free_ev[slot])
            m0 = torch.cuda.Event(enable_timing=True)
            m0.record(copy_stream)
            bt = h.to_batch(dev, buffers=dev_buf[slot])          # copies + the expansion launch, on the copy stream
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(copy_stream)
            h2d_marks.append((m0, ev))
        return bt, ev, slot

    def e2e_step(staged):
        b, ev, slot = staged
        torch.cuda.current_stream().wait_event(ev)
        loss = train_step(step_model, opt, b)
        free_ev[slot] = torch.cuda.Event()
        free_ev[slot].record()
        return loss

    n_e2e_warm = max(2, args.warmup // 2)
    staged = upload(host[0], 0)
    for i in range(n_e2e_warm):
        nxt = upload(host[(i + 1) % len(host)], (i + 1) & 1)
        float(e2e_step(staged).item())
        staged = nxt
    torch.cuda.synchronize()
    barrier()
    # every step's loss is copied to pinned host memory right behind the step and READ (host side) two steps later,
    # so the host never idles the GPU while it prepares the next step's launches
    LAG = 2                                                       # the host reads a loss LAG steps after its step
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(LAG + 1)]
    loss_ev = [torch.cuda.Event() for _ in range(LAG + 1)]
    # The region is timed E2E_REPS times back to back and the MEDIAN repetition is reported (all are listed in
    # `reps_ms_per_step`): a single host hiccup on a fresh box -- seen once as 9.2 instead of 3.5 ms per step -- would
    # otherwise decide a number that is bound by the host's launch rate by design.
    E2E_REPS = 3
    rep_ms, host_ts = [], []
    for rep in range(E2E_REPS):
        for ev in free_ev:
            if ev is not None:
                ev.synchronize()
        barrier()
        losses = []
        t0 = time.perf_counter()
        ev0.record()
        staged = upload(host[0], 0)
        host_ts = []
        for i in range(args.steps):
            host_ts.append(time.perf_counter())
            nxt = upload(host[(i + 1) % len(host)], (i + 1) & 1) if i + 1 < args.steps else None
            loss = e2e_step(staged)
            loss_host[i % (LAG + 1)].copy_(loss.detach(), non_blocking=True)   # D2H read of this step's loss
            loss_ev[i % (LAG + 1)].record()
            if i >= LAG:
                loss_ev[(i - LAG) % (LAG + 1)].synchronize()
                losses.append(float(loss_host[(i - LAG) % (LAG + 1)]))
            staged = nxt
        for i in range(max(0, args.steps - LAG), args.steps):
            loss_ev[i % (LAG + 1)].synchronize()
            losses.append(float(loss_host[i % (LAG + 1)]))
        ev1.record()
        barrier()
        assert len(losses) == args.steps and all(v == v for v in losses), "e2e: every step's loss must reach the host"
        wall_ms = (time.perf_counter() - t0) * 1e3
        rep_ms.append(max_over_ranks(max(ev0.elapsed_time(ev1), wall_ms)))
    e2e_ms = statistics.median(rep_ms)
    e2e_value = world * BATCH * args.steps / (e2e_ms / 1e3)
    gc.enable()
    h2d_ms = sorted(a.elapsed_time(b) for a, b in h2d_marks[-args.steps:])
    h2d_ms_med = h2d_ms[len(h2d_ms) // 2] if h2d_ms else float("nan")
    if rank == 0:
        host_ts.append(t0 + wall_ms / 1e3)
        print("e2e host ms between step starts: " + " ".join(f"{(b - a) * 1e3:.2f}" for a, b in zip(host_ts, host_ts[1:])),
              file=sys.stderr)

    peaks = measured_peaks()
    ncu_calls, ncu_when = load_ncu_calls()

    def attribute(step_fn, bs, mols_per_batch, reps=5):
        """Per-call attribution of one step: every C-ABI call bracketed by CUDA events (rank 0 only; a separate,
        instrumented pass) -> (per-call table sorted by time, roofline object of the dominant call)."""
        per, step_ms = {}, []
        for i in range(reps):
            bb = bs[i % len(bs)]
            drop_index_cache(bb)
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            lib.start_profile()
            s0.record()
            step_fn(bb)
            s1.record()
            recs = lib.stop_profile()
            step_ms.append(s0.elapsed_time(s1))
            ctx = {"N": bb.x.size(0), "E": bb.edge_index.size(1), "B": mols_per_batch}
            seen = {}
            for name, a, ms in recs:
                bound, nbytes, flops = call_cost(name, a, ctx)
                key = call_key(name, a)
                seen[key] = seen.get(key, 0) + 1
                k2 = f"{key}#{seen[key]}" if seen[key] > 1 else key
                e = per.setdefault(k2, {"bound": bound, "bytes": [], "flops": [], "ms": []})
                e["ms"].append(ms), e["bytes"].append(nbytes), e["flops"].append(flops)
        step_med = statistics.median(step_ms)
        table = []
        for key, v in per.items():
            ms = statistics.mean(v["ms"])
            nbytes, flops = statistics.mean(v["bytes"]), statistics.mean(v["flops"])
            if v["bound"] == "tensor":
                ach, peak, unit = flops / (ms * 1e-3) / 1e12, peaks["tensor"], "TFLOP/s"
            else:
                ach, peak, unit = nbytes / (ms * 1e-3) / 1e9, peaks["hbm"], "GB/s"
            cap = ncu_calls.get(key) or {}
            table.append({"call": key, "bound": v["bound"], "ms": round(ms, 4), "share_of_step": round(ms / step_med, 4),
                          "achieved": round(ach, 2), "peak": peak, "unit": unit, "frac": round(ach / peak, 4),
                          "alg_bytes": int(nbytes), "alg_flops": int(flops), "ncu_dram_bytes": cap.get("dram_bytes"),
                          "ncu": {k: v for k, v in cap.items() if k != "dram_bytes"} or None})
        table.sort(key=lambda k: -k["ms"])
        top = table[0]
        roof = {"kernel": top["call"], "bound": top["bound"], "achieved": top["achieved"], "peak": top["peak"],
                "unit": top["unit"], "frac": top["frac"], "traffic": top["ncu_dram_bytes"],
                "traffic_source": (f"ncu --set full over one step of this command ({NCU_CALLS_FILE.relative_to(ROOT)}, "
                                   f"captured {ncu_when}): dram__bytes_read.sum + dram__bytes_write.sum of the call's "
                                   "kernels") if top["ncu_dram_bytes"] is not None else None,
                "peak_source": f"{peaks['source']} ({'bf16 sustained' if top['bound'] == 'tensor' else 'HBM copy'})",
                # an fp32-accurate tensor-core GEMM is three TF32 passes at half the bf16 rate: its own ceiling
                "ceiling_3xtf32": round(peaks["tensor"] / 6, 1) if top["bound"] == "tensor" else None,
                "frac_of_3xtf32_ceiling": round(top["achieved"] / (peaks["tensor"] / 6), 4) if top["bound"] == "tensor" else None,
                "share_of_step": top["share_of_step"], "instrumented_step_ms": round(step_med, 3),
                "libmgs_share_of_step": round(sum(k["ms"] for k in table) / step_med, 4)}
        return table, roof

    def timed_leg(step_fn, bs, steps, warm, finish=None):
        """`warm` untimed steps, barrier, `steps` timed steps (+ `finish()`, e.g. the final gather) between CUDA events,
        barrier, max over ranks -> total milliseconds."""
        for i in range(warm):
            drop_index_cache(bs[i % len(bs)])
            step_fn(bs[i % len(bs)], i, False)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            drop_index_cache(bs[i % len(bs)])
            step_fn(bs[i % len(bs)], i, True)
        if finish is not None:
            finish()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    def compact(table, n=6):
        return [{k: r[k] for k in ("call", "bound", "ms", "share_of_step", "achieved", "unit", "frac")} for r in table[:n]]

    # ---------------- per-kernel attribution of the training step (rank 0) ----------------
    roofline, kernels = None, []
    if rank == 0:
        kernels, roofline = attribute(lambda bb: train_step(model, opt, bb), batches, BATCH)   # un-wrapped module: no
        roofline["calls"] = kernels                                  # collective while the other ranks idle
    barrier()

    other = {}
    if not args.no_other_configs:
        max_atoms = max(b.x.size(0) for b in batches)
        if world > 1:
            t = torch.tensor([max_atoms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            max_atoms = int(t.item())
        K = args.steps

        # ---------- configs[1]: inference, batch 4096 per GPU, final gather of the predictions ----------
        model.eval()
        pred_buf = torch.empty(K, BATCH, device=dev)
        pred_all = torch.empty(world, K, BATCH, device=dev) if world > 1 else pred_buf
        pred_host = torch.empty(pred_all.shape, dtype=torch.float32).pin_memory() if rank == 0 else None

        def infer_step(bb, i, timed):
            with torch.no_grad():
                out = model(bb)
            if timed:
                pred_buf[i].copy_(out.view(-1))

        def infer_finish():
            if world > 1:
                dist.all_gather_into_tensor(pred_all.view(-1), pred_buf.view(-1))
            if rank == 0:
                pred_host.copy_(pred_all, non_blocking=True)

        ms = timed_leg(infer_step, batches, K, 3, infer_finish)
        other["configs[1] inference"] = {
            "value": round(world * BATCH * K / (ms / 1e3), 1), "unit": UNIT, "ms_per_step": round(ms / K, 4),
            "what": f"forward only (eval, no_grad) incl. K0, batch {BATCH} per GPU, {K} steps, then ONE final gather of the "
                    f"[{world} x {K} x {BATCH}] predictions (NCCL all_gather) and their D2H copy on rank 0, inside the timed region"}
        # end to end: every step's inputs come from pinned host memory
        staged = upload(host[0], 0)
        for ev in free_ev:
            if ev is not None:
                ev.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(K):
            nxt = upload(host[(i + 1) % len(host)], (i + 1) & 1) if i + 1 < K else None
            bb, ev, slot = staged
            torch.cuda.current_stream().wait_event(ev)
            infer_step(bb, i, True)
            free_ev[slot] = torch.cuda.Event()
            free_ev[slot].record()
            staged = nxt
        infer_finish()
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
        other["configs[1] inference"]["e2e"] = {
            "value": round(world * BATCH * K / (ms_e2e / 1e3), 1), "unit": UNIT, "ms_per_step": round(ms_e2e / K, 4),
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * BATCH * (world if rank == 0 else 1)}
        if rank == 0:
            with torch.no_grad():
                tab, roof = attribute(lambda bb: model(bb), batches, BATCH, reps=3)
            other["configs[1] inference"]["roofline"] = roof
            other["configs[1] inference"]["calls"] = compact(tab)
        barrier()

        # ---------- configs[3]: per-atom gradient-L2 importance, final gather of [N_atoms] ----------
        imp_buf = torch.zeros(K, max_atoms, device=dev)
        imp_all = torch.empty(world, K, max_atoms, device=dev) if world > 1 else imp_buf
        imp_host = torch.empty(imp_all.shape, dtype=torch.float32).pin_memory() if rank == 0 else None

        def imp_step(bb, i, timed):
            imp = ref_trunks.atom_importance(model, bb)            # gnnexplainer.py:647-652, parameters frozen
            if timed:
                imp_buf[i, : imp.numel()].copy_(imp)

        def imp_finish():
            if world > 1:
                dist.all_gather_into_tensor(imp_all.view(-1), imp_buf.view(-1))
            if rank == 0:
                imp_host.copy_(imp_all, non_blocking=True)

        ms = timed_leg(imp_step, batches, K, 3, imp_finish)
        atoms = sum(batches[i % len(batches)].x.size(0) for i in range(K))
        other["configs[3] atom importance"] = {
            "value": round(world * BATCH * K / (ms / 1e3), 1), "unit": UNIT, "ms_per_step": round(ms / K, 4),
            "atoms_per_s": round(world * atoms / (ms / 1e3), 1),
            "what": f"forward + backward w.r.t. x only (parameters frozen) + per-atom L2 norm, incl. K0, batch {BATCH} per GPU, "
                    f"{K} steps, then ONE final gather of the per-atom importances ([{world} x {K} x {max_atoms}] fp32, NCCL "
                    "all_gather) and their D2H copy on rank 0, inside the timed region",
            "gather_bytes": int(imp_all.numel() * 4)}
        if rank == 0:
            tab, roof = attribute(lambda bb: ref_trunks.atom_importance(model, bb), batches, BATCH, reps=3)
            other["configs[3] atom importance"]["roofline"] = roof
            other["configs[3] atom importance"]["calls"] = compact(tab)
        barrier()
        model.train()
        del imp_buf, imp_all, pred_buf, pred_all

        # ---------- configs[4]: stress shape, training step ----------
        SB, SK = args.stress_batch, max(3, K // 4)
        torch.manual_seed(BASE_SEED)
        stress = ref_trunks.build_trunk("stress", mnn).to(dev).train()
        use_mgs_linear(stress)
        stress_step_model = stress
        if world > 1:
            stress_step_model = DDP(stress, device_ids=[local_rank], bucket_cap_mb=2, gradient_as_bucket_view=True,
                                    static_graph=True)
        sopt = make_adam(stress.parameters(), lr=1e-4)
        sb = [synth_batch(SB, batch_seed(BASE_SEED, rank, 100 + i), device=dev, fixed_atoms=94) for i in range(2)]
        torch.cuda.reset_peak_memory_stats()
        ms = timed_leg(lambda bb, i, timed: train_step(stress_step_model, sopt, bb), sb, SK, 3)
        other["configs[4] stress"] = {
            "value": round(world * SB * SK / (ms / 1e3), 1), "unit": UNIT, "ms_per_step": round(ms / SK, 3),
            "atoms_per_s": round(world * sb[0].x.size(0) * SK / (ms / 1e3), 1),
            "batch_per_gpu": SB, "atoms_per_batch": sb[0].x.size(0), "edges_per_batch": sb[0].edge_index.size(1),
            "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2), "steps": SK,
            "what": "training step (K0 + GATConv(35, 32, heads=8) + SAGEConv(256, 256) + max||mean + MLP 512-1500-128-1 + MSE "
                    "+ backward + Adam" + (" + DDP all-reduce" if world > 1 else "") + "), every molecule 94 atoms"}
        if rank == 0:
            tab, roof = attribute(lambda bb: train_step(stress, sopt, bb), sb, SB, reps=2)
            other["configs[4] stress"]["roofline"] = roof
            other["configs[4] stress"]["calls"] = compact(tab)
        barrier()
        del stress, stress_step_model, sopt, sb
        torch.cuda.empty_cache()

    # ---------------- the complete train.py step: GNN trunk + ECFP CNNNet + CombinedNet, mse + kl ----------
    if not args.no_other_configs:
        from m_gat_graphsage_b200.attention import molecule_attention, use_mgs_attention

        def train_py_leg(bsz, per_molecule, mgs_linear, steps):
            torch.manual_seed(BASE_SEED)
            full = ref_trunks.TrainPyModel(mnn).to(dev).train()
            use_mgs_attention(full)                                   # ModifiedGATLayer -> K5 (no [N, N] matrices)
            if mgs_linear:
                use_mgs_linear(full)                                  # every nn.Linear incl. CNNNet.fc1 (131072 -> 256) on K4
            step_full = DDP(full, device_ids=[local_rank], gradient_as_bucket_view=True) if world > 1 else full
            fopt = make_adam(full.parameters(), lr=1e-3, weight_decay=1e-4)                         # train.py:216-222
            fb = [synth_batch(bsz, batch_seed(BASE_SEED, rank, 200 + i), device=dev) for i in range(2)]
            gen = torch.Generator(device=dev).manual_seed(BASE_SEED + rank)
            ecfp = [(torch.rand(bsz, 1, 1024, device=dev, generator=gen) < 0.05).float() for _ in range(2)]

            def step(bb, i, timed):
                fopt.zero_grad(set_to_none=True)
                e = ecfp[i % 2]
                if per_molecule:
                    with molecule_attention(bb.batch):
                        out, combined = step_full(bb, e)
                else:
                    out, combined = step_full(bb, e)
                loss = F.mse_loss(out, bb.y.view(-1, 1)) + 0.001 * ref_trunks.kl_loss(combined)
                loss.backward()
                fopt.step()

            torch.cuda.reset_peak_memory_stats()
            ms = timed_leg(step, fb, steps, 3)
            res = {"value": round(world * bsz * steps / (ms / 1e3), 1), "unit": UNIT, "ms_per_step": round(ms / steps, 3),
                   "batch_per_gpu": bsz, "steps": steps, "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
            # share of the step spent in the GNN trunk alone (same batch, same attention mode, trunk forward + backward)
            trunk = full.gat_graphsage_model

            def trunk_step(bb, i, timed):
                if per_molecule:
                    with molecule_attention(bb.batch):
                        out = trunk(bb)
                else:
                    out = trunk(bb)
                out.sum().backward()
                for prm in trunk.parameters():
                    prm.grad = None

            ms_t = timed_leg(trunk_step, fb, steps, 2)
            res["gnn_trunk_ms_per_step"] = round(ms_t / steps, 3)
            res["cnn_and_head_share"] = round(1.0 - (ms_t / steps) / (ms / steps), 3)
            del full, step_full, fopt, fb, ecfp
            torch.cuda.empty_cache()
            return res

        K4 = max(3, args.steps // 4)
        legs = {}
        legs["batch 128 (the script's), whole-batch attention, stock cuDNN/cuBLAS CNN"] = train_py_leg(128, False, False, K4)
        legs["batch 128, whole-batch attention, nn.Linear on K4"] = train_py_leg(128, False, True, K4)
        legs["batch 4096, per-molecule attention, stock cuDNN/cuBLAS CNN"] = train_py_leg(4096, True, False, K4)
        legs["batch 4096, per-molecule attention, nn.Linear on K4"] = train_py_leg(4096, True, True, K4)
        other["train.py full step"] = {
            "what": "one optimisation step of /root/reference/train.py:236-249: ModifiedGATLayer (K5) -> SAGEConv -> max pool -> "
                    "MLP, CNNNet over a synthetic 1024-bit fingerprint, CombinedNet, loss = mse + 0.001 kl, backward, "
                    "Adam(lr=1e-3, weight_decay=1e-4); 34.6 M parameters (138.6 MB of gradients"
                    + (", all-reduced by DDP" if world > 1 else "") + "); convolutions and, in the 'stock' legs, all nn.Linear "
                    "layers are stock PyTorch fp32 (TF32 off)", "legs": legs}
        barrier()

    # ---------------- batch-1 latency: the loop of test.py:175-208 / gnnexplainer.py:1414-1420 (one molecule per forward,
    # one D2H read per molecule) -- eager launches vs one CUDA-graph replay per molecule (graphed.GraphedStep) ------------
    if rank == 0 and not args.no_other_configs:
        from m_gat_graphsage_b200.graphed import GraphedStep
        from oracle import pyg_oracle as O_
        torch.manual_seed(BASE_SEED)
        one = ref_trunks.build_trunk("model1", mnn).to(dev).eval()
        use_mgs_linear(one)
        stock = ref_trunks.build_trunk("model1", O_).to(dev).eval()
        mols = synth_batch(256, batch_seed(BASE_SEED, 0, 300), device=dev).to_data_list()
        mols = [Batch.from_data_list([m]) for m in mols]
        host_out = torch.empty(1, 1).pin_memory()

        def loop(fn, n):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(n):
                host_out.copy_(fn(mols[i % len(mols)]))                 # `.cpu()` per molecule, as test.py:197 does
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / n * 1e3

        with torch.no_grad():
            graphed = GraphedStep(one, 1, max_nodes=128, max_edges=320)
            res1 = {}
            for name_, fn in (("stock PyTorch eager (oracle ops on this GPU)", lambda m: stock(m)),
                              ("eager launches (this repository)", lambda m: one(m)),
                              ("one CUDA-graph replay per molecule (GraphedStep)", lambda m: graphed(m))):
                loop(fn, 64)
                res1[name_] = {"ms_per_molecule": round(loop(fn, 512), 4)}
                res1[name_]["molecules_per_s"] = round(1e3 / res1[name_]["ms_per_molecule"], 1)
        other["batch-1 latency (test.py loop)"] = {
            "what": "model1 trunk, one molecule per forward incl. K0, prediction copied to the host after every molecule "
                    "(test.py:175-208); host-clock latency per molecule over 512 molecules of 11-94 atoms", "legs": res1,
            "graph_replays": graphed.replays, "eager_fallbacks": graphed.eager}
        del one, stock, graphed
    barrier()

    # ---------------- stock PyTorch eager on this GPU: the oracle's op chains run op by op on the B200 ----------------
    gpu_eager = None
    if rank == 0 and world == 1 and not args.no_other_configs:
        gpu_eager = gpu_eager_reference(dev, batches, steps=max(5, args.steps // 2), warmup=4)

    # ---------------- CPU baseline on this box's host cores (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(steps=3, warmup=1, budget_s=25.0)

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {
            "workload": "BASELINE configs[2] training step: K0 CSR build + GATConv(35,35,heads=10) + SAGEConv(350,350) "
                        "+ global max||mean pool + MLP 700-1500-128-1 (ablation/model1.py trunk), MSE, backward, "
                        "Adam(lr=1e-4); 4096 synthetic molecules per GPU per step (11-94 atoms, mean 31.8, deg<=6)",
            "batch_per_gpu": BATCH, "atoms_per_batch": ctx0["N"], "edges_per_batch": ctx0["E"],
            "parameters": n_params, "parallelism": f"dp{world}" + (" (DDP static_graph, NCCL all-reduce of 6.0 MB grads in 2 MB buckets, overlapped with backward)" if world > 1 else ""),
            "l2": f"{N_DISTINCT_BATCHES} distinct batches cycled; per-step working set ~3 GB >> 126 MB L2 (inputs larger than L2)",
            "size_distribution": "assumption: n=clip(round(exp(N(ln30,0.35^2))),11,94) (SURVEY.md Appendix C)",
        },
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": round(e2e_ms / args.steps, 4),
                "h2d_ms_per_step": round(h2d_ms_med, 4), "h2d_GBps": round(h2d / max(h2d_ms_med, 1e-9) / 1e6, 1),
                "reps_ms_per_step": [round(m / args.steps, 4) for m in rep_ms],
                "h2d_bytes_per_step_unpacked": h2d_fp32,
                "what": "pinned host wire image of every batch (atom features as 1 bit each: they are exactly 0 / 1 in the "
                        "reference's featurisation; edge_index int32; B + 1 segment pointers; y) -> H2D on a copy stream (step "
                        "i+1 uploads while step i computes) -> one expansion launch rebuilds x[N,35] f32 / edge_index[2,E] i64 "
                        "/ batch[N] i64 bit-identically -> same step -> loss copied to pinned host memory every step, read two "
                        "steps later; the K-step region is timed three times, the median repetition is reported"},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    line["config"]["preconditioning_steps"] = precond
    line["config"]["other_configs"] = other
    if cpu is not None:
        cpu["gpu_eager_baseline"] = gpu_eager
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement (PyG itself is not installable: "port") on all host cores
# --------------------------------------------------------------------------------------------------
def cpu_reference(steps, warmup, budget_s):
    import ref_trunks
    from m_gat_graphsage_b200.synth import batch_seed, synth_batch
    from oracle import pyg_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(BASE_SEED)
    model = ref_trunks.Model1Trunk(O).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    # calibrate the sample so that (steps + warmup) steps fit the budget
    probe = synth_batch(128, batch_seed(BASE_SEED, 0, 999))
    train_step(model, opt, probe)
    t0 = time.perf_counter()
    train_step(model, opt, probe)
    per_mol = (time.perf_counter() - t0) / 128
    sample = int(min(BATCH, max(128, budget_s / max(per_mol, 1e-9) / (steps + warmup))))
    sample = max(128, (sample // 128) * 128)
    batches = [synth_batch(sample, batch_seed(BASE_SEED, 0, i)) for i in range(2)]
    for i in range(warmup):
        train_step(model, opt, batches[i % 2])
    times = []
    for i in range(steps):
        t0 = time.perf_counter()
        train_step(model, opt, batches[i % 2])
        times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": round(sample * steps / total, 1), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} training steps of {sample} molecules each (same model, optimiser and generator as the GPU arm; "
                      f"restatement of PyG in PyTorch CPU, {cores} threads; PyG itself is not installable here)",
            "ms_per_step": round(1e3 * total / steps, 2), "molecules_per_step": sample}


def gpu_eager_reference(dev, batches, steps, warmup):
    """The "kernel to beat on the same box" (SURVEY.md 2.1 / 8d): the oracle's restatement of the PyG op chains
    (index_select, scatter_add_, scatter_reduce_, addmm, ...) run op by op by stock PyTorch eager ON THE B200 -- what
    the reference would do on this GPU if PyG were installed without extension kernels.  Baseline leg only."""
    import ref_trunks
    from oracle import pyg_oracle as O
    torch.manual_seed(BASE_SEED)
    model = ref_trunks.Model1Trunk(O).to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
    for i in range(warmup):
        train_step(model, opt, batches[i % len(batches)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.reset_peak_memory_stats()
    e0.record()
    for i in range(steps):
        train_step(model, opt, batches[i % len(batches)])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"value": round(BATCH * steps / (ms / 1e3), 1), "unit": UNIT, "ms_per_step": round(ms / steps, 3), "steps": steps,
            "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2),
            "what": "same model, batch, optimiser as `value`, computed by stock PyTorch eager kernels on this GPU from the "
                    "oracle's op-for-op restatement of PyG (ATen index_select / scatter_add_ / scatter_reduce_ / cuBLAS sgemm, "
                    "fp32, TF32 off); PyG itself is not installable here"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    n_steps, n_warm = max(1, args.steps), max(0, args.warmup)
    res = cpu_reference(steps=n_steps, warmup=n_warm, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world,
        "steps": n_steps, "warmup": n_warm, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[2] training step on the reference's CPU path (model1 trunk, Adam 1e-4); "
                               f"each step a bounded sample of {res['molecules_per_step']} of the 4096 molecules",
                   "batch_per_gpu": BATCH},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu-steps", type=int, default=0, help="profiling aid: run N steps inside cudaProfilerStart/Stop and exit")
    ap.add_argument("--call-log", default=None, help="with --ncu-steps: write the ordered list of C-ABI calls (and how many "
                                                      "kernels each launched) of the profiled steps to this JSON file")
    ap.add_argument("--no-activation-fusion", action="store_true", help="keep the models' ReLUs as separate PyTorch launches")
    ap.add_argument("--stress-batch", type=int, default=16384, help="molecules per GPU of the configs[4] leg")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the configs[1] / [3] / [4] legs and the GPU eager baseline")
    args = ap.parse_args()
    if args.impl == "ours" and args.warmup < 3:
        args.warmup = 3                       # timing rules: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
