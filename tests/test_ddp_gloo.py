"""CPU, world_size 2, gloo: the N > 1 host logic of bench.py / SURVEY.md section 8(e) -- molecules are sharded
by rank (disjoint seeds), every rank runs forward/backward on its own shard, DDP all-reduces (averages) the
gradients, timing is the max over ranks.  Runs on the oracle operators (the CUDA operators refuse CPU tensors);
what is under test is the sharding + collective wiring, not the kernels."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str) -> None:
    sys.path.insert(0, str(ROOT))
    import ref_trunks
    from m_gat_graphsage_b200.data import Data
    from m_gat_graphsage_b200.synth import batch_seed, synth_batch
    from oracle import pyg_oracle as O
    from torch.nn.parallel import DistributedDataParallel as DDP

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    model = ref_trunks.build_trunk("graphsage", O, seed=42).train()
    for m in model.modules():                       # dropout off: both arms must see the same function
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    model.forward = lambda data, f=model.forward: f(data)
    shards = [synth_batch(24, batch_seed(42, r, 0)) for r in range(world)]
    assert not torch.equal(shards[0].x[:50], shards[1].x[:50]), "ranks must draw disjoint molecule streams"

    def loss_of(net, b):
        torch.manual_seed(0)                        # same F.dropout masks in both arms (p=0.2 in SAGENet.forward)
        return F.mse_loss(net(Data(x=b.x, edge_index=b.edge_index, batch=b.batch)).view(-1), b.y)

    # reference: every shard on one process, gradients averaged by hand
    ref_grads = None
    for b in shards:
        model.zero_grad()
        loss_of(model, b).backward()
        g = [p.grad.clone() for p in model.parameters()]
        ref_grads = g if ref_grads is None else [a + c for a, c in zip(ref_grads, g)]
    ref_grads = [g / world for g in ref_grads]

    ddp = DDP(model)
    model.zero_grad()
    loss_of(ddp, shards[rank]).backward()           # DDP all-reduce (mean) happens here
    for p, g in zip(model.parameters(), ref_grads):
        assert torch.allclose(p.grad, g, rtol=1e-5, atol=1e-7), "DDP gradient != mean of per-shard gradients"

    # bench.py's timing rule: max over ranks; value = molecules of ALL ranks / that time
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t) == float(world)
    dist.barrier()
    Path(out_dir, f"ok{rank}").write_text("ok")
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_gradient_average(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
