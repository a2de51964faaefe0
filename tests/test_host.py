"""CPU: host-side logic -- Data / Batch / DataLoader collation (row a1), the synthetic generator, the
torch_geometric import shim, the C-ABI library surface, and loud failure without CUDA."""
import ctypes
import re
import sys
from pathlib import Path

import pytest
import torch

from m_gat_graphsage_b200 import _lib
from m_gat_graphsage_b200.data import Batch, Data, DataLoader
from m_gat_graphsage_b200.synth import MAX_ATOMS, MIN_ATOMS, batch_seed, synth_batch
from oracle import pyg_oracle as O

ROOT = Path(__file__).resolve().parents[1]


def _molecule(n, e, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 35, generator=g)
    ei = torch.randint(0, n, (2, e), generator=g)
    return x, ei


def test_collate_matches_oracle():
    mols = [_molecule(n, e, s) for s, (n, e) in enumerate([(5, 8), (1, 0), (12, 30), (3, 4)])]
    datas = []
    for k, (x, ei) in enumerate(mols):
        d = Data(x=x, edge_index=ei)
        d.y = torch.tensor(float(k))           # 0-dim target, train.py:190
        d.y_original = torch.tensor(float(10 * k))
        datas.append(d)
    b = Batch.from_data_list(datas)
    x, ei, bv, ptr = O.collate_oracle(mols)
    assert torch.equal(b.x, x) and torch.equal(b.edge_index, ei)
    assert torch.equal(b.batch, bv) and torch.equal(b.ptr, ptr)
    assert b.num_graphs == 4 and b.y.shape == (4,) and b.y_original.tolist() == [0.0, 10.0, 20.0, 30.0]
    assert getattr(b.batch, "_mgs_num_graphs") == 4
    back = b.to_data_list()
    assert all(torch.equal(d.x, m[0]) and torch.equal(d.edge_index, m[1]) for d, m in zip(back, mols))


def test_dataloader_tuple_dataset_like_train_py():
    """train.py:192,209: list of (Data, ecfp[1,1024]) tuples -> [Batch, Tensor[B,1,1024]]."""
    items = []
    for k in range(7):
        x, ei = _molecule(4 + k, 6, k)
        d = Data(x=x, edge_index=ei)
        d.y = torch.tensor(float(k))
        items.append((d, torch.full((1, 1024), float(k))))
    loader = DataLoader(items, batch_size=3, shuffle=False)
    assert len(loader) == 3
    batches = list(loader)
    bd, ecfp = batches[0]
    assert isinstance(bd, Batch) and ecfp.shape == (3, 1, 1024)
    assert bd.y.view(-1, 1).shape == (3, 1)
    assert batches[-1][0].num_graphs == 1
    shuffled = DataLoader(items, batch_size=7, shuffle=True)
    (bd2, e2), = list(shuffled)
    assert sorted(e2[:, 0, 0].tolist()) == [float(k) for k in range(7)]


def _loader_items(count, with_extra):
    items = []
    for k in range(count):
        n = 1 + (k * 7) % 13
        x, ei = _molecule(n, 0 if k % 5 == 1 else 2 * n, k)     # some molecules without bonds (single atoms)
        d = Data(x=x, edge_index=ei)
        d.y = torch.tensor(float(k))                             # 0-dim -> stacked [B]
        d.y_original = torch.tensor([float(-k)])                 # [1] -> concatenated [B]
        d.edge_attr = torch.full((ei.size(1), 3), float(k))      # per-bond rows follow edge_index
        items.append((d, torch.full((1, 8), float(k))) if with_extra else d)
    return items


@pytest.mark.parametrize("with_extra", [False, True])
@pytest.mark.parametrize("batch_size,shuffle,drop_last", [(1, False, False), (6, True, False), (6, True, True), (64, False, False)])
def test_dataloader_flat_gather_is_bit_identical_to_python_collation(with_extra, batch_size, shuffle, drop_last):
    """row a1: the pre-collated gather path yields the very batches Batch.from_data_list builds."""
    items = _loader_items(41, with_extra)
    mk = lambda fast: DataLoader(items, batch_size=batch_size, shuffle=shuffle, drop_last=drop_last, fast=fast,
                                 generator=torch.Generator().manual_seed(3))
    slow, fast_loader = mk(False), mk(True)
    a, b = list(slow), list(fast_loader)
    assert fast_loader._flat is not None, "fast path must be the one that ran"
    assert len(a) == len(b) == len(slow)
    for s_, f_ in zip(a, b):
        if with_extra:
            assert torch.equal(s_[1], f_[1])
            s_, f_ = s_[0], f_[0]
        assert isinstance(f_, Batch) and s_.keys() == f_.keys() and s_.num_graphs == f_.num_graphs
        for k in s_.keys():
            assert s_[k].dtype == f_[k].dtype and torch.equal(s_[k], f_[k]), k
        assert getattr(f_.batch, "_mgs_num_graphs") == f_.num_graphs
    # a second epoch reshuffles (new permutation from the same generator state) but stays a permutation
    if shuffle and not drop_last:
        again = list(fast_loader)
        ys = torch.cat([(t[0] if with_extra else t).y for t in again])
        assert sorted(ys.tolist()) == [float(k) for k in range(41)]


def test_dataloader_falls_back_on_non_tensor_attributes():
    items = _loader_items(5, False)
    for k, d in enumerate(items):
        d.smiles = "C" * (k + 1)
    loader = DataLoader(items, batch_size=2)
    out = list(loader)
    assert loader._flat is None and not loader._fast
    assert out[0].smiles == ["C", "CC"]


def test_data_attribute_protocol():
    d = Data(x=torch.zeros(3, 35), edge_index=torch.zeros(2, 0, dtype=torch.long))
    assert d.batch is None and d.y is None and d.num_nodes == 3 and d.num_edges == 0
    d.batch = torch.zeros(3, dtype=torch.long)   # test.py:186
    assert "batch" in d and d.to("cpu") is d and d.cpu() is d
    with pytest.raises(AttributeError):
        d.nonexistent


def test_synth_invariants():
    b = synth_batch(300, 42)
    N = b.x.size(0)
    n = b.ptr[1:] - b.ptr[:-1]
    assert int(n.min()) >= MIN_ATOMS and int(n.max()) <= MAX_ATOMS
    src, dst = b.edge_index
    assert not bool((src == dst).any())
    key = src * N + dst
    assert bool((key[1:] > key[:-1]).all()), "edges must be in adj.nonzero() (row-major) order, no duplicates"
    rev = set(map(tuple, torch.stack([dst, src], 1).tolist()))
    assert rev == set(map(tuple, b.edge_index.t().tolist())), "bonds are symmetric"
    assert torch.equal(b.batch[src], b.batch[dst]), "no bonds between molecules"
    deg = torch.bincount(dst, minlength=N)
    assert int(deg.max()) <= 6
    assert set(b.x.unique().tolist()) <= {0.0, 1.0}
    assert torch.equal(b.x[:, 10:17].argmax(1), deg.clamp(max=6)), "degree one-hot is consistent"
    ratio = b.edge_index.size(1) / N
    assert 1.9 < ratio < 2.3
    b2 = synth_batch(300, 42)
    assert torch.equal(b.x, b2.x) and torch.equal(b.edge_index, b2.edge_index)
    assert batch_seed(42, 0, 0) != batch_seed(42, 1, 0) != batch_seed(42, 0, 1)
    s = synth_batch(4, 1, fixed_atoms=94)
    assert s.x.size(0) == 4 * 94


def test_shim_imports_like_reference_scripts():
    shim = str(ROOT / "m_gat_graphsage_b200" / "shim")
    sys.path.insert(0, shim)
    try:
        for k in [k for k in sys.modules if k == "torch_geometric" or k.startswith("torch_geometric.")]:
            del sys.modules[k]
        from torch_geometric.data import Data as D2, DataLoader as DL2  # noqa: F401  train.py:8
        from torch_geometric.nn import GATConv, SAGEConv, global_max_pool, global_mean_pool as gap  # noqa: F401
        from torch_geometric.explain import Explainer, GNNExplainer  # noqa: F401  gnnexplainer.py:7
        from torch_geometric.explain.config import ExplainerConfig, ModelConfig  # noqa: F401
        sage, gat = SAGEConv(35, 35), GATConv(35, 35, heads=10)
        assert sorted(sage.state_dict()) == ["lin_l.bias", "lin_l.weight", "lin_r.weight"]
        assert sorted(gat.state_dict()) == ["att_dst", "att_src", "bias", "lin.weight"]
        o_sage, o_gat = O.SAGEConv(35, 35), O.GATConv(35, 35, heads=10)
        sage.load_state_dict(o_sage.state_dict(), strict=True)
        gat.load_state_dict(o_gat.state_dict(), strict=True)
    finally:
        sys.path.remove(shim)


def test_operators_reject_cpu_tensors_loudly():
    from m_gat_graphsage_b200 import nn as mnn
    x = torch.randn(4, 35)
    ei = torch.tensor([[0, 1], [1, 0]])
    for call in (lambda: mnn.SAGEConv(35, 8)(x, ei), lambda: mnn.GATConv(35, 8)(x, ei),
                 lambda: mnn.global_max_pool(x, torch.zeros(4, dtype=torch.long)),
                 lambda: mnn.Linear(35, 8)(x)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()


def test_library_exports_every_declared_symbol(lib_built):
    header = (ROOT / "include" / "mgs.h").read_text()
    declared = set(re.findall(r"\b(mgs_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 24
    lib = ctypes.CDLL(str(lib_built))
    for name in declared:
        assert hasattr(lib, name), f"libmgs.so does not export {name}"
    assert declared == set(_lib.SIGNATURES), "ctypes table and include/mgs.h disagree"
    loaded = _lib.load()
    assert loaded.mgs_version() == 100


def test_library_validates_arguments_without_touching_the_gpu(lib_built):
    lib = _lib.load()
    rc = lib.mgs_linear_fwd(0, 0, 4, 0, 0, 0, 8, 0, 0, 0, 0, 0, 0, 0, 8, 0, 0, 0, 0)   # K = 0
    assert rc == 1 and b"bad sizes" in lib.mgs_last_error_string()
    rc = lib.mgs_pool_fwd(0, 4, 0, 2, 4, 7, 0, 4, 0)                              # unknown mode
    assert rc == 1 and b"unknown mode" in lib.mgs_last_error_string()
    assert lib.mgs_csr_workspace_bytes(100, 300) >= 3 * 300 * 4
    assert lib.mgs_linear_wgrad_workspace_bytes(0, 8, 8) == 0


def test_message_passing_base_accepts_pyg_constructor_arguments():
    """gnn/chebnet.py:50-54: `class ChebConv(MessagePassing)` calls super().__init__(aggr='add') and never propagates."""
    import torch.nn as nn
    from m_gat_graphsage_b200.nn import MessagePassing

    class ChebLike(MessagePassing):
        def __init__(self, cin, cout, K, **kwargs):
            super().__init__(aggr="add", **kwargs)
            self.K, self.lin = K, nn.Linear(cin, cout)

        def forward(self, x, edge_index):
            lap = torch.zeros(x.size(0), x.size(0))
            lap[edge_index[0], edge_index[1]] = -1
            lap = lap + torch.diag(lap.sum(1))
            return self.lin(x + lap @ x)

    layer = ChebLike(35, 16, 3)
    x, ei = _molecule(9, 14, 0)
    assert layer(x, ei).shape == (9, 16) and layer.aggr == "add" and sorted(layer.state_dict()) == ["lin.bias", "lin.weight"]


def test_wire_batch_packs_one_bit_per_feature():
    """data.WireBatch: the 0 / 1 atom features as one bit each, int32 edge ids, segment pointers -- decoded on the host
    here (the device expansion is tested on the GPU)."""
    import torch
    from m_gat_graphsage_b200.data import WireBatch
    from m_gat_graphsage_b200.synth import synth_batch
    b = synth_batch(50, 4)
    w = WireBatch.from_batch(b, pin=False)
    assert w.xbits.dtype == torch.int64 and w.edge_index.dtype == torch.int32 and w.ptr.dtype == torch.int32
    x = ((w.xbits.unsqueeze(1) >> torch.arange(35)) & 1).float()
    assert torch.equal(x, b.x) and torch.equal(w.edge_index.long(), b.edge_index) and torch.equal(w.ptr.long(), b.ptr)
    plain = sum(t.numel() * t.element_size() for t in (b.x, b.edge_index, b.batch, b.y))
    assert w.nbytes * 6 < plain
    b.x[3, 5] = 0.5
    with pytest.raises(ValueError, match="0.0 / 1.0"):
        WireBatch.from_batch(b, pin=False)


def test_pooled_halves_node_gradient_paths():
    """nn._PooledHalves (pure autograd, no kernel): the readout's `torch.cat([gmp, gap], dim=1)` hands back two adjacent views
    of one buffer -- returned as the gradient of the fused [B, 2F] result without a copy; any other use (one half only,
    halves consumed separately) falls back to one cat with zeros.  Gradients equal those of plain slicing."""
    from m_gat_graphsage_b200.nn import _PooledHalves
    g0 = torch.Generator().manual_seed(2)
    B, F = 7, 5
    base = torch.randn(B, 2 * F, generator=g0)
    w = torch.randn(B, 2 * F, generator=g0)

    def run(fn):
        both = base.clone().requires_grad_(True)
        fn(both).backward()
        return both.grad

    # (1) the reference readout: cat of both halves
    got = run(lambda b: (torch.cat(_PooledHalves.apply(b * 1.0, F), dim=1) * w).sum())
    want = run(lambda b: (torch.cat([(b * 1.0)[:, :F], (b * 1.0)[:, F:]], dim=1) * w).sum())
    assert torch.equal(got, want)
    # (2) only the max half is used (train.py:119)
    got = run(lambda b: (_PooledHalves.apply(b * 1.0, F)[0] * w[:, :F]).sum())
    want = run(lambda b: ((b * 1.0)[:, :F] * w[:, :F]).sum())
    assert torch.equal(got, want)
    # (3) halves consumed separately, in swapped order (not adjacent in one buffer)
    def swapped(b):
        mx, mean = _PooledHalves.apply(b * 1.0, F)
        return (torch.cat([mean, mx], dim=1) * w).sum()
    want = run(lambda b: (torch.cat([(b * 1.0)[:, F:], (b * 1.0)[:, :F]], dim=1) * w).sum())
    assert torch.equal(run(swapped), want)


def test_fused_adam_refuses_cpu_parameters():
    """accel.FusedAdam has no CPU fallback (the product path fails loudly without the CUDA library / device)."""
    from m_gat_graphsage_b200.accel import FusedAdam
    p = torch.zeros(3, requires_grad=True)
    p.grad = torch.ones(3)
    opt = FusedAdam([p], lr=1e-3)
    with pytest.raises((RuntimeError, _lib.MgsLibraryError)):
        opt.step()
    with pytest.raises(ValueError):
        FusedAdam([p], lr=-1.0)
