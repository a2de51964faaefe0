"""CPU: the oracle restatement against (i) an independent dense formulation, (ii) explicit left folds,
(iii) ATen's tie semantics, (iv) the committed golden fixtures.  PyG itself is not installable here
(parity unpinned, see oracle/pyg_oracle.py); these tests are what stands in for it."""
import numpy as np
import pytest
import torch

import ref_trunks
from m_gat_graphsage_b200.data import Data
from m_gat_graphsage_b200.synth import random_graph, synth_batch
from oracle import dense_check as D
from oracle import pyg_oracle as O

torch.manual_seed(0)


def _graphs():
    b = synth_batch(6, 7)
    yield "molecules", b.x, b.edge_index
    x, ei = random_graph(40, 150, 3)
    yield "random-multigraph", x, ei
    x, ei = random_graph(25, 60, 4, self_loops=True)
    yield "pre-existing-self-loops", x, ei
    x, ei = random_graph(10, 0, 5)
    yield "no-edges", x, ei
    x, ei = random_graph(12, 8, 6)           # most atoms isolated
    yield "isolated-atoms", x, ei


@pytest.mark.parametrize("name,x,ei", list(_graphs()), ids=lambda v: v if isinstance(v, str) else None)
def test_sage_matches_dense(name, x, ei):
    conv = O.SAGEConv(35, 17)
    out = conv(x, ei)
    ref = D.sage_dense(x, ei, conv.lin_l.weight, conv.lin_l.bias, conv.lin_r.weight)
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("heads,ch,concat", [(10, 35, True), (1, 128, True), (8, 32, True), (3, 7, False)])
@pytest.mark.parametrize("name,x,ei", list(_graphs()), ids=lambda v: v if isinstance(v, str) else None)
def test_gat_matches_dense(name, x, ei, heads, ch, concat):
    conv = O.GATConv(35, ch, heads=heads, concat=concat)
    with torch.no_grad():
        conv.bias.uniform_(-1, 1)
    out = conv(x, ei)
    ref = D.gat_dense(x, ei, conv.lin.weight, conv.att_src, conv.att_dst, conv.bias, heads, ch, 0.2, concat)
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-5)


def test_gat_sage_gradients_match_dense_fp64():
    x, ei = random_graph(18, 50, 11, self_loops=True)
    x = x.double().requires_grad_(True)
    gat = O.GATConv(35, 6, heads=4).double()
    sage = O.SAGEConv(35, 9).double()
    w = torch.randn(18, 24, dtype=torch.double)
    (gx1,) = torch.autograd.grad((gat(x, ei) * w).sum(), x)
    (gx2,) = torch.autograd.grad(
        (D.gat_dense(x, ei, gat.lin.weight, gat.att_src, gat.att_dst, gat.bias, 4, 6) * w).sum(), x)
    assert torch.allclose(gx1, gx2, rtol=1e-9, atol=1e-11)
    w2 = torch.randn(18, 9, dtype=torch.double)
    (gs1,) = torch.autograd.grad((sage(x, ei) * w2).sum(), x)
    (gs2,) = torch.autograd.grad(
        (D.sage_dense(x, ei, sage.lin_l.weight, sage.lin_l.bias, sage.lin_r.weight) * w2).sum(), x)
    assert torch.allclose(gs1, gs2, rtol=1e-9, atol=1e-11)


def test_gat_gradcheck_fp64():
    x, ei = random_graph(7, 14, 2)
    gat = O.GATConv(5, 3, heads=2).double()
    xs = torch.randn(7, 5, dtype=torch.double, requires_grad=True)
    assert torch.autograd.gradcheck(lambda t: gat(t, ei), (xs,), eps=1e-6, atol=1e-5)


@pytest.mark.parametrize("kind", ["max", "mean", "add"])
def test_pools_match_loop(kind):
    x = torch.randn(50, 9)
    batch = torch.sort(torch.randint(0, 7, (50,)))[0]
    batch[batch == 3] = 4                      # molecule 3 is empty
    fn = {"max": O.global_max_pool, "mean": O.global_mean_pool, "add": O.global_add_pool}[kind]
    out = fn(x, batch, 8)                      # molecule 7 may be empty too
    ref = D.pool_loop(x, batch, 8, kind)
    assert torch.allclose(out, ref, rtol=1e-6, atol=1e-6)
    assert torch.equal(out[3], torch.zeros(9))


def test_scatter_add_is_ascending_left_fold():
    """CPU scatter_add_ folds in ascending row order from 0.0 -- the order the CUDA
    aggregation kernels reproduce bit-exactly."""
    g = torch.Generator().manual_seed(1)
    src = torch.randn(400, 3, generator=g) * 1e3
    idx = torch.randint(0, 20, (400,), generator=g)
    out = O.scatter(src, idx, 20, "sum")
    ref = torch.zeros(20, 3)
    for r in range(400):
        ref[idx[r]] = ref[idx[r]] + src[r]
    assert torch.equal(out, ref)


def test_max_pool_tie_gradient_semantics():
    """ATen scatter_reduce('amax') backward: even split over exact ties; the zero-initialised
    destination counts as one more tie when the max is exactly 0."""
    x = torch.tensor([[1.0, 0.0], [1.0, 0.0], [0.5, 0.0], [2.0, -1.0]], requires_grad=True)
    batch = torch.tensor([0, 0, 0, 1])
    out = O.global_max_pool(x, batch, 2)
    out.sum().backward()
    exp = torch.tensor([[0.5, 0.25], [0.5, 0.25], [0.0, 0.25], [1.0, 1.0]])
    assert torch.equal(x.grad, exp)


def test_csr_oracle_matches_naive():
    b = synth_batch(5, 3)
    N = b.x.size(0)
    c = O.csr_oracle(b.edge_index, N)
    src, dst = b.edge_index.numpy()
    for i in range(N):
        ids = [e for e in range(len(dst)) if dst[e] == i]
        seg = slice(c["rowptr"][i], c["rowptr"][i + 1])
        assert list(c["perm"][seg]) == ids
        assert list(c["col"][seg]) == [src[e] for e in ids]
    assert np.array_equal(c["perm"][c["csc_pos"]], c["permt"])
    assert np.array_equal(O.graph_ptr_oracle(b.batch, 5), b.ptr.numpy().astype(np.int32))


@pytest.mark.parametrize("name", ["model1", "gat", "graphsage", "train"])
def test_oracle_reproduces_golden(name):
    """The fixtures were produced by the REFERENCE's model classes (ast-extracted) over the oracle ops;
    re-deriving them from ref_trunks + oracle must give the same numbers (guards both against drift)."""
    from pathlib import Path
    fx = torch.load(Path(__file__).parent / "golden" / f"{name}.pt", weights_only=False)
    torch.set_num_threads(1)
    model = ref_trunks.build_trunk(name, O, seed=fx["weights_seed"]).eval()
    for k, v in fx["state_checksum"].items():
        assert abs(float(model.state_dict()[k].double().abs().sum()) - v) <= 1e-9 * max(1.0, abs(v)), k
    d = Data(x=fx["x"], edge_index=fx["edge_index"], batch=fx["batch"])
    out = model(d)
    assert torch.allclose(out, fx["logits"], rtol=1e-6, atol=1e-7)
    imp = ref_trunks.atom_importance(model, d)
    assert torch.allclose(imp, fx["atom_importance"], rtol=1e-5, atol=1e-8)


def test_modified_gat_attention_restatement_matches_the_layer_code():
    """train.py:87-99: the dense restatement (and the folded single projection the CUDA path uses) reproduce
    the layer as the reference wrote it, conv1d-on-length-1 and broadcasting matmul included."""
    from m_gat_graphsage_b200.attention import folded_projection
    torch.manual_seed(3)
    layer = ref_trunks.ModifiedGATLayer(35, 35).double()
    x = torch.randn(57, 35, dtype=torch.float64)
    want = layer(x)
    q, k, v = layer.query_transform(x), layer.key_transform(x), layer.value_transform(x)
    k3 = k.unsqueeze(2)
    k_new = layer.linear_transform(torch.cat((layer.conv3(k3), layer.conv5(k3), k3), 1).transpose(1, 2)).squeeze(1)
    assert torch.allclose(O.modified_gat_attention(q, k_new, v), want, rtol=0, atol=1e-12)
    w, b = folded_projection(layer)
    y = x @ w.t() + b
    assert torch.allclose(y[:, :35], q, atol=1e-12) and torch.allclose(y[:, 70:], v, atol=1e-12)
    assert torch.allclose(y[:, 35:70], k_new, atol=1e-12)
    # molecule-restricted softmax == running the layer one molecule at a time (test.py:185-190)
    batch = torch.tensor([0] * 20 + [1] * 1 + [2] * 36)
    per_mol = torch.cat([layer(x[batch == g]) for g in range(3)])
    assert torch.allclose(O.modified_gat_attention(q, k_new, v, batch), per_mol, atol=1e-12)


def test_gcn_and_gin_oracle_match_dense_adjacency():
    """gnn/gcn.py / gnn/gin.py operators: the scatter restatement == dense D^-1/2 (A + I) D^-1/2 X W / (A + I) X."""
    torch.manual_seed(0)
    b = synth_batch(6, 3)
    n = b.x.size(0)
    x = torch.randn(n, 35, dtype=torch.float64)
    a = torch.zeros(n, n, dtype=torch.float64)
    a[b.edge_index[1], b.edge_index[0]] = 1.0                     # a[i, j] = 1 for an edge j -> i
    gcn = O.GCNConv(35, 20).double()
    a_hat = a + torch.eye(n, dtype=torch.float64)
    dinv = a_hat.sum(1).pow(-0.5)
    want = (dinv[:, None] * a_hat * dinv[None, :]) @ (x @ gcn.lin.weight.t()) + gcn.bias
    assert torch.allclose(gcn(x, b.edge_index), want, atol=1e-12)
    imp = O.GCNConv(35, 20, improved=True).double()
    a2 = a + 2 * torch.eye(n, dtype=torch.float64)
    d2 = a2.sum(1).pow(-0.5)
    assert torch.allclose(imp(x, b.edge_index), (d2[:, None] * a2 * d2[None, :]) @ (x @ imp.lin.weight.t()) + imp.bias,
                          atol=1e-12)
    mlp = torch.nn.Sequential(torch.nn.Linear(35, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32)).double()
    gin = O.GINConv(mlp)
    assert torch.allclose(gin(x, b.edge_index), mlp((a + torch.eye(n, dtype=torch.float64)) @ x), atol=1e-12)
    assert sorted(gcn.state_dict()) == ["bias", "lin.weight"] and "eps" in gin.state_dict()


def test_oracle_matches_real_pyg_dump():
    """If a maintainer has run ``tools/dump_pyg_golden.py`` on a machine WITH torch_geometric, the oracle is compared
    with the real PyG outputs and gradients stored under tests/golden/pyg/ (this is what would lift "parity
    unpinned"); in this repository's containers PyG cannot be installed, so the files are absent and the test skips."""
    from pathlib import Path
    import pytest
    d = Path(__file__).parent / "golden" / "pyg"
    files = sorted(d.glob("*.pt")) if d.is_dir() else []
    if not files:
        pytest.skip("tests/golden/pyg/ is empty: torch_geometric is not installable here (tools/dump_pyg_golden.py)")
    for f in files:
        fx = torch.load(f, weights_only=False)
        if f.name == "pools.pt":
            for pname, want in fx["pools"].items():
                x = fx["x"].clone().requires_grad_(True)
                out = getattr(O, pname)(x, fx["batch"])
                assert torch.equal(out, want["out"]), pname
                assert torch.equal(torch.autograd.grad(out.sum(), x)[0], want["x_grad"]), pname
            continue
        conv = getattr(O, fx["layer"])(**fx["kwargs"])
        conv.load_state_dict(fx["state_dict"], strict=True)
        x = fx["x"].clone().requires_grad_(True)
        out = conv(x, fx["edge_index"])
        assert torch.allclose(out, fx["out"], rtol=1e-6, atol=1e-6), f.name
        grads = torch.autograd.grad((out * fx["cotangent"]).sum(), [x] + list(conv.parameters()))
        assert torch.allclose(grads[0], fx["x_grad"], rtol=1e-5, atol=1e-6), f.name
        for (k, _), g in zip(conv.named_parameters(), grads[1:]):
            assert torch.allclose(g, fx["param_grads"][k], rtol=1e-5, atol=1e-5), (f.name, k)
