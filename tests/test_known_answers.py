"""Known-answer vectors worked out BY HAND from PyG's published operator definitions on a
4-atom molecule -- every expected number below is a fraction a reader can verify with pencil and paper; nothing here
is produced by ``oracle/`` or by the CUDA kernels.  They pin the oracle (CPU, runs everywhere) and the CUDA operators
(``-m gpu``) to something other than each other; what they cannot pin is PyG's floating-point summation order,
which the bit-exactness tests derive from ATen's ``scatter_add_`` (tests/test_oracle.py).

Molecule: a star, atom 1 bonded to atoms 0, 2 and 3.  Directed edges sorted by (source, target) like the
reference's ``adj.nonzero()`` (train.py:47-54):  (0,1) (1,0) (1,2) (1,3) (2,1) (3,1).
In-neighbours: N(0) = {1}, N(1) = {0, 2, 3}, N(2) = {1}, N(3) = {1}.
"""
import math

import pytest
import torch

from oracle import pyg_oracle as O

EI = torch.tensor([[0, 1, 1, 1, 2, 3],
                   [1, 0, 2, 3, 1, 1]])
X = torch.tensor([[1., 2.], [3., 4.], [5., 6.], [7., 9.]])
LN2 = math.log(2.0)


def _ops(kind):
    if kind == "oracle":
        return O, torch.device("cpu")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from m_gat_graphsage_b200 import nn as mnn
    return mnn, torch.device("cuda:0")


KINDS = ["oracle", pytest.param("cuda", marks=pytest.mark.gpu)]


def _close(got, want, tol=2e-6):
    want = torch.tensor(want, dtype=torch.float64)
    got = got.detach().cpu().double()
    assert got.shape == want.shape, (got.shape, want.shape)
    err = float((got - want).abs().max()) / max(float(want.abs().max()), 1.0)
    assert err <= tol, f"max error {err:.2e}\ngot  {got.tolist()}\nwant {want.tolist()}"


# ------------------------------------------------------------------------------------------------ SAGEConv (A.2)
@pytest.mark.parametrize("kind", KINDS)
def test_sageconv_known_answer_forward_and_backward(kind):
    """out_i = W_l mean_{j in N(i)} x_j + b_l + W_r x_i with W_l = [[1,0],[0,2]], b_l = [1/2,-1/2], W_r = [[1,1],[0,1]].

    mean: agg_0 = x_1 = [3,4]; agg_1 = (x_0+x_2+x_3)/3 = [13/3, 17/3]; agg_2 = agg_3 = [3,4]
    W_l agg: [3,8]; [13/3, 34/3]; [3,8]; [3,8]          W_r x: [3,2]; [7,4]; [11,6]; [16,9]
    Backward of sum(out): g W_l = [1,2], g W_r = [1,2] for every atom;
      d x_j = [1,2] + sum_{i : j in N(i)} [1,2] / |N(i)|  ->  atoms 0,2,3: [1,2](1 + 1/3); atom 1: [1,2](1 + 3)
      d W_l[o,:] = sum_i agg_i = [40/3, 53/3]; d W_r[o,:] = sum_i x_i = [16, 21]; d b_l = [4, 4]."""
    ops, dev = _ops(kind)
    conv = ops.SAGEConv(2, 2)
    with torch.no_grad():
        conv.lin_l.weight.copy_(torch.tensor([[1., 0.], [0., 2.]]))
        conv.lin_l.bias.copy_(torch.tensor([0.5, -0.5]))
        conv.lin_r.weight.copy_(torch.tensor([[1., 1.], [0., 1.]]))
    conv = conv.to(dev)
    x = X.clone().to(dev).requires_grad_(True)
    out = conv(x, EI.to(dev))
    _close(out, [[6.5, 9.5], [13 / 3 + 7.5, 34 / 3 + 3.5], [14.5, 13.5], [19.5, 16.5]])
    out.sum().backward()
    _close(x.grad, [[4 / 3, 8 / 3], [4., 8.], [4 / 3, 8 / 3], [4 / 3, 8 / 3]])
    _close(conv.lin_l.weight.grad, [[40 / 3, 53 / 3], [40 / 3, 53 / 3]])
    _close(conv.lin_r.weight.grad, [[16., 21.], [16., 21.]])
    _close(conv.lin_l.bias.grad, [4., 4.])


@pytest.mark.parametrize("kind", KINDS)
def test_sageconv_isolated_atom_gets_zero_mean(kind):
    """An atom without in-edges: mean over the empty set is 0 (scatter 'mean' divides by clamp(count, 1)), so
    out = b_l + W_r x.  Two atoms, one edge 0 -> 1: agg_0 = 0, agg_1 = x_0."""
    ops, dev = _ops(kind)
    conv = ops.SAGEConv(2, 2)
    with torch.no_grad():
        conv.lin_l.weight.copy_(torch.tensor([[1., 0.], [0., 2.]]))
        conv.lin_l.bias.copy_(torch.tensor([0.5, -0.5]))
        conv.lin_r.weight.copy_(torch.tensor([[1., 1.], [0., 1.]]))
    conv = conv.to(dev)
    out = conv(X[:2].to(dev), torch.tensor([[0], [1]], device=dev))
    _close(out, [[0.5 + 3., -0.5 + 2.], [1. + 0.5 + 7., 4. - 0.5 + 4.]])


# ------------------------------------------------------------------------------------------------ GATConv (A.1)
def _gat(ops, dev, heads, channels, w, att_src, att_dst, bias, **kw):
    conv = ops.GATConv(2, channels, heads=heads, **kw)
    with torch.no_grad():
        conv.lin.weight.copy_(torch.tensor(w))
        conv.att_src.copy_(torch.tensor(att_src).view(1, heads, channels))
        conv.att_dst.copy_(torch.tensor(att_dst).view(1, heads, channels))
        conv.bias.copy_(torch.tensor(bias))
    return conv.to(dev)


@pytest.mark.parametrize("kind", KINDS)
def test_gatconv_uniform_attention_known_answer(kind):
    """att_src = att_dst = 0: every logit is leaky_relu(0) = 0, so alpha is uniform over N(i) + the self loop PyG
    appends: out_i = mean_{j in N(i) u {i}} x_j W^T + bias with W = I, bias = [1/4, -1].
      out_0 = (x_1+x_0)/2 = [2,3]; out_1 = (x_0+x_2+x_3+x_1)/4 = [4, 21/4]; out_2 = (x_1+x_2)/2 = [4,5];
      out_3 = (x_1+x_3)/2 = [5, 13/2].
    Backward of sum(out) w.r.t. x (alpha does not depend on x here): d x_j = sum over the softmaxes j takes part in of
    1/(|N(i)|+1): atoms 0,2,3: 1/2 (own) + 1/4 (atom 1's) = 3/4; atom 1: 1/4 + 3/2 = 7/4."""
    ops, dev = _ops(kind)
    conv = _gat(ops, dev, 1, 2, [[1., 0.], [0., 1.]], [0., 0.], [0., 0.], [0.25, -1.0])
    x = X.clone().to(dev).requires_grad_(True)
    out = conv(x, EI.to(dev))
    _close(out, [[2.25, 2.], [4.25, 4.25], [4.25, 4.], [5.25, 5.5]])
    out.sum().backward()
    _close(x.grad, [[0.75, 0.75], [1.75, 1.75], [0.75, 0.75], [0.75, 0.75]])


@pytest.mark.parametrize("kind", KINDS)
def test_gatconv_softmax_known_answer_positive_and_negative_logits(kind):
    """One head, one channel, W = [[1, 0]] so xh = x[:,0], att_dst = 0, bias = 0.

    (a) att_src = ln 2, x[:,0] = [1,3,5,7]: logit of edge j->i = leaky_relu(xh_j ln 2) = xh_j ln 2 (positive branch),
        exp = 2^xh_j = 2, 8, 32, 128.
          out_0 = (8*3 + 2*1)/10 = 13/5;              out_1 = (2*1 + 32*5 + 128*7 + 8*3)/170 = 1082/170;
          out_2 = (8*3 + 32*5)/40 = 23/5;             out_3 = (8*3 + 128*7)/136 = 920/136.
    (b) att_src = -ln 2, x[:,0] = [5,10,5,15]: logit = leaky_relu(-xh_j ln 2) = -0.2 xh_j ln 2 (negative branch, slope
        0.2), exp = 2^(-xh_j/5) = 1/2, 1/4, 1/2, 1/8.
          out_0 = (10/4 + 5/2)/(3/4) = 20/3;          out_1 = (5/2 + 5/2 + 15/8 + 10/4)/(11/8) = 75/11;
          out_2 = (10/4 + 5/2)/(3/4) = 20/3;          out_3 = (10/4 + 15/8)/(3/8) = 35/3."""
    ops, dev = _ops(kind)
    conv = _gat(ops, dev, 1, 1, [[1., 0.]], [LN2], [0.], [0.])
    _close(conv(X.to(dev), EI.to(dev)), [[13 / 5], [1082 / 170], [23 / 5], [920 / 136]])
    conv = _gat(ops, dev, 1, 1, [[1., 0.]], [-LN2], [0.], [0.])
    xb = torch.tensor([[5., 0.], [10., 0.], [5., 0.], [15., 0.]])
    _close(conv(xb.to(dev), EI.to(dev)), [[20 / 3], [75 / 11], [20 / 3], [35 / 3]])


@pytest.mark.parametrize("kind", KINDS)
def test_gatconv_two_heads_concat_and_mean(kind):
    """Two heads with one channel each, W = [[1,0],[0,1]] (head 0 sees x[:,0], head 1 sees x[:,1]); head 0 has
    att_src = ln 2 (answer (a) above), head 1 has zero attention vectors (uniform mean of x[:,1] = [2,4,6,9]:
    3, 21/4, 5, 13/2).  concat=True: [head0 | head1] + bias [H*C]; concat=False: mean over heads + bias [C]."""
    ops, dev = _ops(kind)
    h0, h1 = [13 / 5, 1082 / 170, 23 / 5, 920 / 136], [3., 21 / 4, 5., 13 / 2]
    conv = _gat(ops, dev, 2, 1, [[1., 0.], [0., 1.]], [LN2, 0.], [0., 0.], [1., -1.])
    _close(conv(X.to(dev), EI.to(dev)), [[a + 1., b - 1.] for a, b in zip(h0, h1)])
    conv = _gat(ops, dev, 2, 1, [[1., 0.], [0., 1.]], [LN2, 0.], [0., 0.], [0.5], concat=False)
    _close(conv(X.to(dev), EI.to(dev)), [[(a + b) / 2 + 0.5] for a, b in zip(h0, h1)])


@pytest.mark.parametrize("kind", KINDS)
def test_gatconv_pre_existing_self_loops_are_replaced_not_doubled(kind):
    """PyG removes self loops that are already in edge_index and appends exactly one per atom: with uniform attention
    a two-atom molecule (0 <-> 1) WITH an explicit (0,0) edge still gives out_0 = (x_0 + x_1)/2, not (2 x_0 + x_1)/3."""
    ops, dev = _ops(kind)
    conv = _gat(ops, dev, 1, 2, [[1., 0.], [0., 1.]], [0., 0.], [0., 0.], [0., 0.])
    ei = torch.tensor([[0, 0, 1], [0, 1, 0]], device=dev)
    _close(conv(X[:2].to(dev), ei), [[2., 3.], [2., 3.]])


# ------------------------------------------------------------------------------------------------ pools (A.3)
@pytest.mark.parametrize("kind", KINDS)
def test_pools_known_answer_with_an_empty_molecule(kind):
    """batch = [0,0,1,1], three molecules (the last one empty): max / mean / add per column; empty -> 0."""
    ops, dev = _ops(kind)
    x = torch.tensor([[1., -2.], [3., 4.], [-5., 6.], [7., -9.]], device=dev)
    batch = torch.tensor([0, 0, 1, 1], device=dev)
    _close(ops.global_max_pool(x, batch, 3), [[3., 4.], [7., 6.], [0., 0.]])
    _close(ops.global_mean_pool(x, batch, 3), [[2., 1.], [1., -1.5], [0., 0.]])
    _close(ops.global_add_pool(x, batch, 3), [[4., 2.], [2., -3.], [0., 0.]])


@pytest.mark.parametrize("kind", KINDS)
def test_max_pool_gradient_splits_evenly_over_exact_ties(kind):
    """ATen's ``scatter_reduce_('amax')`` backward (the reference's CPU path, A.3): the gradient is divided evenly over
    the sources that equal the result -- and the zero-initialised destination counts as one more tie when the maximum
    is exactly 0.  Molecule 0: values [2, 2, 1] -> 1/2, 1/2, 0.  Molecule 1: values [0, 0, 0] -> 1/4 each (three
    sources + the destination).  Molecule 2: values [-1, 0] -> 0, 1/2."""
    ops, dev = _ops(kind)
    x = torch.tensor([[2.], [2.], [1.], [0.], [0.], [0.], [-1.], [0.]], device=dev, requires_grad=True)
    batch = torch.tensor([0, 0, 0, 1, 1, 1, 2, 2], device=dev)
    out = ops.global_max_pool(x, batch, 3)
    _close(out, [[2.], [0.], [0.]])
    out.sum().backward()
    _close(x.grad, [[0.5], [0.5], [0.], [0.25], [0.25], [0.25], [0.], [0.5]])


@pytest.mark.parametrize("kind", KINDS)
def test_mean_pool_gradient(kind):
    ops, dev = _ops(kind)
    x = torch.ones(5, 2, device=dev, requires_grad=True)
    batch = torch.tensor([0, 0, 0, 1, 1], device=dev)
    (ops.global_mean_pool(x, batch, 2) * torch.tensor([[3., 6.], [2., 4.]], device=dev)).sum().backward()
    _close(x.grad, [[1., 2.]] * 3 + [[1., 2.]] * 2)


# ------------------------------------------------------------------------------------------------ K0 (row a2)
@pytest.mark.parametrize("kind", KINDS)
def test_csr_known_answer(kind):
    """Sorted CSR by destination of the star: in-edges of atom 1 are edges 0, 4, 5 (ascending edge id = stable sort)."""
    want = dict(rowptr=[0, 1, 4, 5, 6], col=[1, 0, 2, 3, 1, 1], perm=[1, 0, 4, 5, 2, 3],
                colptr=[0, 1, 4, 5, 6], row=[1, 0, 2, 3, 1, 1], permt=[0, 1, 2, 3, 4, 5])
    if kind == "oracle":
        got = O.csr_oracle(EI, 4)
        for k, v in want.items():
            assert got[k].tolist() == v, k
        assert O.graph_ptr_oracle(torch.tensor([0, 0, 2, 2, 2]), 4).tolist() == [0, 2, 2, 5, 5]
        return
    _, dev = _ops(kind)
    from m_gat_graphsage_b200.graph import build_graph_index, graph_ptr
    gi = build_graph_index(EI.to(dev), 4)
    for k, v in want.items():
        assert getattr(gi, k).cpu().tolist() == v, k
    assert graph_ptr(torch.tensor([0, 0, 2, 2, 2], device=dev), 4).cpu().tolist() == [0, 2, 2, 5, 5]
