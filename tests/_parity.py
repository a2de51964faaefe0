"""Error metrics and tie-class helpers shared by the parity tests.

Two metrics are computed for every comparison and both are asserted (BASELINE.json north_star: logits 1e-5
relative, gradients / importances 1e-4 relative):

* ``rel_max``  = max|a - b| / max|b|                               (error relative to the tensor's scale)
* ``rel_elem`` = max_i |a_i - b_i| / max(|b_i|, floor * max|b|)    (element-wise relative error; ``floor`` is the
  absolute floor below which an element is compared against the floor instead of against itself: entries of an
  fp32 sum that are small because large terms cancelled carry the rounding error of the large terms and have no
  meaningful relative error of their own)

Every call is also appended to ``REPORT`` so that a GPU run can dump the measured numbers
(``gpurun_out/parity_report.json``; a copy is committed under ``profiles/``).
"""
from __future__ import annotations

import json
import os
from pathlib import Path

import torch

REPORT = []
ELEM_FLOOR = 1e-2          # elements below 1 % of the tensor's max are compared against that floor


def _d(t):
    return t.detach().cpu().double()


def rel_max(a, b) -> float:
    a, b = _d(a), _d(b)
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


def rel_elem(a, b, floor: float = ELEM_FLOOR) -> float:
    a, b = _d(a), _d(b)
    scale = max(float(b.abs().max()), 1e-30)
    return float(((a - b).abs() / b.abs().clamp(min=floor * scale)).max())


def check(a, b, tol: float, what: str, elem_factor: float = 10.0, floor: float = ELEM_FLOOR):
    """Assert ``rel_max <= tol`` and ``rel_elem <= elem_factor * tol`` (an element at the floor, i.e. 100x smaller
    than the largest one, may carry 10x the relative error of the tensor as a whole); record both."""
    assert tuple(a.shape) == tuple(b.shape), f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    rm, re = rel_max(a, b), rel_elem(a, b, floor)
    REPORT.append({"what": what, "rel_max": rm, "rel_elem": re, "tol": tol, "numel": int(b.numel())})
    assert rm <= tol, f"{what}: max|diff|/max|ref| = {rm:.3e} > {tol:g}"
    assert re <= elem_factor * tol, f"{what}: element-wise relative error {re:.3e} > {elem_factor * tol:g} (floor {floor:g})"
    return rm, re


def dump_report() -> None:
    out = Path(os.environ.get("GRAFT_REPO_ROOT", Path(__file__).resolve().parents[1])) / "gpurun_out"
    if REPORT and out.is_dir():
        path = out / "parity_report.json"
        old = json.loads(path.read_text()) if path.exists() else []
        path.write_text(json.dumps(old + REPORT, indent=0))
        REPORT.clear()


# ------------------------------------------------------------------------------------------------------------------
# max-pool ties (SURVEY.md section 7 "hard parts")
# ------------------------------------------------------------------------------------------------------------------
def tied_molecules(h: torch.Tensor, batch: torch.Tensor, num_graphs: int, rtol: float = 1e-5,
                   kink: bool = True) -> torch.Tensor:
    """-> bool [B]: molecules in which some column of the pre-pooling embedding ``h`` has a positive maximum that a
    second atom reaches within ``rtol`` (relative to the tensor's scale).  Which of such twins receives the pooled
    gradient hinges on 1-ulp differences that no two GEMM implementations reproduce; everywhere else the arg-max is
    unambiguous and per-atom gradients must agree.  ``kink``: ``h`` is the output of a ReLU, so a maximum within
    ``rtol`` of zero is a near tie with the ReLU's kink (one rounding away the column's maximum is exactly 0, every
    atom ties with the zero-initialised destination and ReLU'(0) = 0 removes the gradient altogether)."""
    h, batch = h.detach().cpu().double(), batch.cpu()
    B, F = num_graphs, h.size(1)
    idx = batch.view(-1, 1).expand_as(h)
    top = torch.full((B, F), float("-inf"), dtype=h.dtype).scatter_reduce_(0, idx, h, "amax", include_self=True)
    is_top = h == top[batch]
    second = torch.full((B, F), float("-inf"), dtype=h.dtype).scatter_reduce_(
        0, idx, h.masked_fill(is_top, float("-inf")), "amax", include_self=True)
    n_top = torch.zeros(B, F, dtype=h.dtype).scatter_add_(0, idx, is_top.double())
    scale = float(h.abs().max())
    near = ((top - second) <= rtol * scale) | (n_top > 1)
    flagged = near & (top > 0)
    if kink:
        flagged |= (top > 0) & (top <= rtol * scale)
    return flagged.any(dim=1)


def kink_molecules(pre_activations, num_graphs: int, rtol: float = 1e-5) -> torch.Tensor:
    """-> bool [B]: molecules with a ReLU pre-activation within ``rtol`` (of that tensor's scale) of zero.
    ``pre_activations``: iterable of ``(tensor [rows, F], molecule id per row [rows])``.  The gradient of a ReLU network
    is discontinuous there: a unit that is "on" in one fp32 implementation and "off" in another moves the molecule's
    input gradient by O(1), between ANY two implementations (also CPU fp32 vs fp64)."""
    out = torch.zeros(num_graphs, dtype=torch.bool)
    for t, rows in pre_activations:
        t = t.detach().cpu()
        hit = (t.abs() <= rtol * float(t.abs().max())).any(dim=1)
        out.index_put_((rows.cpu()[hit],), torch.ones(int(hit.sum()), dtype=torch.bool))
    return out


def feature_classes(x: torch.Tensor, batch: torch.Tensor):
    """-> (class id per atom [N], number of classes): atoms of the SAME molecule with IDENTICAL input feature rows.
    A tie swap between twins moves gradient from atom k to the atom phi(k) that the local isomorphism between the two
    twins' neighbourhoods maps it to, and x[phi(k)] == x[k]: sums of d pred / d x rows over these classes are
    invariant under every such swap (finer than per-molecule sums, no threshold involved)."""
    x, batch = x.detach().cpu(), batch.cpu()
    key = torch.cat([batch.view(-1, 1).to(x.dtype), x], dim=1)
    _, inv = torch.unique(key, dim=0, return_inverse=True)
    return inv, int(inv.max()) + 1 if inv.numel() else 0


def class_sums(g: torch.Tensor, cls: torch.Tensor, n: int) -> torch.Tensor:
    g = g.detach().cpu().double()
    return torch.zeros(n, g.size(1), dtype=g.dtype).index_add_(0, cls, g)


class RecordingOps:
    """Operator namespace that forwards to ``ops`` and remembers the tensor handed to ``global_max_pool`` (the
    pre-pooling embedding the tie analysis needs)."""

    def __init__(self, ops):
        self._ops = ops
        self.pool_input = None

    def __getattr__(self, name):
        return getattr(self._ops, name)

    def global_max_pool(self, x, batch, size=None):
        self.pool_input = x.detach()
        return self._ops.global_max_pool(x, batch, size)


def gat_logits(conv, x_in: torch.Tensor, edge_index: torch.Tensor):
    """Pre-LeakyReLU attention logits ``a_src[j] + a_dst[i]`` of an oracle GATConv for every edge incl. self loops
    -> ``(logits [E', H], destination atom per row)``: LeakyReLU has a kink at 0 like ReLU (slope 1 vs 0.2)."""
    with torch.no_grad():
        N, H, C = x_in.size(0), conv.heads, conv.out_channels
        xh = conv.lin(x_in).view(N, H, C)
        a_s, a_d = (xh * conv.att_src).sum(-1), (xh * conv.att_dst).sum(-1)
        keep = edge_index[0] != edge_index[1]
        loop = torch.arange(N)
        src, dst = torch.cat([edge_index[0][keep], loop]), torch.cat([edge_index[1][keep], loop])
        return a_s[src] + a_d[dst], dst


def smooth_molecules(model, rec_ops, data, num_graphs: int, relu_modules, gat_modules=(), pool_after_relu=True,
                     tie_rtol: float = 1e-5, kink_rtol: float = 2e-6):
    """Run the ORACLE ``model`` (built on ``rec_ops = RecordingOps(oracle)``) on ``data`` and classify molecules:
    -> ``(out, smooth [B], tied_only [B])``.  ``relu_modules``: names of sub-modules whose output feeds a ReLU;
    ``gat_modules``: names of GATConv sub-modules (LeakyReLU kink on the attention logits).  ``smooth`` molecules
    have a unique pooling arg-max in every column and no pre-activation at a kink: per-atom gradients of two fp32
    implementations must agree there.  ``tied_only``: near ties in the max pool but no kink: gradients agree after
    summing over tie classes (``feature_classes``)."""
    pre, gat_in, hooks = {}, {}, []
    for nm in relu_modules:
        hooks.append(getattr(model, nm).register_forward_hook(lambda m, i, o, nm=nm: pre.__setitem__(nm, o.detach())))
    for nm in gat_modules:
        hooks.append(getattr(model, nm).register_forward_hook(
            lambda m, i, o, nm=nm: gat_in.__setitem__(nm, (i[0].detach(), i[1]))))
    out = model(data)
    for h in hooks:
        h.remove()
    batch = data.batch.cpu()
    n_atoms = batch.numel()
    acts = [(t, batch if t.size(0) == n_atoms else torch.arange(num_graphs)) for t in pre.values()]
    for nm, (x_in, ei) in gat_in.items():
        e, dst = gat_logits(getattr(model, nm), x_in, ei)
        acts.append((e, batch[dst]))
    kink = kink_molecules(acts, num_graphs, kink_rtol)
    tied = tied_molecules(rec_ops.pool_input, batch, num_graphs, tie_rtol, kink=pool_after_relu)
    return out, ~(tied | kink), tied & ~kink
