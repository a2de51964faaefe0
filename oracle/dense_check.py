"""Independent dense-adjacency formulation of the same operators.

TEST INFRASTRUCTURE (see ``oracle/pyg_oracle.py`` header).  Used only to
cross-check the scatter-based restatement, because ``torch_geometric`` itself
cannot be imported here.  Nothing below shares code with ``pyg_oracle``:
SAGE is a row-normalised adjacency matmul, GAT is a masked dense softmax over
``A + I``, pools are Python loops over graphs.
Intended for small N only (O(N^2) memory).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def dense_adj(edge_index: torch.Tensor, n: int, dtype=torch.float32) -> torch.Tensor:
    """A[i, j] = multiplicity of edge j -> i (target-major)."""
    a = torch.zeros(n, n, dtype=dtype)
    ones = torch.ones(edge_index.size(1), dtype=dtype)
    a.index_put_((edge_index[1], edge_index[0]), ones, accumulate=True)
    return a


def sage_dense(x, edge_index, w_l, b_l, w_r):
    a = dense_adj(edge_index, x.size(0), x.dtype)
    deg = a.sum(dim=1, keepdim=True).clamp(min=1)
    agg = (a @ x) / deg
    out = agg @ w_l.t()
    if b_l is not None:
        out = out + b_l
    if w_r is not None:
        out = out + x @ w_r.t()
    return out


def gat_dense(x, edge_index, w, att_src, att_dst, bias, heads, out_channels,
              negative_slope=0.2, concat=True):
    n = x.size(0)
    xh = (x @ w.t()).view(n, heads, out_channels)
    a_s = (xh * att_src.view(1, heads, out_channels)).sum(-1)          # [n, H]
    a_d = (xh * att_dst.view(1, heads, out_channels)).sum(-1)
    adj = dense_adj(edge_index, n, x.dtype)
    adj.fill_diagonal_(0)                       # remove_self_loops
    adj = adj + torch.eye(n, dtype=x.dtype)     # add_self_loops (multiplicity kept for multi-edges)
    # e[i, j, h] = leaky(a_s[j] + a_d[i])
    e = F.leaky_relu(a_s.unsqueeze(0) + a_d.unsqueeze(1), negative_slope)   # [i, j, H]
    mask = (adj > 0).unsqueeze(-1)
    e_m = e.masked_fill(~mask, float("-inf"))
    m = e_m.max(dim=1, keepdim=True)[0]
    p = (e_m - m).exp() * adj.unsqueeze(-1)     # multiplicity-weighted
    alpha = p / (p.sum(dim=1, keepdim=True) + 1e-16)
    out = torch.einsum("ijh,jhc->ihc", alpha, xh)
    out = out.reshape(n, heads * out_channels) if concat else out.mean(dim=1)
    if bias is not None:
        out = out + bias
    return out


def pool_loop(x, batch, num_graphs, kind):
    rows = []
    for g in range(num_graphs):
        xs = x[batch == g]
        if xs.size(0) == 0:
            rows.append(x.new_zeros(x.size(1)))
        elif kind == "max":
            rows.append(xs.max(dim=0)[0])
        elif kind == "mean":
            rows.append(xs.mean(dim=0))
        else:
            rows.append(xs.sum(dim=0))
    return torch.stack(rows)
