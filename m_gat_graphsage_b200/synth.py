"""Synthetic molecule batches shaped like the reference's featurisation.

The reference ships no data (``README.md:11`` mentions ``data/`` which is
absent); it publishes only the atom-count *range* 11-94 (``README.md:127``).
Everything below that range is an ASSUMPTION and is
repeated in every benchmark report:

* atoms per molecule ``n = clip(round(exp(N(ln 30, 0.35^2))), 11, 94)`` (mean ~31);
  the stress config uses ``n = 94`` for all molecules;
* topology: a random tree (atom k bonds to one of the previous <= 6 atoms,
  degree cap 4; ~1 % "hub" atoms have cap 6 and attract bonds) plus
  ``max(1, n // 12)`` ring-closure bonds of ring size 5-6 where the degree
  cap allows => E/N ~ 2.0-2.2, degree <= 6;
* edges: both directions, no self loops, ordered exactly like
  ``adj.nonzero()`` (row-major, /root/reference/train.py:47-54);
* features: 35 = 10 symbol + 7 degree + 7 implicit valence + 5 hybridisation
  + 1 aromatic + 5 total-Hs one-hots (train.py:33-43), values exactly 0.0 / 1.0,
  the degree one-hot consistent with the generated degree, ~1 % of
  hybridisation groups all-zero (the ``Unknown`` fall-through of train.py:19-22).

The generator is vectorised over molecules and runs on any device, so the
1 M / 4 M molecule configs are created on the GPU outside timed regions.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from .data import Batch, _tag_num_graphs

NUM_FEATURES = 35
MIN_ATOMS, MAX_ATOMS = 11, 94

_SYMBOL_P = [0.72, 0.12, 0.12, 0.02, 0.01, 0.002, 0.005, 0.002, 0.0005, 0.0005]
_VALENCE_P = [0.35, 0.30, 0.20, 0.15, 0.0, 0.0, 0.0]
_HYBRID_P = [0.03, 0.45, 0.49, 0.01, 0.01, 0.01]          # last = "Unknown" -> all-zero group
_HS_P = [0.40, 0.30, 0.18, 0.10, 0.02]


def _categorical(p, n, gen, device):
    probs = torch.tensor(p, dtype=torch.float32, device=device)
    return torch.multinomial(probs, n, replacement=True, generator=gen)


def synth_batch(num_graphs: int, seed: int, device="cpu", fixed_atoms: Optional[int] = None,
                with_targets: bool = True) -> Batch:
    """One collated batch of ``num_graphs`` synthetic molecules (deterministic in ``seed``)."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    B = int(num_graphs)

    if fixed_atoms is not None:
        n = torch.full((B,), int(fixed_atoms), dtype=torch.long, device=dev)
    else:
        z = torch.randn(B, generator=gen, device=dev)
        n = torch.exp(math.log(30.0) + 0.35 * z).round().long().clamp_(MIN_ATOMS, MAX_ATOMS)
    nmax = int(n.max())

    deg = torch.zeros(B, nmax, dtype=torch.long, device=dev)
    adj = torch.zeros(B, nmax, nmax, dtype=torch.bool, device=dev)
    hub = torch.rand(B, nmax, generator=gen, device=dev) < 0.01
    cap = torch.where(hub, 6, 4)
    rows = torch.arange(B, device=dev)

    def bond(mask, a, b):
        r = rows[mask]
        adj[r, a[mask], b[mask]] = True
        adj[r, b[mask], a[mask]] = True
        deg.index_put_((r, a[mask]), torch.ones_like(r), accumulate=True)
        deg.index_put_((r, b[mask]), torch.ones_like(r), accumulate=True)

    # --- random tree ------------------------------------------------------------
    for k in range(1, nmax):
        active = n > k
        w = min(6, k)
        r = torch.randint(0, w, (B,), generator=gen, device=dev)
        cand = k - 1 - r
        full = deg[rows, cand] >= cap[rows, cand]
        cand = torch.where(full, torch.full_like(cand, k - 1), cand)
        # hub attraction: first non-saturated hub inside the window wins a coin flip
        win = torch.arange(k - w, k, device=dev)
        hub_ok = hub[:, win] & (deg[:, win] < cap[:, win])
        any_hub = hub_ok.any(dim=1) & (torch.rand(B, generator=gen, device=dev) < 0.7)
        first_hub = win[hub_ok.float().argmax(dim=1)]
        parent = torch.where(any_hub, first_hub, cand)
        bond(active, parent, torch.full_like(parent, k))

    # --- ring closures ----------------------------------------------------------
    rings = torch.clamp(n // 12, min=1)
    for r_i in range(int(rings.max())):
        size = torch.randint(5, 7, (B,), generator=gen, device=dev)
        a = (torch.rand(B, generator=gen, device=dev) * (n - size + 1).clamp(min=0).float()).long()
        b = a + size - 1
        ok = (r_i < rings) & (b < n) & (n >= size)
        b = b.clamp(max=nmax - 1)
        ok &= (deg[rows, a] < cap[rows, a]) & (deg[rows, b] < cap[rows, b]) & ~adj[rows, a, b]
        bond(ok, a, b)

    # --- flatten: valid atoms, row-major nonzero() edge order -----------------------
    valid = torch.arange(nmax, device=dev).unsqueeze(0) < n.unsqueeze(1)       # [B, nmax]
    ptr = torch.zeros(B + 1, dtype=torch.long, device=dev)
    ptr[1:] = torch.cumsum(n, 0)
    gb, gi, gj = adj.nonzero(as_tuple=True)
    edge_index = torch.stack([ptr[gb] + gi, ptr[gb] + gj])
    batch_vec = torch.repeat_interleave(torch.arange(B, device=dev), n)
    N = int(batch_vec.numel())
    degree = deg[valid]

    # --- 35 one-hot features ----------------------------------------------------------
    x = torch.zeros(N, NUM_FEATURES, dtype=torch.float32, device=dev)
    ar = torch.arange(N, device=dev)
    x[ar, _categorical(_SYMBOL_P, N, gen, dev)] = 1.0
    x[ar, 10 + degree.clamp(max=6)] = 1.0
    x[ar, 17 + _categorical(_VALENCE_P, N, gen, dev)] = 1.0
    hyb = _categorical(_HYBRID_P, N, gen, dev)
    keep = hyb < 5
    x[ar[keep], 24 + hyb[keep]] = 1.0
    x[:, 29] = (torch.rand(N, generator=gen, device=dev) < 0.35).float()
    x[ar, 30 + _categorical(_HS_P, N, gen, dev)] = 1.0

    out = Batch(x=x, edge_index=edge_index)
    out._store["batch"] = _tag_num_graphs(batch_vec, B)
    out._store["ptr"] = ptr
    out.__dict__["_num_graphs"] = B
    if with_targets:
        out._store["y"] = torch.randn(B, generator=gen, device=dev)
    return out


def batch_seed(base_seed: int, rank: int, batch_idx: int) -> int:
    """Per-rank, per-batch seed: disjoint molecule streams per rank."""
    h = (base_seed + rank) * 1_000_003 + batch_idx * 7_919 + 12_345
    return h % (2 ** 31 - 1)


def random_graph(num_nodes: int, num_edges: int, seed: int, num_features: int = 35,
                 self_loops: bool = False, device="cpu"):
    """Unstructured random directed multigraph with dense random features (tie-free tests)."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    src = torch.randint(0, num_nodes, (num_edges,), generator=gen)
    dst = torch.randint(0, num_nodes, (num_edges,), generator=gen)
    if not self_loops:
        clash = src == dst
        dst[clash] = (dst[clash] + 1) % max(num_nodes, 1)
    x = torch.randn(num_nodes, num_features, generator=gen)
    return x.to(device), torch.stack([src, dst]).to(device)
