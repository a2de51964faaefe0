"""One CUDA graph per (model, batch size) for the reference's own hyper-parameters.

The reference trains with 64 / 128 molecules per step (ablation/model1.py:109, train.py:209): ~2-4 k atoms, a
few hundred microseconds of GPU work behind ~110 kernel launches -- the step is bound by the host's launch
rate (1.7-2.2 ms), not by the B200.  ``GraphedStep`` captures K0 + forward + loss + backward + optimiser once
and replays it per batch: the host cost of a step becomes one graph launch plus a handful of copies.

Batches differ in atom / bond count, a captured graph has fixed shapes, so every batch is *padded* into static
buffers of a fixed capacity:

* padding atoms: zero feature rows appended after the real atoms and dealt evenly to ``P`` extra molecules (ids
  ``B .. B+P-1``, ~90 atoms each at most: the pooling kernels walk a molecule's atoms with one thread per feature
  chunk, one 5000-atom padding molecule cost 1 ms), so the ``batch`` vector stays sorted and the real molecules'
  pooled rows are rows ``0 .. B-1``;
* padding bonds: self loops spread round-robin over the padding atoms (bounded in-degree).

SAGEConv / GATConv / the pools never mix molecules, so rows of real atoms and real molecules are what the
unpadded step computes (the padding molecule is outside the loss, its gradient rows are zero); weight gradients sum
the same non-zero terms (padding contributes exact zeros) in a different order, i.e. agree to fp32 rounding.
NOT valid for layers that mix the atoms of a batch: ``ModifiedGATLayer`` with whole-batch attention (use
``attention.molecule_attention`` semantics there) and ``BatchNorm1d`` over atoms in training mode (gnn/gin.py), whose
statistics would include the padding atoms.  A batch that does not fit (or is not full) runs eagerly through
the same code, so results never depend on which path ran.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .data import Data, _tag_num_graphs
from .graph import drop_cached_index


class GraphedStep:
    """``step(batch) -> loss`` (training, ``optimizer`` given) or ``step(batch) -> out[B, ...]`` (inference).

    ``model(data)`` takes a ``Data`` with ``x / edge_index / batch``; ``loss_fn(out[:B], y)`` is captured with it.
    The optimiser must support capture (``torch.optim.Adam(..., capturable=True)``, fused or not).
    The returned tensor is a static buffer that the next call overwrites.
    """

    def __init__(self, model: torch.nn.Module, num_graphs: int, max_nodes: int, max_edges: int,
                 optimizer: Optional[torch.optim.Optimizer] = None, loss_fn: Optional[Callable] = None,
                 num_features: int = 35, warmup: int = 3, device=None, pool: Optional[tuple] = None,
                 forward: Optional[Callable] = None):
        self.model, self.opt, self.loss_fn = model, optimizer, loss_fn
        # forward(model, data) -> out, default model(data); e.g. the train.py trunk with per-molecule attention:
        #   forward=lambda m, d: molecule_scope(m, d)  with  `with attention.molecule_attention(d.batch): return m(d)`
        self.forward = forward if forward is not None else (lambda m, d: m(d))
        self.B, self.n_cap, self.e_cap = int(num_graphs), int(max_nodes), int(max_edges)
        dev = torch.device(device) if device is not None else next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedStep needs a CUDA device: the hot path has no CPU fallback")
        if optimizer is not None and loss_fn is None:
            raise ValueError("training needs a loss_fn(out, y)")
        self.device, self.warmup, self.pool = dev, int(warmup), pool
        self.P = max(1, -(-int(0.15 * self.n_cap) // 90))             # padding molecules
        self.x = torch.zeros(self.n_cap, num_features, device=dev)
        self.edge_index = torch.zeros(2, self.e_cap, dtype=torch.long, device=dev)
        self.batch = _tag_num_graphs(torch.full((self.n_cap,), self.B, dtype=torch.long, device=dev), self.B + self.P)
        self.y = torch.zeros(self.B, device=dev)
        self._iota = torch.arange(max(self.e_cap, self.n_cap), device=dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.result: Optional[torch.Tensor] = None
        self.replays = self.eager = 0

    # -- padding ---------------------------------------------------------------------------------
    def fits(self, batch) -> bool:
        n, e = batch.x.size(0), batch.edge_index.size(1)
        if getattr(batch, "num_graphs", None) != self.B or n >= self.n_cap or e > self.e_cap:
            return False
        return (self.e_cap - e) <= 4 * (self.n_cap - n)          # padding self loops per padding atom

    def _fill(self, batch) -> None:
        n, e = batch.x.size(0), batch.edge_index.size(1)
        self.x[:n].copy_(batch.x, non_blocking=True)
        self.x[n:].zero_()
        self.edge_index[:, :e].copy_(batch.edge_index, non_blocking=True)
        if e < self.e_cap:
            pad = (self._iota[:self.e_cap - e] % (self.n_cap - n)) + n
            self.edge_index[:, e:] = pad
        self.batch[:n].copy_(batch.batch, non_blocking=True)
        nd = self.n_cap - n
        self.batch[n:] = torch.div(self._iota[:nd] * self.P, nd, rounding_mode="floor") + self.B
        if self.opt is not None:
            self.y.copy_(batch.y.view(-1), non_blocking=True)

    # -- the step itself (eager and captured run the same code) ------------------------------------
    def _body(self, data, y, num_graphs):
        out = self.forward(self.model, data)[:num_graphs]
        if self.opt is None:
            return out
        loss = self.loss_fn(out, y)
        loss.backward()
        self.opt.step()
        return loss

    def _static_data(self) -> Data:
        return Data(x=self.x, edge_index=self.edge_index, batch=self.batch)

    def _capture(self) -> None:
        params = [p for p in self.model.parameters()]
        saved_p = [p.detach().clone() for p in params]
        had_state = self.opt is not None and len(self.opt.state) > 0
        saved_s = None
        if had_state:
            saved_s = {p: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                       for p, st in self.opt.state.items()}
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):                          # allocator / cuBLAS / lazy-init warm-up
                if self.opt is not None:
                    self.opt.zero_grad(set_to_none=True)
                with torch.set_grad_enabled(self.opt is not None):
                    self._body(self._static_data(), self.y, self.B)
        torch.cuda.current_stream(self.device).wait_stream(side)
        if self.opt is not None:
            self.opt.zero_grad(set_to_none=True)
        # the warm-up left the CSR / segment pointers of THIS batch cached on the static tensors: drop them, the graph
        # has to contain the builder kernels (K0) so that every replay indexes its own batch
        drop_cached_index(self.edge_index, self.batch)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=self.pool):
            with torch.set_grad_enabled(self.opt is not None):
                self.result = self._body(self._static_data(), self.y, self.B)
        drop_cached_index(self.edge_index, self.batch)       # (they point into the graph's memory pool)
        # the warm-up steps must not count as training: put parameters and optimiser state back IN PLACE (the
        # captured graph holds their addresses)
        with torch.no_grad():
            for p, s in zip(params, saved_p):
                p.copy_(s)
            if self.opt is not None:
                for p, st in self.opt.state.items():
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            if saved_s is not None and p in saved_s and k in saved_s[p]:
                                v.copy_(saved_s[p][k])
                            else:
                                v.zero_()

    def __call__(self, batch):
        if not self.fits(batch):
            self.eager += 1
            if self.opt is not None:
                self.opt.zero_grad(set_to_none=True)
            with torch.set_grad_enabled(self.opt is not None):
                return self._body(batch, batch.y.view(-1) if self.opt is not None else None, batch.num_graphs)
        self._fill(batch)
        if self.graph is None:
            self._capture()
        self.graph.replay()
        self.replays += 1
        return self.result
