"""Host-side mirror of ``torch_geometric.data`` for the hot path.

Only what the reference scripts touch is here:

* ``Data(x=, edge_index=, batch=)`` with free attribute assignment
  (``data.y`` / ``data.y_original`` -- /root/reference/train.py:189-191;
  ``data.batch = ...`` -- test.py:186), ``.to(device)`` / ``.cpu()``
  (test.py:188, gnnexplainer.py:642).
* ``DataLoader(list, batch_size, shuffle)`` over a list of ``Data`` or of
  ``(Data, Tensor[1,1024])`` tuples (train.py:192,209-210) yielding
  ``[Batch, Tensor[B,1,1024]]``; ``len(loader)`` (train.py:278).
* ``Batch.from_data_list`` collation rules of PyG: tensors whose key contains
  ``"index"`` are concatenated along the last dim and offset by the running
  node count, everything else along dim 0; 0-dim tensors are stacked to
  ``[B]`` (train.py:243 relies on this for ``y``); ``batch`` and ``ptr`` are
  generated.

Unlike PyG, collation also records ``num_graphs`` on the ``batch`` index
tensor itself so that ``global_*_pool`` never needs ``batch.max().item()``
(a host sync in the reference path).
"""
from __future__ import annotations

from typing import Any, Iterable, List, Optional, Sequence

import numpy as np
import torch
from torch.utils.data import DataLoader as _TorchDataLoader
from torch.utils.data.dataloader import default_collate

_HINT = "_mgs_num_graphs"


def _tag_num_graphs(batch_vec: torch.Tensor, num_graphs: int) -> torch.Tensor:
    """Remember the graph count on the index tensor (no device sync later)."""
    setattr(batch_vec, _HINT, int(num_graphs))
    return batch_vec


class Data:
    """A plain attribute bag of tensors describing one graph (or a collated batch)."""

    def __init__(self, x: Optional[torch.Tensor] = None, edge_index: Optional[torch.Tensor] = None,
                 edge_attr: Optional[torch.Tensor] = None, y: Any = None, pos=None, **kwargs):
        self.__dict__["_store"] = {}
        for k, v in dict(x=x, edge_index=edge_index, edge_attr=edge_attr, y=y, pos=pos, **kwargs).items():
            if v is not None:
                self._store[k] = v

    # --- attribute protocol -------------------------------------------------
    def __getattr__(self, key: str):
        store = self.__dict__.get("_store", {})
        if key in store:
            return store[key]
        if key in ("x", "edge_index", "edge_attr", "y", "pos", "batch", "ptr"):
            return None
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{key}'")

    def __setattr__(self, key: str, value: Any) -> None:
        if key.startswith("_") and key != "_store":
            self.__dict__[key] = value
        elif value is None:
            self._store.pop(key, None)
        else:
            self._store[key] = value

    def __delattr__(self, key: str) -> None:
        self._store.pop(key, None)

    def __getitem__(self, key: str):
        return self._store[key]

    def __setitem__(self, key: str, value: Any) -> None:
        setattr(self, key, value)

    def __contains__(self, key: str) -> bool:
        return key in self._store

    def __iter__(self):
        return iter(self._store.items())

    def keys(self) -> List[str]:
        return list(self._store.keys())

    def to_dict(self) -> dict:
        return dict(self._store)

    def __getstate__(self):
        return {"_store": self._store, **{k: v for k, v in self.__dict__.items() if k != "_store"}}

    def __setstate__(self, state):
        self.__dict__.update(state)

    # --- sizes ---------------------------------------------------------------
    @property
    def num_nodes(self) -> Optional[int]:
        if "x" in self._store:
            return int(self._store["x"].size(0))
        if "batch" in self._store:
            return int(self._store["batch"].numel())
        if "edge_index" in self._store and self._store["edge_index"].numel() > 0:
            return int(self._store["edge_index"].max()) + 1
        return None

    @property
    def num_edges(self) -> int:
        ei = self._store.get("edge_index")
        return 0 if ei is None else int(ei.size(1))

    @property
    def num_node_features(self) -> int:
        x = self._store.get("x")
        return 0 if x is None else (1 if x.dim() == 1 else int(x.size(-1)))

    num_features = num_node_features

    # --- movement ------------------------------------------------------------
    def _apply(self, fn):
        for k, v in list(self._store.items()):
            if isinstance(v, torch.Tensor):
                nv = fn(v)
                if hasattr(v, _HINT) and nv is not v:
                    setattr(nv, _HINT, getattr(v, _HINT))
                self._store[k] = nv
        return self

    def to(self, device, *args, **kwargs):
        return self._apply(lambda t: t.to(device, *args, **kwargs))

    def cpu(self):
        return self._apply(lambda t: t.cpu())

    def cuda(self, device=None, non_blocking: bool = False):
        return self._apply(lambda t: t.cuda(device, non_blocking=non_blocking))

    def pin_memory(self):
        return self._apply(lambda t: t.pin_memory())

    def clone(self):
        out = type(self).__new__(type(self))
        out.__dict__["_store"] = {k: (v.clone() if isinstance(v, torch.Tensor) else v)
                                  for k, v in self._store.items()}
        for k, v in self.__dict__.items():
            if k != "_store":
                out.__dict__[k] = v
        return out

    def __repr__(self) -> str:
        parts = []
        for k, v in self._store.items():
            parts.append(f"{k}={list(v.shape)}" if isinstance(v, torch.Tensor) else f"{k}={v!r}")
        return f"{type(self).__name__}({', '.join(parts)})"


def _cat_dim(key: str, value: torch.Tensor) -> int:
    return -1 if ("index" in key or key == "face") else 0


def _is_index_key(key: str) -> bool:
    return "index" in key or key == "face"


class Batch(Data):
    """Several graphs collated into one disconnected graph (PyG ``Batch``)."""

    @classmethod
    def from_data_list(cls, data_list: Sequence[Data]) -> "Batch":
        if len(data_list) == 0:
            raise ValueError("cannot collate an empty list of graphs")
        keys = data_list[0].keys()
        out = cls()
        counts = []
        for d in data_list:
            n = d.num_nodes
            if n is None:
                raise ValueError("every graph needs `x` (or a `batch`) to define its node count")
            counts.append(n)
        ptr = [0]
        for n in counts:
            ptr.append(ptr[-1] + n)
        for key in keys:
            if key in ("batch", "ptr"):
                continue
            vals = [d[key] for d in data_list]
            v0 = vals[0]
            if isinstance(v0, torch.Tensor):
                if v0.dim() == 0:
                    out._store[key] = torch.stack(vals)
                elif _is_index_key(key):
                    out._store[key] = torch.cat([v + off for v, off in zip(vals, ptr[:-1])],
                                                dim=_cat_dim(key, v0))
                else:
                    out._store[key] = torch.cat(vals, dim=0)
            elif isinstance(v0, (int, float)):
                out._store[key] = torch.tensor(vals)
            else:
                out._store[key] = vals
        dev = data_list[0].x.device if data_list[0].x is not None else None
        counts_t = torch.tensor(counts, dtype=torch.long, device=dev)
        batch_vec = torch.repeat_interleave(torch.arange(len(counts), dtype=torch.long, device=dev),
                                            counts_t, output_size=ptr[-1])
        out._store["batch"] = _tag_num_graphs(batch_vec, len(counts))
        out._store["ptr"] = torch.tensor(ptr, dtype=torch.long, device=dev)
        out.__dict__["_num_graphs"] = len(counts)
        return out

    @property
    def num_graphs(self) -> int:
        n = self.__dict__.get("_num_graphs")
        if n is not None:
            return n
        if "ptr" in self._store:
            return int(self._store["ptr"].numel()) - 1
        if "batch" in self._store:
            hinted = getattr(self._store["batch"], _HINT, None)
            return hinted if hinted is not None else int(self._store["batch"].max()) + 1
        raise ValueError("batch has neither `ptr` nor `batch`")

    def _apply(self, fn):
        super()._apply(fn)
        if "batch" in self._store and self.__dict__.get("_num_graphs") is not None:
            _tag_num_graphs(self._store["batch"], self.__dict__["_num_graphs"])
        return self

    def to_data_list(self) -> List[Data]:
        ptr = self._store["ptr"].tolist()
        out = []
        for g in range(len(ptr) - 1):
            lo, hi = ptr[g], ptr[g + 1]
            d = Data()
            for key, v in self._store.items():
                if key in ("batch", "ptr"):
                    continue
                if isinstance(v, torch.Tensor):
                    if _is_index_key(key):
                        m = (v[0] >= lo) & (v[0] < hi)
                        d._store[key] = v[:, m] - lo
                    elif v.size(0) == ptr[-1]:
                        d._store[key] = v[lo:hi]
                    elif v.size(0) == len(ptr) - 1:
                        d._store[key] = v[g]
                else:
                    d._store[key] = v[g]
            out.append(d)
        return out


class WireBatch:
    """Compact host-side image of a ``Batch`` for the host -> device copy (pinned memory):

    * ``xbits``  uint64 ``[N]``   -- the atom features as one bit each (``F <= 64``; the reference's 35 features are
      one-hot groups, exactly 0.0 / 1.0: train.py:33-43) -- 8 bytes per atom instead of ``4 F`` = 140;
    * ``edge_index`` int32 ``[2, E]`` (atom ids fit 31 bits), ``ptr`` int32 ``[B + 1]`` instead of the ``[N]`` int64 batch vector;
    * ``y`` (optional) as it is.

    4096 molecules (130 k atoms, 266 k edges): 3.2 MB instead of 23.6 MB.  ``to_batch(device)`` issues the (non-blocking)
    copies and ONE expansion launch (``mgs_wire_expand``) and returns an ordinary device ``Batch`` whose tensors are
    bit-identical to ``batch.to(device)``.  Features that are not all 0 / 1 cannot be packed: ``from_batch`` raises."""

    def __init__(self, xbits, num_features, edge_index, ptr, y=None):
        self.xbits, self.num_features, self.edge_index, self.ptr, self.y = xbits, int(num_features), edge_index, ptr, y

    @classmethod
    def from_batch(cls, batch: "Batch", pin: bool = True) -> "WireBatch":
        x = batch.x.detach().cpu()
        if x.dim() != 2 or x.size(1) > 64:
            raise ValueError("WireBatch packs [N, F <= 64] feature matrices")
        if not bool(((x == 0) | (x == 1)).all()):
            raise ValueError("WireBatch: features must be exactly 0.0 / 1.0 (the reference's one-hot featurisation)")
        weights = (torch.ones(x.size(1), dtype=torch.int64) << torch.arange(x.size(1), dtype=torch.int64))
        xbits = (x.to(torch.int64) * weights).sum(dim=1)            # bit f of word n = x[n, f]  (F <= 63 stays positive;
        ei = batch.edge_index.detach().cpu()                        #  F = 64 wraps into the sign bit, same bits)
        if x.size(0) >= 2 ** 31:
            raise ValueError("WireBatch: atom ids must fit int32")
        if "ptr" in batch._store:
            ptr = batch._store["ptr"].detach().cpu()
        else:
            counts = torch.bincount(batch.batch.detach().cpu(), minlength=batch.num_graphs)
            ptr = torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)])
        y = batch._store.get("y")
        out = cls(xbits, x.size(1), ei.to(torch.int32).contiguous(), ptr.to(torch.int32).contiguous(),
                  None if y is None else y.detach().cpu().contiguous())
        return out.pin_memory() if pin and torch.cuda.is_available() else out

    def pin_memory(self) -> "WireBatch":
        for k in ("xbits", "edge_index", "ptr", "y"):
            v = getattr(self, k)
            if v is not None and not v.is_pinned():
                setattr(self, k, v.pin_memory())
        return self

    @property
    def nbytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in (self.xbits, self.edge_index, self.ptr, self.y) if v is not None)

    def to_batch(self, device, buffers: Optional[dict] = None) -> "Batch":
        """Copies + one expansion kernel on the current stream of ``device``.  ``buffers``: optional dict of preallocated
        device tensors (``xbits``, ``edge_index32``, ``ptr32``, ``ptr``, ``y``, ``x``, ``edge_index``, ``batch``) at least as
        large as needed -- a training loop reuses them instead of allocating per step."""
        from . import _lib
        from .graph import device_guard, stream_ptr
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("WireBatch.to_batch: the expansion runs on a CUDA device (no CPU fallback)")
        N, E, B, F = int(self.xbits.numel()), int(self.edge_index.size(1)), int(self.ptr.numel()) - 1, self.num_features

        def buf(name, shape, dtype):
            n = 1
            for s_ in shape:
                n *= s_
            if buffers is not None and name in buffers and buffers[name].numel() >= n and buffers[name].dtype == dtype:
                return buffers[name].view(-1)[:n].view(shape)
            return torch.empty(shape, dtype=dtype, device=dev)

        xbits = buf("xbits", (N,), torch.int64)
        ei32 = buf("edge_index32", (2, E), torch.int32)
        ptr32 = buf("ptr32", (B + 1,), torch.int32)
        xbits.copy_(self.xbits, non_blocking=True)
        ei32.copy_(self.edge_index, non_blocking=True)
        ptr32.copy_(self.ptr, non_blocking=True)
        x = buf("x", (N, F), torch.float32)
        ei = buf("edge_index", (2, E), torch.int64)
        bvec = buf("batch", (N,), torch.int64)
        lib = _lib.load()
        with device_guard(dev):
            rc = lib.mgs_wire_expand(xbits.data_ptr(), N, F, x.data_ptr(), F, ei32.data_ptr(), E, ei.data_ptr(),
                                     ptr32.data_ptr(), B, bvec.data_ptr(), stream_ptr())
        _lib.check(rc, "mgs_wire_expand")
        out = Batch(x=x, edge_index=ei)
        out._store["batch"] = _tag_num_graphs(bvec, B)
        ptr64 = buf("ptr", (B + 1,), torch.int64)
        ptr64.copy_(ptr32)
        out._store["ptr"] = ptr64
        out.__dict__["_num_graphs"] = B
        if self.y is not None:
            yb = buf("y", tuple(self.y.shape), self.y.dtype)
            yb.copy_(self.y, non_blocking=True)
            out._store["y"] = yb
        return out


class Collater:
    """PyG's collate dispatch: ``Data`` -> ``Batch``, tensors -> stacked, tuples -> element-wise."""

    def __call__(self, batch: List[Any]) -> Any:
        elem = batch[0]
        if isinstance(elem, Data):
            return Batch.from_data_list(batch)
        if isinstance(elem, torch.Tensor):
            return default_collate(batch)
        if isinstance(elem, float):
            return torch.tensor(batch, dtype=torch.float)
        if isinstance(elem, int):
            return torch.tensor(batch)
        if isinstance(elem, str):
            return batch
        if isinstance(elem, dict):
            return {k: self([b[k] for b in batch]) for k in elem}
        if isinstance(elem, (tuple, list)):
            return [self(list(s)) for s in zip(*batch)]
        raise TypeError(f"DataLoader found invalid type: {type(elem)}")


class _HostRandomSampler(torch.utils.data.Sampler):
    """RandomSampler whose permutation is always drawn on the host.  With a CUDA default device
    (``m_gat_graphsage_b200.run``) the stock sampler would call ``torch.randperm`` on the GPU with a CPU
    generator and fail; indices are host data anyway."""

    def __init__(self, data_source, generator=None):
        self.data_source = data_source
        self.generator = generator

    def __len__(self) -> int:
        return len(self.data_source)

    def __iter__(self):
        with torch.device("cpu"):
            gen = self.generator
            if gen is None:
                gen = torch.Generator(device="cpu")
                gen.manual_seed(int(torch.empty((), dtype=torch.int64).random_().item()))
            yield from torch.randperm(len(self.data_source), generator=gen).tolist()


class _FlatGraphs:
    """A list of ``Data`` collated ONCE: every tensor attribute concatenated over all molecules plus, per attribute,
    the extent of every molecule along the concatenation axis.  A mini-batch is then a handful of vectorised
    gathers (on the device the flat tensors live on) instead of a Python loop over molecules:

        host collation (``Batch.from_data_list``)   ~35 k molecules/s  (116 ms per 4096-molecule batch)
        flat gather on the GPU                       > 5 M molecules/s

    -- the message-passing step runs at 1.2 M molecules/s, so a per-epoch Python collation (train.py:209-210,236)
    would throttle it 30x.  The batches are bit-identical to ``Batch.from_data_list`` of the same molecules."""

    def __init__(self, data_list: Sequence[Data], device=None):
        if len(data_list) == 0:
            raise ValueError("empty dataset")
        self.keys = [k for k in data_list[0].keys() if k not in ("batch", "ptr")]
        self.num = len(data_list)
        counts = []
        for d in data_list:
            if list(k for k in d.keys() if k not in ("batch", "ptr")) != self.keys:
                raise TypeError("graphs with different attribute sets")
            n = d.num_nodes
            if n is None:
                raise ValueError("every graph needs `x` (or a `batch`) to define its node count")
            counts.append(n)
        # extent tables ("groups"): attributes with the same per-molecule extents share one gather index;
        # group 0 is always the atoms
        self.tables = [np.concatenate([[0], np.cumsum(np.asarray(counts, dtype=np.int64))])]
        self.flat, self.group, self.kind = {}, {}, {}
        unit = np.arange(self.num + 1, dtype=np.int64)
        for key in self.keys:
            vals = [d[key] for d in data_list]
            v0 = vals[0]
            if isinstance(v0, (int, float)):
                vals = [torch.tensor(v) for v in vals]
                v0 = vals[0]
            if not isinstance(v0, torch.Tensor):
                raise TypeError(f"attribute {key!r} is not a tensor")
            if v0.dim() == 0:
                kind, flat, table = "rows", torch.stack(vals), unit
            else:
                kind = "index" if _is_index_key(key) else "cat"     # index: LOCAL node ids, joined on the last dim
                dim = -1 if kind == "index" else 0
                flat = torch.cat(vals, dim=dim)
                table = np.concatenate([[0], np.cumsum(np.asarray([v.size(dim) for v in vals], dtype=np.int64))])
                if kind == "cat" and np.array_equal(table, unit):
                    kind = "rows"                                    # one row per molecule: gather by molecule id
            if kind != "rows":
                for g, t in enumerate(self.tables):
                    if np.array_equal(t, table):
                        self.group[key] = g
                        break
                else:
                    self.group[key] = len(self.tables)
                    self.tables.append(table)
            self.kind[key] = kind
            self.flat[key] = flat if device is None else flat.to(device)
        self.device = next(iter(self.flat.values())).device
        self._iota = torch.arange(1 << 16, device=self.device)

    def _arange(self, n: int) -> torch.Tensor:
        if n > self._iota.numel():
            self._iota = torch.arange(max(n, 2 * self._iota.numel()), device=self.device)
        return self._iota[:n]

    def batch(self, ids: Sequence[int]) -> "Batch":
        ids_np = np.asarray(ids, dtype=np.int64)
        b, G, dev = ids_np.size, len(self.tables), self.device
        # every small integer table of this batch goes to the device in ONE (pinned, asynchronous) copy:
        # [ids | per group: sizes, old start - new start, new start | new node ptr (b+1)]
        n_pack = b * (1 + 3 * G) + b + 1
        pack_t = torch.empty(n_pack, dtype=torch.long, device="cpu", pin_memory=dev.type == "cuda")
        pack = pack_t.numpy()
        pack[:b] = ids_np
        totals = []
        for g, t in enumerate(self.tables):
            starts = t[ids_np]
            sizes = t[ids_np + 1] - starts
            o = b * (1 + 3 * g)
            new_start = pack[o + 2 * b:o + 3 * b]
            new_start[0] = 0
            np.cumsum(sizes[:-1], out=new_start[1:])
            pack[o:o + b] = sizes
            pack[o + b:o + 2 * b] = starts - new_start
            tot = int(sizes.sum())
            totals.append(tot)
            if g == 0:
                pack[n_pack - b - 1:n_pack - 1] = new_start
                pack[n_pack - 1] = tot
        pd = pack_t.to(dev, non_blocking=True) if dev.type == "cuda" else pack_t
        ids_d = pd[:b]
        seg, idx = {}, {}

        def segments(g):
            if g not in seg:
                o = b * (1 + 3 * g)
                seg[g] = torch.repeat_interleave(self._arange(b), pd[o:o + b], output_size=totals[g])
                idx[g] = self._arange(totals[g]) + pd[o + b:o + 2 * b][seg[g]]
            return seg[g], idx[g]

        out = Batch()
        for key in self.keys:
            flat, kind = self.flat[key], self.kind[key]
            if kind == "rows":
                out._store[key] = flat.index_select(0, ids_d)
                continue
            sg, ix = segments(self.group[key])
            if kind == "index":
                out._store[key] = flat.index_select(-1, ix) + pd[3 * b:4 * b][sg]      # + new first atom of its molecule
            else:
                out._store[key] = flat.index_select(0, ix)
        out._store["batch"] = _tag_num_graphs(segments(0)[0], b)
        out._store["ptr"] = pd[n_pack - b - 1:]
        out.__dict__["_num_graphs"] = int(b)
        out.__dict__["_ids"] = ids_d                          # molecule ids of this batch, on the device
        return out


class DataLoader(_TorchDataLoader):
    """``torch_geometric.data.DataLoader`` / ``torch_geometric.loader.DataLoader``.

    ``fast=True`` (default): a list of ``Data`` (or of ``(Data, Tensor, ...)`` tuples, train.py:192) is collated once
    into flat tensors on first use (``_FlatGraphs``) and every batch is gathered from them -- on ``device`` when
    given, else on the default device if that is CUDA (``m_gat_graphsage_b200.run``), else where the tensors
    are.  Same batches, bit for bit, as the per-batch Python collation (``fast=False``).  The flat copy is taken at
    the first ``iter()``: edits to the ``Data`` objects after that are not seen (the reference builds its lists once,
    train.py:169-193); heterogeneous attribute sets, non-tensor attributes or ``num_workers > 0`` use the Python path."""

    def __init__(self, dataset: Iterable, batch_size: int = 1, shuffle: bool = False, fast: bool = True, device=None,
                 **kwargs):
        kwargs.pop("collate_fn", None)
        if shuffle and "sampler" not in kwargs and "batch_sampler" not in kwargs:
            kwargs["sampler"] = _HostRandomSampler(dataset, kwargs.pop("generator", None))
            shuffle = False
        super().__init__(dataset, batch_size, shuffle, collate_fn=Collater(), **kwargs)
        self._fast = bool(fast) and kwargs.get("num_workers", 0) == 0 and isinstance(dataset, (list, tuple))
        self._fast_device = device
        self._flat = None

    def _build_flat(self):
        ds = self.dataset
        dev = self._fast_device
        if dev is None:
            d0 = torch.get_default_device() if hasattr(torch, "get_default_device") else torch.device("cpu")
            dev = d0 if d0.type == "cuda" else None
        with torch.device("cpu"):
            elem = ds[0]
            if isinstance(elem, Data):
                return (_FlatGraphs(ds, dev), None)
            if isinstance(elem, (tuple, list)) and len(elem) > 0 and isinstance(elem[0], Data) and \
                    all(isinstance(t, torch.Tensor) for t in elem[1:]):
                graphs = _FlatGraphs([e[0] for e in ds], dev)
                extras = [torch.stack([e[i] for e in ds]) for i in range(1, len(elem))]
                return (graphs, [t.to(graphs.device) for t in extras])
        raise TypeError("dataset elements are neither Data nor (Data, Tensor, ...) tuples")

    def _fast_iter(self):
        graphs, extras = self._flat
        for ids in self.batch_sampler:
            b = graphs.batch(ids)
            if extras is None:
                yield b
            else:
                yield [b] + [t.index_select(0, b._ids) for t in extras]

    def __iter__(self):
        if self._fast and self._flat is None:
            try:
                self._flat = self._build_flat()
            except (TypeError, ValueError, RuntimeError):
                self._fast = False                            # heterogeneous / non-tensor attributes: Python collation
        if self._fast:
            return self._fast_iter()
        # the iterator draws its base seed with torch.empty(()).random_(): keep that on the host even when the
        # default device is CUDA
        with torch.device("cpu"):
            return super().__iter__()
