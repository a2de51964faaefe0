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


class DataLoader(_TorchDataLoader):
    """``torch_geometric.data.DataLoader`` / ``torch_geometric.loader.DataLoader``."""

    def __init__(self, dataset: Iterable, batch_size: int = 1, shuffle: bool = False, **kwargs):
        kwargs.pop("collate_fn", None)
        if shuffle and "sampler" not in kwargs and "batch_sampler" not in kwargs:
            kwargs["sampler"] = _HostRandomSampler(dataset, kwargs.pop("generator", None))
            shuffle = False
        super().__init__(dataset, batch_size, shuffle, collate_fn=Collater(), **kwargs)

    def __iter__(self):
        # the iterator draws its base seed with torch.empty(()).random_(): keep that on the host even when the
        # default device is CUDA
        with torch.device("cpu"):
            return super().__iter__()
