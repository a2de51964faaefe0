"""Mirror of the ``torch_geometric.explain`` surface the reference uses
(/root/reference/gnnexplainer.py:7-8, 620-631, 669-680): ``Explainer``, ``GNNExplainer``,
``ExplainerConfig``, ``ModelConfig`` and the ``Explanation`` result with ``.node_mask``,
``.edge_mask``, ``.prediction``.  Algorithm per SURVEY.md Appendix A.4: a learnable node-attribute mask
multiplied into ``x`` and a learnable per-edge mask whose sigmoid is multiplied into the MESSAGES of
every ``MessagePassing`` layer (our K1 / K2 kernels take it as ``edge_weight`` and return its
gradient), trained with Adam against the model's own prediction.

The forward/backward inside the loop is the same CUDA hot path as training; the loop itself is host
code, like the reference's.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Dict, Optional

import torch
import torch.nn.functional as F

from .nn import MessagePassing


@dataclass
class ModelConfig:
    mode: str = "regression"
    task_level: str = "graph"
    return_type: Optional[str] = "raw"

    def __post_init__(self):
        self.mode = getattr(self.mode, "value", self.mode)
        self.task_level = getattr(self.task_level, "value", self.task_level)
        self.return_type = getattr(self.return_type, "value", self.return_type)
        if self.mode not in ("regression", "binary_classification", "multiclass_classification"):
            raise ValueError(f"unknown mode {self.mode!r}")
        if self.task_level not in ("graph", "node", "edge"):
            raise ValueError(f"unknown task_level {self.task_level!r}")
        if self.mode == "regression" and self.return_type not in (None, "raw"):
            raise ValueError("regression models must have return_type='raw'")


@dataclass
class ExplainerConfig:
    explanation_type: str = "model"
    node_mask_type: Optional[str] = None
    edge_mask_type: Optional[str] = None

    def __post_init__(self):
        if self.explanation_type not in ("model", "phenomenon"):
            raise ValueError(f"unknown explanation_type {self.explanation_type!r}")
        if self.node_mask_type not in (None, "object", "common_attributes", "attributes"):
            raise ValueError(f"unknown node_mask_type {self.node_mask_type!r}")
        if self.edge_mask_type not in (None, "object"):
            raise ValueError(f"unsupported edge_mask_type {self.edge_mask_type!r}")
        if self.node_mask_type is None and self.edge_mask_type is None:
            raise ValueError("either node_mask_type or edge_mask_type must be set")


class Explanation:
    """Attribute bag returned by ``Explainer.__call__``."""

    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)

    def __getattr__(self, key):  # missing attributes read as None, like PyG's storage
        if key.startswith("__"):
            raise AttributeError(key)
        return None

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]


def set_masks(model: torch.nn.Module, mask: torch.Tensor, apply_sigmoid: bool = True) -> int:
    n = 0
    for module in model.modules():
        if isinstance(module, MessagePassing):
            module._explain = True
            module._edge_mask = mask
            module._apply_sigmoid = apply_sigmoid
            n += 1
    return n


def clear_masks(model: torch.nn.Module) -> None:
    for module in model.modules():
        if isinstance(module, MessagePassing):
            module._explain = False
            module._edge_mask = None
            module._apply_sigmoid = True


class GNNExplainer:
    coeffs: Dict[str, Any] = {
        "edge_size": 0.005, "edge_reduction": "sum",
        "node_feat_size": 1.0, "node_feat_reduction": "mean",
        "edge_ent": 1.0, "node_feat_ent": 0.1, "EPS": 1e-15,
    }

    def __init__(self, epochs: int = 100, lr: float = 0.01, init_node_mask=None, init_edge_mask=None, **kwargs):
        self.epochs, self.lr = epochs, lr
        self.coeffs = dict(type(self).coeffs)
        self.coeffs.update(kwargs)
        #: optional fixed starting masks (PyG draws them with ``torch.randn``); the parity tests start the CUDA path
        #: and the CPU oracle (oracle/explainer_oracle.py) from the same point
        self._init_node_mask, self._init_edge_mask = init_node_mask, init_edge_mask
        self.node_mask = self.edge_mask = None
        self.hard_node_mask = self.hard_edge_mask = None
        self.explainer_config: Optional[ExplainerConfig] = None
        self.model_config: Optional[ModelConfig] = None

    def connect(self, explainer_config: ExplainerConfig, model_config: ModelConfig) -> None:
        self.explainer_config, self.model_config = explainer_config, model_config

    # -- A.4 ------------------------------------------------------------------------------------
    def _initialize_masks(self, x: torch.Tensor, edge_index: torch.Tensor) -> None:
        cfg = self.explainer_config
        (N, Fdim), E, dev = x.size(), edge_index.size(1), x.device
        if cfg.node_mask_type is None:
            self.node_mask = None
        elif self._init_node_mask is not None:
            self.node_mask = torch.nn.Parameter(self._init_node_mask.detach().clone().to(dev))
        elif cfg.node_mask_type == "object":
            self.node_mask = torch.nn.Parameter(torch.randn(N, 1, device=dev) * 0.1)
        elif cfg.node_mask_type == "attributes":
            self.node_mask = torch.nn.Parameter(torch.randn(N, Fdim, device=dev) * 0.1)
        else:
            self.node_mask = torch.nn.Parameter(torch.randn(1, Fdim, device=dev) * 0.1)
        if cfg.edge_mask_type is None:
            self.edge_mask = None
        elif self._init_edge_mask is not None:
            self.edge_mask = torch.nn.Parameter(self._init_edge_mask.detach().clone().to(dev))
        else:
            std = torch.nn.init.calculate_gain("relu") * math.sqrt(2.0 / (2 * N))
            self.edge_mask = torch.nn.Parameter(torch.randn(E, device=dev) * std)

    def _loss(self, y_hat: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        mode = self.model_config.mode
        if mode == "regression":
            loss = F.mse_loss(y_hat, y)
        elif mode == "binary_classification":
            loss = F.binary_cross_entropy_with_logits(y_hat.view_as(y), y.float())
        else:
            loss = F.cross_entropy(y_hat, y)
        eps = self.coeffs["EPS"]
        if self.hard_edge_mask is not None:
            m = self.edge_mask[self.hard_edge_mask].sigmoid()
            loss = loss + self.coeffs["edge_size"] * getattr(torch, self.coeffs["edge_reduction"])(m)
            ent = -m * torch.log(m + eps) - (1 - m) * torch.log(1 - m + eps)
            loss = loss + self.coeffs["edge_ent"] * ent.mean()
        if self.hard_node_mask is not None:
            m = self.node_mask[self.hard_node_mask].sigmoid()
            loss = loss + self.coeffs["node_feat_size"] * getattr(torch, self.coeffs["node_feat_reduction"])(m)
            ent = -m * torch.log(m + eps) - (1 - m) * torch.log(1 - m + eps)
            loss = loss + self.coeffs["node_feat_ent"] * ent.mean()
        return loss

    def _train(self, model, x, edge_index, target, index=None, **kwargs) -> None:
        self._initialize_masks(x, edge_index)
        params = [p for p in (self.node_mask, self.edge_mask) if p is not None]
        if self.edge_mask is not None:
            set_masks(model, self.edge_mask, apply_sigmoid=True)
        opt = torch.optim.Adam(params, lr=self.lr)
        for i in range(self.epochs):
            opt.zero_grad()
            h = x if self.node_mask is None else x * self.node_mask.sigmoid()
            y_hat, y = model(h, edge_index, **kwargs), target
            if index is not None:
                y_hat, y = y_hat[index], y[index]
            loss = self._loss(y_hat, y)
            loss.backward()
            opt.step()
            if i == 0 and self.node_mask is not None:
                if self.node_mask.grad is None:
                    raise ValueError("Could not compute gradients for node features")
                self.hard_node_mask = self.node_mask.grad != 0.0
            if i == 0 and self.edge_mask is not None:
                if self.edge_mask.grad is None:
                    raise ValueError("Could not compute gradients for edges: the model's layers do not "
                                     "consume the edge mask")
                self.hard_edge_mask = self.edge_mask.grad != 0.0

    @staticmethod
    def _post_process(mask, hard_mask):
        if mask is None:
            return None
        mask = mask.detach().sigmoid()
        if hard_mask is not None and mask.size() == hard_mask.size():
            mask[~hard_mask] = 0.0
        return mask

    def __call__(self, model, x, edge_index, *, target, index=None, **kwargs) -> Explanation:
        self.hard_node_mask = self.hard_edge_mask = None
        try:
            self._train(model, x, edge_index, target=target, index=index, **kwargs)
            node_mask = self._post_process(self.node_mask, self.hard_node_mask)
            edge_mask = self._post_process(self.edge_mask, self.hard_edge_mask)
        finally:
            clear_masks(model)
            self.node_mask = self.edge_mask = None
            self.hard_node_mask = self.hard_edge_mask = None
        return Explanation(node_mask=node_mask, edge_mask=edge_mask)


class BatchedGNNExplainer(GNNExplainer):
    """GNNExplainer over a whole ``Batch`` of molecules at once.

    The reference explains ~1.2 k molecules one by one, 100 epochs each (gnnexplainer.py:661-690, 1089-1097): pure
    launch latency.  Here the masks of all molecules are one ``[N, F]`` / ``[E]`` parameter pair, the objective is the
    SUM over molecules of the per-molecule GNNExplainer objective (squared error + size / entropy regularisers with
    per-molecule sums and means), and Adam -- element-wise -- updates every molecule's masks exactly as a separate
    run would.  Valid for models whose layers never mix molecules (GATConv / SAGEConv / the pools; not
    ``ModifiedGATLayer``).  Call with ``batch=`` (and ``target`` per molecule); returns the concatenated masks.
    ``init_node_mask`` / ``init_edge_mask`` fix the starting point (tests compare with per-molecule runs)."""

    def __init__(self, epochs: int = 100, lr: float = 0.01, init_node_mask=None, init_edge_mask=None, **kwargs):
        super().__init__(epochs=epochs, lr=lr, init_node_mask=init_node_mask, init_edge_mask=init_edge_mask, **kwargs)

    def _initialize_masks_batched(self, x, edge_index, batch, num_graphs) -> None:
        cfg = self.explainer_config
        (N, Fdim), E, dev = x.size(), edge_index.size(1), x.device
        if cfg.node_mask_type is None:
            self.node_mask = None
        elif self._init_node_mask is not None:
            self.node_mask = torch.nn.Parameter(self._init_node_mask.detach().clone().to(dev))
        elif cfg.node_mask_type == "object":
            self.node_mask = torch.nn.Parameter(torch.randn(N, 1, device=dev) * 0.1)
        elif cfg.node_mask_type == "attributes":
            self.node_mask = torch.nn.Parameter(torch.randn(N, Fdim, device=dev) * 0.1)
        else:
            raise NotImplementedError("node_mask_type='common_attributes' is one mask per molecule: explain them one "
                                      "by one with GNNExplainer")
        if cfg.edge_mask_type is None:
            self.edge_mask = None
        elif self._init_edge_mask is not None:
            self.edge_mask = torch.nn.Parameter(self._init_edge_mask.detach().clone().to(dev))
        else:                                             # per-molecule std, as N differs from molecule to molecule
            n_g = torch.bincount(batch, minlength=num_graphs).clamp_(min=1).to(torch.float32)
            std = torch.nn.init.calculate_gain("relu") * torch.sqrt(2.0 / (2.0 * n_g))
            self.edge_mask = torch.nn.Parameter(torch.randn(E, device=dev) * std[batch[edge_index[0]]])

    @staticmethod
    def _segment_sum(v, seg, num):
        return torch.zeros(num, dtype=v.dtype, device=v.device).index_add_(0, seg, v)

    def _loss_batched(self, y_hat, y, batch, edge_graph, num_graphs) -> torch.Tensor:
        mode = self.model_config.mode
        if mode == "regression":
            per_graph = ((y_hat.view(num_graphs, -1) - y.view(num_graphs, -1)) ** 2).mean(dim=1)
        elif mode == "binary_classification":
            per_graph = F.binary_cross_entropy_with_logits(y_hat.view(num_graphs, -1), y.view(num_graphs, -1).float(),
                                                           reduction="none").mean(dim=1)
        else:
            per_graph = F.cross_entropy(y_hat, y.view(-1), reduction="none")
        loss = per_graph.sum()
        eps = self.coeffs["EPS"]
        if self.hard_edge_mask is not None:
            hard = self.hard_edge_mask.to(torch.float32)
            m = self.edge_mask.sigmoid()
            cnt = self._segment_sum(hard, edge_graph, num_graphs).clamp_(min=1.0)
            size = self._segment_sum(m * hard, edge_graph, num_graphs)
            if self.coeffs["edge_reduction"] == "mean":
                size = size / cnt
            ent = -m * torch.log(m + eps) - (1 - m) * torch.log(1 - m + eps)
            ent = self._segment_sum(ent * hard, edge_graph, num_graphs) / cnt
            loss = loss + (self.coeffs["edge_size"] * size + self.coeffs["edge_ent"] * ent).sum()
        if self.hard_node_mask is not None:
            hard = self.hard_node_mask.to(torch.float32)
            m = self.node_mask.sigmoid()
            cnt = self._segment_sum(hard.sum(dim=1), batch, num_graphs).clamp_(min=1.0)
            size = self._segment_sum((m * hard).sum(dim=1), batch, num_graphs)
            if self.coeffs["node_feat_reduction"] == "mean":
                size = size / cnt
            ent = -m * torch.log(m + eps) - (1 - m) * torch.log(1 - m + eps)
            ent = self._segment_sum((ent * hard).sum(dim=1), batch, num_graphs) / cnt
            loss = loss + (self.coeffs["node_feat_size"] * size + self.coeffs["node_feat_ent"] * ent).sum()
        return loss

    def __call__(self, model, x, edge_index, *, target, batch=None, index=None, **kwargs) -> Explanation:
        if batch is None:
            return super().__call__(model, x, edge_index, target=target, index=index, **kwargs)
        if index is not None:
            raise NotImplementedError("index= selects one output of one graph: use GNNExplainer for that")
        from .graph import resolve_num_graphs
        num_graphs = resolve_num_graphs(batch, None)
        edge_graph = batch[edge_index[0]]
        self.hard_node_mask = self.hard_edge_mask = None
        try:
            self._initialize_masks_batched(x, edge_index, batch, num_graphs)
            params = [p for p in (self.node_mask, self.edge_mask) if p is not None]
            if self.edge_mask is not None:
                set_masks(model, self.edge_mask, apply_sigmoid=True)
            opt = torch.optim.Adam(params, lr=self.lr)
            for i in range(self.epochs):
                opt.zero_grad()
                h = x if self.node_mask is None else x * self.node_mask.sigmoid()
                y_hat = model(h, edge_index, batch=batch, **kwargs)
                loss = self._loss_batched(y_hat, target, batch, edge_graph, num_graphs)
                loss.backward()
                opt.step()
                if i == 0 and self.node_mask is not None:
                    self.hard_node_mask = self.node_mask.grad != 0.0
                if i == 0 and self.edge_mask is not None:
                    if self.edge_mask.grad is None:
                        raise ValueError("Could not compute gradients for edges: the model's layers do not "
                                         "consume the edge mask")
                    self.hard_edge_mask = self.edge_mask.grad != 0.0
            node_mask = self._post_process(self.node_mask, self.hard_node_mask)
            edge_mask = self._post_process(self.edge_mask, self.hard_edge_mask)
        finally:
            clear_masks(model)
            self.node_mask = self.edge_mask = None
            self.hard_node_mask = self.hard_edge_mask = None
        return Explanation(node_mask=node_mask, edge_mask=edge_mask)


class Explainer:
    def __init__(self, model: torch.nn.Module, algorithm: GNNExplainer, explanation_type="model",
                 model_config=None, node_mask_type=None, edge_mask_type=None, threshold_config=None):
        if isinstance(model_config, dict):
            model_config = ModelConfig(**model_config)
        self.model = model
        self.algorithm = algorithm
        self.explanation_type = getattr(explanation_type, "value", explanation_type)
        self.model_config = model_config or ModelConfig()
        self.node_mask_type = getattr(node_mask_type, "value", node_mask_type)
        self.edge_mask_type = getattr(edge_mask_type, "value", edge_mask_type)
        self.threshold_config = threshold_config
        self.explainer_config = ExplainerConfig(self.explanation_type, self.node_mask_type, self.edge_mask_type)
        self.algorithm.connect(self.explainer_config, self.model_config)

    @torch.no_grad()
    def get_prediction(self, *args, **kwargs) -> torch.Tensor:
        training = self.model.training
        self.model.eval()
        out = self.model(*args, **kwargs)
        self.model.train(training)
        return out

    def get_target(self, prediction: torch.Tensor) -> torch.Tensor:
        mode = self.model_config.mode
        if mode == "binary_classification":
            return (prediction > (0 if self.model_config.return_type == "raw" else 0.5)).long().view(-1)
        if mode == "multiclass_classification":
            return prediction.argmax(dim=-1)
        return prediction

    def __call__(self, x, edge_index, *, target=None, index=None, **kwargs) -> Explanation:
        prediction = None
        if self.explanation_type == "phenomenon":
            if target is None:
                raise ValueError("target must be given for explanation_type='phenomenon'")
        else:
            prediction = self.get_prediction(x, edge_index, **kwargs)
            target = self.get_target(prediction)
        training = self.model.training
        self.model.eval()
        explanation = self.algorithm(self.model, x, edge_index, target=target, index=index, **kwargs)
        self.model.train(training)
        explanation.prediction = prediction
        explanation.target = target
        explanation.index = index
        explanation.x, explanation.edge_index = x, edge_index
        for k, v in kwargs.items():
            setattr(explanation, k, v)
        return explanation


# ------------------------------------------------------------------------------------------------
# per-atom gradient-L2 importance (gnnexplainer.py:640-659, aggregated :1427-1429)
# ------------------------------------------------------------------------------------------------
class frozen_parameters:
    """Context manager: ``requires_grad_(False)`` on every trainable parameter of ``model`` for the duration.
    A custom autograd Function sees ``needs_input_grad`` per input, not per ``autograd.grad`` call, so with trainable
    parameters the weight-gradient GEMMs, bias column sums and attention-vector reductions would be launched and
    thrown away (as the reference's ``prediction.backward()`` does): one third of the backward."""

    def __init__(self, model: torch.nn.Module):
        self._params = [p for p in model.parameters() if p.requires_grad]

    def __enter__(self):
        for p in self._params:
            p.requires_grad_(False)
        return self

    def __exit__(self, *exc):
        for p in self._params:
            p.requires_grad_(True)
        return False


def atom_importance(model: torch.nn.Module, x: torch.Tensor, edge_index: torch.Tensor,
                    batch: torch.Tensor = None) -> torch.Tensor:
    """``|| d sum(pred) / d x_i ||_2`` per atom for a model with the reference's explainable signature
    ``model(x, edge_index, batch)`` (gnnexplainer.py:103-112).  Batched over molecules: valid per molecule because
    GATConv / SAGEConv / the pools never mix molecules."""
    x = x.detach().clone().requires_grad_(True)
    with frozen_parameters(model):
        pred = model(x, edge_index, batch)
        grad, = torch.autograd.grad(pred.sum(), x)
    return torch.norm(grad, dim=1)
