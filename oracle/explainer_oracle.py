"""CPU oracle for the explainer path of ``/root/reference/gnnexplainer.py`` (TEST INFRASTRUCTURE, see
``oracle/pyg_oracle.py`` for the rules and for the "parity unpinned" statement: the algorithm lives in PyG's
``torch_geometric.explain.algorithm.GNNExplainer``, not under ``/root/reference``).

Restates SURVEY.md Appendix A.4 as two plain functions on the oracle operators:

* ``gnn_explainer``       -- ``Explainer(model, GNNExplainer(epochs, lr), explanation_type='model',
                             node_mask_type='attributes', edge_mask_type='object', ModelConfig('regression',
                             'graph', 'raw'))(x, edge_index, batch=...)`` for ONE molecule
                             (gnnexplainer.py:620-631, 669-673);
* ``gradient_importance`` -- ``simple_gradient_explanation`` (gnnexplainer.py:640-659): ``prediction.backward()``
                             and ``torch.norm(x.grad, dim=1)``.

The masks' starting values are arguments (PyG draws them with ``torch.randn``) so that the CUDA path can be
compared from the same starting point.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

EDGE_SIZE, EDGE_ENT = 0.005, 1.0            # coeffs of torch_geometric GNNExplainer: edge_size (sum), edge_ent (mean)
FEAT_SIZE, FEAT_ENT = 1.0, 0.1              # node_feat_size (mean), node_feat_ent (mean)
EPS = 1e-15


def default_masks(num_nodes: int, num_feats: int, num_edges: int, generator=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """PyG's initialisation: ``node_mask = randn(N, F) * 0.1``; ``edge_mask = randn(E) * std`` with
    ``std = calculate_gain('relu') * sqrt(2 / (2 N))``."""
    std = math.sqrt(2.0) * math.sqrt(2.0 / (2 * num_nodes))
    return (torch.randn(num_nodes, num_feats, generator=generator) * 0.1,
            torch.randn(num_edges, generator=generator) * std)


def _entropy(m: torch.Tensor) -> torch.Tensor:
    return -m * torch.log(m + EPS) - (1.0 - m) * torch.log(1.0 - m + EPS)


def _message_passing_modules(model):
    return [m for m in model.modules() if hasattr(m, "_explain") and hasattr(m, "_edge_mask")]


def gnn_explainer(model: torch.nn.Module, x: torch.Tensor, edge_index: torch.Tensor, *, epochs: int, lr: float,
                  init_node_mask: torch.Tensor, init_edge_mask: torch.Tensor, target: Optional[torch.Tensor] = None,
                  **kwargs):
    """-> ``(node_mask [N, F], edge_mask [E], prediction)`` after ``epochs`` Adam steps (A.4):
    forward on ``x * sigmoid(node_mask)`` with ``sigmoid(edge_mask)`` multiplied into the messages of every
    message-passing layer; loss = mse(y_hat, target) + 0.005 sum(s_e) + mean H(s_e) + mean(s_v) + 0.1 mean H(s_v);
    after the first backward, entries whose gradient is exactly zero are frozen out of the regularisers ("hard
    masks") and are zero in the result."""
    was_training = model.training
    model.eval()
    with torch.no_grad():
        prediction = model(x, edge_index, **kwargs)
    if target is None:
        target = prediction
    node_mask = init_node_mask.detach().clone().requires_grad_(True)
    edge_mask = init_edge_mask.detach().clone().requires_grad_(True)
    layers = _message_passing_modules(model)
    for m in layers:
        m._explain, m._edge_mask, m._apply_sigmoid = True, edge_mask, True
    hard_node = hard_edge = None
    opt = torch.optim.Adam([node_mask, edge_mask], lr=lr)
    try:
        for it in range(epochs):
            opt.zero_grad()
            y_hat = model(x * torch.sigmoid(node_mask), edge_index, **kwargs)
            loss = torch.nn.functional.mse_loss(y_hat, target)
            if hard_edge is not None:
                s = torch.sigmoid(edge_mask[hard_edge])
                loss = loss + EDGE_SIZE * s.sum() + EDGE_ENT * _entropy(s).mean()
            if hard_node is not None:
                s = torch.sigmoid(node_mask[hard_node])
                loss = loss + FEAT_SIZE * s.mean() + FEAT_ENT * _entropy(s).mean()
            loss.backward()
            opt.step()
            if it == 0:
                if edge_mask.grad is None:
                    raise ValueError("Could not compute gradients for edges")
                hard_node, hard_edge = node_mask.grad != 0.0, edge_mask.grad != 0.0
    finally:
        for m in layers:
            m._explain, m._edge_mask, m._apply_sigmoid = False, None, True
        model.train(was_training)
    out_node, out_edge = torch.sigmoid(node_mask.detach()), torch.sigmoid(edge_mask.detach())
    if hard_node is not None:
        out_node[~hard_node] = 0.0
        out_edge[~hard_edge] = 0.0
    return out_node, out_edge, prediction


def gradient_importance(model: torch.nn.Module, x: torch.Tensor, edge_index: torch.Tensor,
                        batch: Optional[torch.Tensor] = None) -> torch.Tensor:
    """gnnexplainer.py:640-659 for one molecule (or a batch of them when the model never mixes molecules):
    ``x.requires_grad_(True); prediction = model(x, edge_index, batch); prediction.backward();
    node_importance = torch.norm(x.grad, dim=1)``."""
    x = x.detach().clone().requires_grad_(True)
    if batch is None:
        batch = torch.zeros(x.size(0), dtype=torch.long)
    prediction = model(x, edge_index, batch)
    prediction.sum().backward()
    return torch.norm(x.grad, dim=1)
