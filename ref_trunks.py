"""The reference scripts' model classes -- the CALLERS of the hot path -- re-declared once.

Every reference script re-declares its own copy of these ``nn.Module`` classes next to module-level
code that reads CSV files with RDKit, so they cannot be imported.  They are restated here
parametrised by the operator namespace ``ops`` (anything exposing ``GATConv``, ``SAGEConv``,
``global_max_pool``, ``global_mean_pool``): ``m_gat_graphsage_b200.nn`` for the CUDA path,
``oracle.pyg_oracle`` for the CPU oracle.  Layer names, constructor arguments, forward wiring and
therefore ``state_dict`` keys are the reference's.  ``tests/golden/make_golden.py`` checks these
mirrors against classes extracted from ``/root/reference`` source by ``ast``.

Used by ``tests/``, ``bench.py`` and ``__graft_entry__.smoke()``; not part of the product package.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class Model1Trunk(nn.Module):
    """``GAT_GraphSAGE`` of /root/reference/ablation/model1.py:53-77 -- the north-star model:
    GATConv(35, 35, heads=10) -> ReLU -> SAGEConv(350, 350) -> ReLU -> [max || mean] -> MLP."""

    def __init__(self, ops, n_output=1, num_features_xd=35, output_dim=128, dropout=0.2, heads=10):
        super().__init__()
        self.ops = ops
        self.conv1 = ops.GATConv(num_features_xd, num_features_xd, heads=heads)
        self.conv2 = ops.SAGEConv(num_features_xd * heads, num_features_xd * heads)
        self.fc_g1 = nn.Linear(num_features_xd * heads * 2, 1500)
        self.fc_g2 = nn.Linear(1500, output_dim)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(dropout)
        self.out = nn.Linear(output_dim, n_output)

    def forward(self, data):
        x, edge_index, batch = data.x, data.edge_index, data.batch
        x = self.relu(self.conv1(x, edge_index))
        x = self.relu(self.conv2(x, edge_index))
        x = torch.cat([self.ops.global_max_pool(x, batch), self.ops.global_mean_pool(x, batch)], dim=1)
        x = self.dropout(self.relu(self.fc_g1(x)))
        return self.out(self.fc_g2(x))


class GATNetTrunk(nn.Module):
    """``GATNet`` of /root/reference/gnn/gat.py:51-71 (attention dropout 0.2, second layer H=1, C=128)."""

    def __init__(self, ops, num_features_xd=35, n_output=1, output_dim=128, dropout=0.2):
        super().__init__()
        self.ops = ops
        self.gcn1 = ops.GATConv(num_features_xd, num_features_xd, heads=10, dropout=dropout)
        self.gcn2 = ops.GATConv(num_features_xd * 10, output_dim, dropout=dropout)
        self.fc_g1 = nn.Linear(output_dim, output_dim)
        self.out = nn.Linear(output_dim, n_output)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(dropout)

    def forward(self, data):
        x, edge_index, batch = data.x, data.edge_index, data.batch
        x = F.dropout(x, p=0.2, training=self.training)
        x = F.elu(self.gcn1(x, edge_index))
        x = F.dropout(x, p=0.2, training=self.training)
        x = self.relu(self.gcn2(x, edge_index))
        x = self.ops.global_max_pool(x, batch)
        return self.out(self.relu(self.fc_g1(x)))


class SAGENetTrunk(nn.Module):
    """``SAGENet`` of /root/reference/gnn/graphsage.py:50-75 (pools WITHOUT a preceding ReLU, :67-68)."""

    def __init__(self, ops, num_features_xd=35, n_output=1, output_dim=128, dropout=0.2):
        super().__init__()
        self.ops = ops
        self.sage1 = ops.SAGEConv(num_features_xd, num_features_xd)
        self.sage2 = ops.SAGEConv(num_features_xd, output_dim)
        self.fc_g1 = nn.Linear(output_dim, output_dim)
        self.fc_g2 = nn.Linear(output_dim, output_dim)
        self.out = nn.Linear(output_dim, n_output)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(dropout)

    def forward(self, data):
        x, edge_index, batch = data.x, data.edge_index, data.batch
        x = F.dropout(x, p=0.2, training=self.training)
        x = F.relu(self.sage1(x, edge_index))
        x = F.dropout(x, p=0.2, training=self.training)
        x = self.sage2(x, edge_index)
        x = self.ops.global_max_pool(x, batch)
        x = self.relu(self.fc_g1(x))
        x = F.dropout(x, p=0.2, training=self.training)
        x = self.relu(self.fc_g2(x))
        return self.out(x)


class ModifiedGATLayer(nn.Module):
    """/root/reference/train.py:77-99 (and 26 copies): dense all-pairs attention over every atom of the
    batch; ignores ``edge_index``.  Plain PyTorch in the reference and here."""

    def __init__(self, in_features, out_features):
        super().__init__()
        self.query_transform = nn.Linear(in_features, out_features)
        self.key_transform = nn.Linear(in_features, out_features)
        self.value_transform = nn.Linear(in_features, out_features)
        self.conv3 = nn.Conv1d(out_features, out_features, kernel_size=3, padding=1)
        self.conv5 = nn.Conv1d(out_features, out_features, kernel_size=5, padding=2)
        self.linear_transform = nn.Linear(out_features * 3, out_features)

    def forward(self, x):
        q, k, v = self.query_transform(x), self.key_transform(x), self.value_transform(x)
        k = k.unsqueeze(2)
        k_cat = torch.cat((self.conv3(k), self.conv5(k), k), dim=1)
        k_new = self.linear_transform(k_cat.transpose(1, 2))
        scores = torch.matmul(q, k_new.transpose(1, 2)) / (k_new.size(-1) ** 0.5)
        w = F.softmax(scores.squeeze(-1), dim=-1)
        return torch.matmul(w, v) + v


class TrainTrunk(nn.Module):
    """``GAT_GraphSAGE`` of /root/reference/train.py:102-124 (== test.py:86-108, gnnexplainer.py:78-100):
    ModifiedGATLayer -> ReLU -> SAGEConv(35, 35) -> ReLU -> global_max_pool -> MLP."""

    def __init__(self, ops, n_output=1, num_features_xd=35, output_dim=128, dropout=0.3):
        super().__init__()
        self.ops = ops
        self.conv1 = ModifiedGATLayer(num_features_xd, num_features_xd)
        self.conv2 = ops.SAGEConv(num_features_xd, num_features_xd)
        self.fc_g1 = nn.Linear(num_features_xd, 1500)
        self.fc_g2 = nn.Linear(1500, output_dim)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(dropout)
        self.out = nn.Linear(output_dim, n_output)

    def forward(self, data):
        x, edge_index, batch = data.x, data.edge_index, data.batch
        x = self.relu(self.conv1(x))
        x = self.relu(self.conv2(x, edge_index))
        x = self.ops.global_max_pool(x, batch)
        x = self.dropout(self.relu(self.fc_g1(x)))
        return self.out(self.fc_g2(x))


class CNNNet(nn.Module):
    """ECFP branch of /root/reference/train.py:127-146: three Conv1d over the 1024-bit
    fingerprint, ``fc1 = Linear(128 * 1024, 256)`` (33.5 M of the model's 34.6 M parameters), ``fc2``.  Plain PyTorch in
    the reference and here (stock cuDNN / cuBLAS; ``accel.use_mgs_linear`` can route fc1 / fc2 to the K4 kernels)."""

    def __init__(self, input_dim=1024, output_dim=1024, dropout=0.3):
        super().__init__()
        self.conv1 = nn.Conv1d(in_channels=1, out_channels=32, kernel_size=3, padding="same")
        self.conv2 = nn.Conv1d(in_channels=32, out_channels=64, kernel_size=3, padding="same")
        self.conv3 = nn.Conv1d(in_channels=64, out_channels=128, kernel_size=3, padding="same")
        self.fc1 = nn.Linear(128 * input_dim, 256)
        self.fc2 = nn.Linear(256, output_dim)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(dropout)

    def forward(self, ecfp):
        x = ecfp.squeeze(1).unsqueeze(1)
        x = self.relu(self.conv1(x))
        x = self.relu(self.conv2(x))
        x = self.relu(self.conv3(x))
        x = x.view(x.size(0), -1)
        x = self.dropout(self.relu(self.fc1(x)))
        return self.fc2(x)


class CombinedNet(nn.Module):
    """Fusion head of /root/reference/train.py:149-160: ``[trunk output | CNN output]`` (1 + 1024) -> 512 -> 1."""

    def __init__(self, input_dim=1025, hidden_dim=512, output_dim=1):
        super().__init__()
        self.fc1 = nn.Linear(input_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, output_dim)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(0.3)

    def forward(self, x):
        return self.fc2(self.dropout(self.relu(self.fc1(x))))


def kl_loss(latent):
    """/root/reference/train.py:70-74: KL term on the batch statistics of the fused embedding."""
    mean = torch.mean(latent, dim=0)
    var = torch.var(latent, dim=0)
    return -0.5 * torch.sum(1 + torch.log(var + 1e-10) - mean.pow(2) - var)


class TrainPyModel(nn.Module):
    """The three networks of /root/reference/train.py:212-214 and the step of :236-249 as one module: GNN trunk
    (``GAT_GraphSAGE``) + ECFP ``CNNNet`` + ``CombinedNet``; ``loss = mse + 0.001 * kl_loss(combined)`` (:244-246)."""

    def __init__(self, ops, lambda_kl=0.001):
        super().__init__()
        self.gat_graphsage_model = TrainTrunk(ops)
        self.cnn_model = CNNNet(input_dim=1024, output_dim=1024)
        self.combined_model = CombinedNet(input_dim=1025, hidden_dim=512, output_dim=1)
        self.lambda_kl = lambda_kl

    def forward(self, data, ecfp):
        g = self.gat_graphsage_model(data)
        combined = torch.cat((g, self.cnn_model(ecfp)), dim=1)
        return self.combined_model(combined), combined

    def loss(self, data, ecfp):
        out, combined = self(data, ecfp)
        return F.mse_loss(out, data.y.view(-1, 1)) + self.lambda_kl * kl_loss(combined)


class StressTrunk(nn.Module):
    """BASELINE.json configs[4]: 8-head GAT hidden 256 + GraphSAGE hidden 256 (model1 wiring, wider)."""

    def __init__(self, ops, num_features_xd=35, hidden=256, heads=8, output_dim=128, n_output=1, dropout=0.2):
        super().__init__()
        self.ops = ops
        self.conv1 = ops.GATConv(num_features_xd, hidden // heads, heads=heads)
        self.conv2 = ops.SAGEConv(hidden, hidden)
        self.fc_g1 = nn.Linear(hidden * 2, 1500)
        self.fc_g2 = nn.Linear(1500, output_dim)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(dropout)
        self.out = nn.Linear(output_dim, n_output)

    forward = Model1Trunk.forward


class GCNNetTrunk(nn.Module):
    """``GCNNet`` of /root/reference/gnn/gcn.py:42-66 (35 atom features as everywhere else in the reference's
    data pipeline; the class default of 5 is overridden at gnn/gcn.py construction)."""

    def __init__(self, ops, n_output=1, num_features_xd=35, dropout=0.1):
        super().__init__()
        self.ops = ops
        self.conv1 = ops.GCNConv(num_features_xd, num_features_xd)
        self.conv2 = ops.GCNConv(num_features_xd, num_features_xd * 2)
        self.conv3 = ops.GCNConv(num_features_xd * 2, num_features_xd * 4)
        self.fc_g1 = nn.Linear(num_features_xd * 4, 1024)
        self.fc_g2 = nn.Linear(1024, n_output)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(dropout)

    def forward(self, data):
        x, edge_index = data.x, data.edge_index
        x = self.relu(self.conv1(x, edge_index))
        x = self.relu(self.conv2(x, edge_index))
        x = self.relu(self.conv3(x, edge_index))
        x = self.ops.global_max_pool(x, data.batch)
        x = self.dropout(self.relu(self.fc_g1(x)))
        return self.fc_g2(x)


class GATGCNTrunk(nn.Module):
    """``GAT_GCN`` of /root/reference/gnn/gat-gcn.py:53-76: GATConv(35, 35, heads=10) -> GCNConv(350, 350) ->
    max || mean pool -> MLP."""

    def __init__(self, ops, n_output=1, num_features_xd=35, output_dim=128, dropout=0.2):
        super().__init__()
        self.ops = ops
        self.conv1 = ops.GATConv(num_features_xd, num_features_xd, heads=10)
        self.conv2 = ops.GCNConv(num_features_xd * 10, num_features_xd * 10)
        self.fc_g1 = nn.Linear(num_features_xd * 10 * 2, 1500)
        self.fc_g2 = nn.Linear(1500, output_dim)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(dropout)
        self.out = nn.Linear(output_dim, n_output)

    def forward(self, data):
        x, edge_index, batch = data.x, data.edge_index, data.batch
        x = self.relu(self.conv1(x, edge_index))
        x = self.relu(self.conv2(x, edge_index))
        x = torch.cat([self.ops.global_max_pool(x, batch), self.ops.global_mean_pool(x, batch)], dim=1)
        x = self.dropout(self.relu(self.fc_g1(x)))
        return self.out(self.fc_g2(x))


class GINNetTrunk(nn.Module):
    """``GINConvNet`` of /root/reference/gnn/gin.py:56-104: 5 x (GINConv(MLP) -> ReLU -> BatchNorm1d) ->
    global_add_pool -> MLP."""

    def __init__(self, ops, n_output=1, num_features_xd=35, dropout=0.2):
        super().__init__()
        self.ops = ops
        dim = 32
        self.dropout = nn.Dropout(dropout)
        self.relu = nn.ReLU()
        for k in range(1, 6):
            fin = num_features_xd if k == 1 else dim
            setattr(self, f"conv{k}", ops.GINConv(nn.Sequential(nn.Linear(fin, dim), nn.ReLU(), nn.Linear(dim, dim))))
            setattr(self, f"bn{k}", nn.BatchNorm1d(dim))
        self.fc1_xd = nn.Linear(dim, 128)
        self.fc1 = nn.Linear(128, 1024)
        self.fc2 = nn.Linear(1024, 256)
        self.out = nn.Linear(256, n_output)

    def forward(self, data):
        x, edge_index, batch = data.x, data.edge_index, data.batch
        for k in range(1, 6):
            x = getattr(self, f"bn{k}")(self.relu(getattr(self, f"conv{k}")(x, edge_index)))
        x = self.ops.global_add_pool(x, batch)
        x = self.dropout(self.relu(self.fc1_xd(x)))
        x = self.dropout(self.relu(self.fc1(x)))
        return self.out(self.relu(self.fc2(x)))


class ExplainableWrapper(nn.Module):
    """``ExplainableGATGraphSAGE`` of /root/reference/gnnexplainer.py:103-112: ``forward(x, edge_index, batch)``."""

    def __init__(self, trunk, data_cls):
        super().__init__()
        self.gat_graphsage = trunk
        self._data_cls = data_cls

    def forward(self, x, edge_index, batch=None, edge_attr=None):
        if batch is None:
            batch = torch.zeros(x.size(0), dtype=torch.long, device=x.device)
        return self.gat_graphsage(self._data_cls(x=x, edge_index=edge_index, batch=batch))


TRUNKS = {"model1": Model1Trunk, "gat": GATNetTrunk, "graphsage": SAGENetTrunk, "train": TrainTrunk,
          "stress": StressTrunk, "gcn": GCNNetTrunk, "gat-gcn": GATGCNTrunk, "gin": GINNetTrunk}


def build_trunk(name: str, ops, seed: int = 42, **kwargs) -> nn.Module:
    torch.manual_seed(seed)
    return TRUNKS[name](ops, **kwargs)


def atom_importance(model: nn.Module, data) -> torch.Tensor:
    """Per-atom gradient-L2 importance, /root/reference/gnnexplainer.py:647-652, batched: valid per
    molecule because GATConv / SAGEConv / pools never mix molecules."""
    x = data.x.detach().clone().requires_grad_(True)
    d = type(data)(x=x, edge_index=data.edge_index, batch=data.batch)
    # Only d pred / d x is wanted (the reference's `prediction.backward()` also fills every parameter gradient and
    # throws it away: SURVEY.md section 8 row a10).  A custom autograd Function sees `needs_input_grad` per INPUT,
    # not per autograd.grad() call, so the parameters are frozen for the duration: the weight-gradient GEMMs,
    # bias column sums and attention-vector reductions (1/3 of the backward) are then never launched.
    from m_gat_graphsage_b200.explain import frozen_parameters
    with frozen_parameters(model):
        pred = model(d)
        grad, = torch.autograd.grad(pred.sum(), x)
    return torch.norm(grad, dim=1)
