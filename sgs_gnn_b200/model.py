"""Drop-in for the reference's model.py (same class names, constructor signatures, forward
signatures and state_dict keys -- SURVEY.md 8b), computed by libsgs_b200 kernels.

    GNNModel(in_channels, hidden_dim, num_classes, dropout_prob=0.3, edge_mlp_type='MLP')   model.py:147-164
      .edge_prob_mlp(node_features, edge_index, random_sampled_edge_index=None,
                     use_checkpoint=False) -> [E,1]                                          model.py:102-133
      .forward(data, edge_index, edge_weight=None) -> [N,C]                                  model.py:155-164

`edge_index` arguments may be the reference's int64 [2,E] tensors or ops.Graph objects.
`use_checkpoint` is accepted and ignored: the fused scorer never materialises the [E,2H]
tensor checkpointing exists to avoid, and its backward always recomputes.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops


class _Lin(nn.Module):
    """PyG dense Linear(bias=False, weight_initializer='glorot'): weight [out, in]."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        a = math.sqrt(6.0 / (in_channels + out_channels))
        nn.init.uniform_(self.weight, -a, a)


class GCNConv(nn.Module):
    """PyG 2.3.1 GCNConv(in, out) with its defaults.  Parameter names `bias`,
    `lin.weight` as in PyG so the reference's name-filtered optimisers (main.py:100,122) and
    state_dict save/load (main.py:231,264) keep working."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.lin = _Lin(in_channels, out_channels)

    def forward(self, x, edge_index, edge_weight=None, relu=False, p_drop=0.0, seed=0):
        g = ops.graph_of(edge_index, x.size(0))
        return ops.gcn_conv(x, self.lin.weight, self.bias, g, edge_weight, relu, p_drop, seed)


class _EdgeProbBase(nn.Module):
    def _drop(self):
        return float(self.dropout.p) if self.training else 0.0

    def score(self, out, graph, ids=None, precomputed=None, seed=None, precision=None):
        """_edge_score (model.py:115-122) on node embeddings `out` for all edges of `graph` or
        the int32 id subset `ids`; returns [n] (not [n,1])."""
        seed = ops.next_seed() if seed is None else seed
        return ops.edge_score(out, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, graph, ids,
                              self._drop(), seed, precomputed, precision)

    def forward(self, node_features, edge_index, random_sampled_edge_index=None, use_checkpoint=False):
        profiler = getattr(self, "gpu_profiler", None)
        n = node_features.size(0)
        ops.seg_begin(profiler, "edge_mlp_pre")
        g_full = ops.graph_of(edge_index, n)
        g_msg = g_full if random_sampled_edge_index is None else ops.graph_of(random_sampled_edge_index, n)
        out = self.embed(node_features, g_msg)
        ops.seg_end(profiler, "edge_mlp_pre")
        ops.seg_begin(profiler, "edge_score")
        self.last_seed = ops.next_seed()
        prob = self.score(out, g_full, seed=self.last_seed)
        ops.seg_end(profiler, "edge_score")
        return prob.unsqueeze(-1)


class EdgeProbGCN(_EdgeProbBase):
    """model.py:91-133."""

    def __init__(self, in_channels, hidden_dim, dropout_prob=0.2):
        super().__init__()
        self.gcn1 = GCNConv(in_channels, hidden_dim)
        self.gcn2 = GCNConv(hidden_dim, hidden_dim)
        self.fc1 = nn.Linear(2 * hidden_dim, hidden_dim)
        self.dropout = nn.Dropout(dropout_prob)
        self.fc2 = nn.Linear(hidden_dim, 1)

    def embed(self, x, graph):
        """out = relu(gcn2(dropout(relu(gcn1(x, g))), g))   (model.py:106-111)"""
        h = self.gcn1(x, graph, None, relu=True, p_drop=self._drop(), seed=ops.next_seed())
        return self.gcn2(h, graph, None, relu=True)


class EdgeProbMLP(_EdgeProbBase):
    """model.py:8-45.  Only the `random_sampled_edge_index=None` form is shape-consistent in the
    reference; the per-edge F->H projection is evaluated once per node
    (relu(fcdim(X)) then gather), which is identical whenever dropout is off; with dropout the
    keep mask is per node rather than per edge endpoint."""

    def __init__(self, in_channels, hidden_dim, dropout_prob=0.2):
        super().__init__()
        self.dropout = nn.Dropout(dropout_prob)
        self.fcdim = nn.Linear(in_channels, hidden_dim)
        self.fc1 = nn.Linear(2 * hidden_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, 1)

    def embed(self, x, graph):
        n = x.size(0)
        empty = getattr(self, "_empty_graph", None)
        if empty is None or empty.num_nodes != n or empty.device != x.device:
            z = torch.empty(0, dtype=torch.int32, device=x.device)
            empty = ops.Graph(z, z.clone(), n)
            self._empty_graph = empty
        # edge-less GCN layer == relu(x W^T + b) per node (deg = loop weight = 1)
        return ops.gcn_conv(x, self.fcdim.weight, self.fcdim.bias, empty, None, True, self._drop(), ops.next_seed())

    def forward(self, node_features, edge_index, random_sampled_edge_index=None, use_checkpoint=False):
        if random_sampled_edge_index is not None:
            raise RuntimeError("EdgeProbMLP scores only the edges it is given: with "
                               "random_sampled_edge_index it returns [q] probabilities, which the sampler "
                               "cannot combine with batch.prob [E] (use --conditional False "
                               "--sparse_edge_mlp False, or --edge_mlp_type GCN)")
        return super().forward(node_features, edge_index, None, use_checkpoint)


class SAGEConv(nn.Module):
    """PyG 2.3.1 SAGEConv(in, out) with its defaults (mean aggregation, root weight, bias on lin_l only).  Parameter
    names `lin_l.weight`, `lin_l.bias`, `lin_r.weight` and nn.Linear-equivalent initialisers as in PyG."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_l = nn.Linear(in_channels, out_channels, bias=True)
        self.lin_r = nn.Linear(in_channels, out_channels, bias=False)

    def forward(self, x, edge_index, relu=False, p_drop=0.0, seed=0):
        g = ops.graph_of(edge_index, x.size(0))
        return ops.sage_conv(x, self.lin_l.weight, self.lin_l.bias, self.lin_r.weight, g, relu, p_drop, seed)


class EdgeProbSAGE(_EdgeProbBase):
    """model.py:47-89: ONE SAGEConv layer (registered as `gcn1`, so main.py:100's 'gcn' name filter picks it up as
    the reference does), dropout on its output, then the shared _edge_score."""

    def __init__(self, in_channels, hidden_dim, dropout_prob=0.2):
        super().__init__()
        self.gcn1 = SAGEConv(in_channels, hidden_dim)
        self.fc1 = nn.Linear(2 * hidden_dim, hidden_dim)
        self.dropout = nn.Dropout(dropout_prob)
        self.fc2 = nn.Linear(hidden_dim, 1)

    def embed(self, x, graph):
        """out = dropout(relu(gcn1(x, g)))   (model.py:63/66)"""
        return self.gcn1(x, graph, relu=True, p_drop=self._drop(), seed=ops.next_seed())


def get_edge_mlp(in_channels, hidden_dim, dropout_prob, edge_mlp_type="MLP"):
    """model.py:135-145."""
    if edge_mlp_type == "MLP":
        return EdgeProbMLP(in_channels, hidden_dim, dropout_prob)
    if edge_mlp_type == "GSAGE":
        return EdgeProbSAGE(in_channels, hidden_dim, dropout_prob)
    if edge_mlp_type == "GCN":
        return EdgeProbGCN(in_channels, hidden_dim, dropout_prob)
    raise NotImplementedError(edge_mlp_type)


class GNNModel(nn.Module):
    """model.py:147-164."""

    def __init__(self, in_channels, hidden_dim, num_classes, dropout_prob=0.3, edge_mlp_type="MLP"):
        super().__init__()
        self.edge_prob_mlp = get_edge_mlp(in_channels, hidden_dim, dropout_prob, edge_mlp_type)
        self.gcn1 = GCNConv(in_channels, hidden_dim)
        self.dropout = nn.Dropout(dropout_prob)
        self.gcn2 = GCNConv(hidden_dim, num_classes)

    def forward(self, data, edge_index, edge_weight=None):
        profiler = getattr(self, "gpu_profiler", None)
        ops.seg_begin(profiler, "gnn_forward")
        x = data.x if hasattr(data, "x") else data
        g = ops.graph_of(edge_index, x.size(0))
        p_drop = float(self.dropout.p) if self.training else 0.0
        h = self.gcn1(x, g, edge_weight, relu=True, p_drop=p_drop, seed=ops.next_seed())
        out = self.gcn2(h, g, edge_weight)
        ops.seg_end(profiler, "gnn_forward")
        return out
