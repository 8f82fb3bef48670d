"""TEST INFRASTRUCTURE ONLY (oracle shim) -- not a product path.

Pure-torch restatement of PyG 2.3.1 ``GCNConv`` with its defaults
(improved=False, cached=False, add_self_loops=True, normalize=True, bias=True,
flow source->target), as used at /root/reference/model.py:94-95,151,153.
The arithmetic lives in a third-party dependency that is not vendored in the
reference (PyTorch-Geometric 2.3.1, pinned in prose at README.md:12-16) and the
reference has no test pinning it: parity at this boundary is UNPINNED.
"""
import math
import torch
import torch.nn as nn


def add_remaining_self_loops(edge_index, edge_weight, num_nodes):
    # PyG utils/loop.py::add_remaining_self_loops, fill_value = 1.0
    row, col = edge_index[0], edge_index[1]
    mask = row != col
    loop_index = torch.arange(num_nodes, dtype=row.dtype, device=row.device)
    loop_index = loop_index.unsqueeze(0).repeat(2, 1)
    if edge_weight is not None:
        loop_w = edge_weight.new_full((num_nodes,), 1.0)
        inv = ~mask
        # existing self loops keep their weight (index assignment: last write wins)
        loop_w[row[inv]] = edge_weight[inv]
        edge_weight = torch.cat([edge_weight[mask], loop_w], dim=0)
    edge_index = torch.cat([edge_index[:, mask], loop_index], dim=1)
    return edge_index, edge_weight


def gcn_norm(edge_index, edge_weight, num_nodes, dtype):
    # PyG nn/conv/gcn_conv.py::gcn_norm (dense edge_index branch)
    if edge_weight is None:
        edge_weight = torch.ones((edge_index.size(1),), dtype=dtype, device=edge_index.device)
    edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, num_nodes)
    row, col = edge_index[0], edge_index[1]
    deg = torch.zeros(num_nodes, dtype=edge_weight.dtype, device=edge_weight.device)
    deg = deg.index_add(0, col, edge_weight)
    dis = deg.pow(-0.5)
    dis = dis.masked_fill(dis == float("inf"), 0.0)
    return edge_index, dis[row] * edge_weight * dis[col]


class _Linear(nn.Module):
    """PyG nn/dense/linear.py::Linear(bias=False, weight_initializer='glorot')."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        a = math.sqrt(6.0 / (in_channels + out_channels))
        nn.init.uniform_(self.weight, -a, a)

    def forward(self, x):
        return x @ self.weight.t()


class GCNConv(nn.Module):
    def __init__(self, in_channels, out_channels, **kwargs):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        # registration order matters for state_dict order: bias first, then lin.weight
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.lin = _Linear(in_channels, out_channels)

    def forward(self, x, edge_index, edge_weight=None):
        n = x.size(0)
        ei, w = gcn_norm(edge_index, edge_weight, n, x.dtype)
        h = self.lin(x)
        msg = h.index_select(0, ei[0]) * w.unsqueeze(-1)
        out = torch.zeros(n, h.size(1), dtype=h.dtype, device=h.device).index_add(0, ei[1], msg)
        return out + self.bias


class _Unavailable(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("out of scope for the oracle shim")


class SAGEConv(nn.Module):
    """Pure-torch restatement of PyG 2.3.1 ``SAGEConv(in, out)`` with its defaults (aggr='mean', root_weight=True,
    bias=True, normalize=False, project=False, flow source->target), as used at /root/reference/model.py:50:
        out_i = lin_l(mean_{j -> i} x_j) + lin_r(x_i),   mean over an empty neighbourhood = 0.
    lin_l = Linear(in, out, bias=True), lin_r = Linear(in, out, bias=False); PyG's default Linear initialisers
    (kaiming_uniform(a=sqrt(5)) weight, uniform(+-1/sqrt(in)) bias) coincide with nn.Linear's, registration order
    lin_l.weight, lin_l.bias, lin_r.weight.  Same caveat as GCNConv: third-party arithmetic, parity UNPINNED."""

    def __init__(self, in_channels, out_channels, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_l = nn.Linear(in_channels, out_channels, bias=True)
        self.lin_r = nn.Linear(in_channels, out_channels, bias=False)

    def forward(self, x, edge_index):
        n = x.size(0)
        src, dst = edge_index[0], edge_index[1]
        agg = torch.zeros(n, x.size(1), dtype=x.dtype, device=x.device).index_add(0, dst, x.index_select(0, src))
        cnt = torch.zeros(n, dtype=x.dtype, device=x.device).index_add(0, dst, torch.ones_like(dst, dtype=x.dtype))
        agg = agg / cnt.clamp(min=1).unsqueeze(-1)
        return self.lin_l(agg) + self.lin_r(x)


GATConv = GINConv = ChebConv = GAT = GIN = _Unavailable
