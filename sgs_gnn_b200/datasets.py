"""Drop-in for the one hot-path-adjacent function of the reference's datasets.py: `add_degree`
(datasets.py:141-156), the degree prior `data.prob` the sampler mixes into the edge scores.  The reference
computes it once on the CPU through torch-sparse; here it runs on whatever device `data.edge_index` lives on
(two bincounts, two gathers and a softmax over E -- data preparation, executed once per graph).
Dataset download / METIS clustering (datasets.py, main.py:41-67) are outside the build's scope."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def degree_prior(edge_index, num_nodes):
    """softmax_e( E^-1/2 / (colcount[row_e] + rowcount[col_e] + 1e-10) ), op order as the reference.
    `edge_index` must be sorted by (row, col) as PyG datasets are (adj.coo() returns that order)."""
    row, col = edge_index[0], edge_index[1]
    e = row.numel()
    colcount = torch.bincount(col, minlength=num_nodes)
    rowcount = torch.bincount(row, minlength=num_nodes)
    deg_in = 1.0 / colcount
    deg_out = 1.0 / rowcount
    prob = (1.0 / deg_in[row]) + (1.0 / deg_out[col])
    prob = 1.0 / (prob + 1e-10)
    return F.softmax(prob * e ** -0.5, dim=0)


def add_degree(data):
    n = data.num_nodes if hasattr(data, "num_nodes") else data.x.size(0)
    ei = data.edge_index
    key = ei[0] * n + ei[1]
    if ei.size(1) > 1 and not bool((key[1:] >= key[:-1]).all()):
        raise RuntimeError("add_degree expects edge_index sorted by (row, col): the reference reads the edges back "
                           "through SparseTensor.coo(), which returns that order")
    data.prob = degree_prior(ei, n)
