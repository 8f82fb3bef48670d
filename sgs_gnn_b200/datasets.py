"""Drop-in for the one hot-path-adjacent function of the reference's datasets.py: `add_degree`
(datasets.py:141-156), the degree prior `data.prob` the sampler mixes into the edge scores.  The reference
computes it once on the CPU through torch-sparse; here it runs on whatever device `data.edge_index` lives on
(two bincounts, two gathers and a softmax over E -- data preparation, executed once per graph).
Dataset download / METIS clustering (datasets.py, main.py:41-67) are outside the build's scope."""
from __future__ import annotations

import torch


def degree_prior(edge_index, num_nodes):
    """softmax_e( E^-1/2 / (colcount[row_e] + rowcount[col_e] + 1e-10) ), every fp32 operation as the reference
    performs it (datasets.py:147-155).  `edge_index` must be sorted by (row, col) as PyG datasets are (adj.coo()
    returns that order).  Computed by libsgs_b200 kernels (sgs_degree_scores + sgs_softmax_f32); a host tensor -- the
    reference prepares its data on the CPU -- is uploaded, processed on the device and the result copied back.  There
    is no CPU implementation: without a GPU this raises like every other op of the package."""
    from . import ops
    if not edge_index.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("add_degree runs on the GPU (sgs_gnn_b200 has no CPU fallback)")
        return degree_prior(edge_index.cuda(), num_nodes).to(edge_index.device)
    g = ops.graph_of(edge_index, num_nodes)
    e = g.num_edges
    counts = torch.empty(2 * num_nodes, dtype=torch.int32, device=edge_index.device)
    scores = torch.empty(e, dtype=torch.float32, device=edge_index.device)
    ops.check(ops.lib().sgs_degree_scores(ops._p(g.src), ops._p(g.dst), e, int(num_nodes), ops._p(counts),
                                          ops._p(scores), ops._stream()), "sgs_degree_scores")
    return ops.softmax_f32(scores)


def add_degree(data):
    n = data.num_nodes if hasattr(data, "num_nodes") else data.x.size(0)
    ei = data.edge_index
    key = ei[0] * n + ei[1]
    if ei.size(1) > 1 and not bool((key[1:] >= key[:-1]).all()):
        raise RuntimeError("add_degree expects edge_index sorted by (row, col): the reference reads the edges back "
                           "through SparseTensor.coo(), which returns that order")
    data.prob = degree_prior(ei, n)
