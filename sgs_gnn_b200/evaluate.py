"""Drop-in for the reference's evaluate.py: `evaluate` (evaluate.py:6-68) and `ensemble_evaluate`
(evaluate.py:70-173; called every epoch from main.py:218 and at main.py:269), same signatures and
return tuple (train_f1, val_f1, test_f1).

Differences that are deliberate:
  * eval mode has no dropout, so the edge probabilities of a batch do not change between the
    `num_samples_eval` ensemble members: the reference recomputes them for every member
    (full-graph 2-layer GCN + scoring of all E edges, x11), here they are computed once per batch
    and only the Exp(1) noise is redrawn -- the results are identical;
  * micro-F1 of single-label predictions is accuracy and is counted on the device
    (utils.calculate_f1), one small D2H per mask instead of sklearn on the host.
"""
from __future__ import annotations

import torch

from . import ops, sampling
from ._train_core import _softmax_prob
from .ops import SAMPLE_TEST
from .utils import calculate_f1


def _member_logits(args, model, batch, q, mode, cache):
    """Logits of one ensemble member for `batch` (already on the device)."""
    n = batch.x.size(0)
    ei = batch.edge_index
    e = ei.shape[1]
    if mode == "learned":
        if e > q:
            g_full = ops.graph_of(ei, n)
            if "p_full" not in cache:
                # model.edge_prob_mlp(batch.x, batch.edge_index): message passing over the FULL graph
                cache["p_full"] = model.edge_prob_mlp(batch.x, g_full).reshape(-1)
            p_full = cache["p_full"]
            r = sampling.sample_edges(p_full, None, q, True, args.degree_bias_coef, validate=True)
            # sampled_edge_weight = (p * st)[mask].clamp(0, 1), istest (sampling.py:94,137-155)
            w = ops.gather_selected(p_full, None, r.sel, SAMPLE_TEST, args.degree_bias_coef, r.S, True)[1]
            return model(batch, g_full.subgraph(r.sel, ascending=True), w)
        return model(batch, ei)
    if mode == "random":
        if e > q:
            return model(batch, sampling.random_edge_sampling(ei, q=q))
        return model(batch, ei)
    if mode == "edge":
        if e > q:
            g_full = ops.graph_of(ei, n)
            r = sampling.sample_random(_softmax_prob(batch.prob), q)
            return model(batch, g_full.subgraph(r.sel, ascending=True))
        return model(batch, ei)
    if mode == "full":
        return model(batch, ei)
    raise ValueError("Invalid mode. Choose 'learned', 'random', or 'full'.")


def _accumulate(out, batch, totals):
    for k, name in enumerate(("train_mask", "val_mask", "test_mask")):
        mask = getattr(batch, name, None)
        if mask is None:
            continue
        cnt = int(mask.sum().item())
        if cnt > 0:
            totals[k][0] += calculate_f1(out, batch.y, mask) * cnt
            totals[k][1] += cnt


def _finish(totals):
    return tuple(t[0] / t[1] if t[1] > 0 else 0 for t in totals)


def evaluate(args, model, cluster_loader, device, q=500, mode=None, temperature=1.0):
    model.eval()
    totals = [[0.0, 0], [0.0, 0], [0.0, 0]]
    with torch.no_grad():
        for batch in cluster_loader:
            batch = batch.to(device)
            out = _member_logits(args, model, batch, q, mode, {})
            _accumulate(out, batch, totals)
    return _finish(totals)


def ensemble_evaluate(args, model, cluster_loader, device, q=500, mode=None, temperature=1.0):
    model.eval()
    totals = [[0.0, 0], [0.0, 0], [0.0, 0]]
    with torch.no_grad():
        for batch in cluster_loader:
            batch = batch.to(device)
            cache = {}
            acc = None
            for _ in range(args.num_samples_eval):
                out = _member_logits(args, model, batch, q, mode, cache)
                acc = out.clone() if acc is None else acc.add_(out)
            out = acc / float(args.num_samples_eval)      # torch.mean(torch.stack(outs), dim=0)
            _accumulate(out, batch, totals)
    return _finish(totals)
