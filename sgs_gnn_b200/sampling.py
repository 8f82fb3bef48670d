"""Drop-in for the reference's sampling.py.

    gumbel_softmax_sampling(batch, edge_probs, edge_index, q=500, temperature=1.0,
                            degree_bias_coef=0.3, log=False, istest=False, epoch=-1)
        -> (BoolTensor[E], Tensor[q])                                     sampling.py:91-155
    random_edge_sampling(edge_index, q)                                   sampling.py:159-163

The live reference function is `torch.multinomial(samples, q, replacement=False)`, i.e. the
top-q of samples / Exp(1)-noise.  Here it is one radix top-q select on the
device (libsgs_b200 K2) with no 2^24 category cap; `temperature`, `log`, `epoch` and
`edge_index` are accepted and unused exactly as in the reference.

Noise: by default drawn on the device from a counter-based generator.  For parity tests the
reference's own Exp(1) tensor can be injected with `inject_noise([...])` (consumed in draw
order), and the normaliser S = sum(p) with `inject_S`.
"""
from __future__ import annotations

import collections

import torch

from . import ops
from .ops import SAMPLE_RAW, SAMPLE_TEST, SAMPLE_TRAIN

_noise_queue = collections.deque()
_S_queue = collections.deque()


def inject_noise(tensors):
    """Queue Exp(1) noise tensors (one per upcoming draw, CUDA float32 [E])."""
    _noise_queue.extend(tensors)


def inject_S(tensors):
    _S_queue.extend(tensors)


def clear_injected():
    _noise_queue.clear()
    _S_queue.clear()


def _next_noise():
    return _noise_queue.popleft() if _noise_queue else None


def _next_S():
    return _S_queue.popleft() if _S_queue else None


def sample_edges(p, prob, q, istest=False, degree_bias_coef=0.3, want_mask=False, validate=True):
    """Internal entry: returns ops.TopQ (sel int32 ascending, optional uint8 mask, state, S)."""
    mode = SAMPLE_TEST if istest else SAMPLE_TRAIN
    return ops.sample_topq(p, prob, q, mode, degree_bias_coef, noise=_next_noise(), S=_next_S(),
                           want_mask=want_mask, validate=validate)


def sample_random(scores, q, validate=True):
    """Baseline draw of training_hybrid.py:45-48 on precomputed scores = softmax(batch.prob).
    Returns edge ids in ascending order (the reference's top-k order is irrelevant to the
    unweighted GCN that consumes them)."""
    return ops.sample_topq(scores, None, q, SAMPLE_RAW, 0.0, noise=_next_noise(), validate=validate)


def gumbel_softmax_sampling(batch, edge_probs, edge_index, q=500, temperature=1.0, degree_bias_coef=0.3,
                            log=False, istest=False, epoch=-1):
    p = edge_probs
    if p.dim() != 1:
        p = p.reshape(-1)
    r = sample_edges(p.detach(), None if istest else batch.prob, q, istest, degree_bias_coef, want_mask=True)
    mask = r.mask.view(torch.bool)
    mode = SAMPLE_TEST if istest else SAMPLE_TRAIN
    if p.requires_grad:
        w = ops.StraightThroughWeightsFn.apply(p, None if istest else batch.prob, r.sel, r.S, mode,
                                               degree_bias_coef)
    else:
        w = ops.gather_selected(p, None if istest else batch.prob, r.sel, mode, degree_bias_coef, r.S, True)[1]
    return mask, w


def random_edge_sampling(edge_index, q):
    num_edges = edge_index.shape[1]
    sampled = torch.randperm(num_edges, device=edge_index.device)[:q]
    return edge_index[:, sampled]
