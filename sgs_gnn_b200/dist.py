"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL / NVLink).

What shards and what is exchanged (DESIGN.md section "Multi-GPU"):

* Independent batches (the reference's cluster mini-batches, main.py:41-67) are the natural
  data-parallel unit: every rank runs the learned-sparsifier step on its own batch, the
  conditional gate is decided on the summed correct-counts (one 2 x int64 all-reduce) and the
  small weight gradients (~0.5 M floats) are all-reduced before the replicated Adam steps.
  This is the path bench.py measures at N > 1 ("scaling": "weak").

* One large graph can also be sharded by destination-node range (`shard_by_destination`).
  The top-q sampler then runs as a DISTRIBUTED RADIX SELECT: every rank histograms its own
  keys and only the 2048-bin digit histograms are all-reduced (3 x 16 KiB), so all ranks derive
  the same threshold tau; threshold ties are resolved by an exclusive scan of per-rank tie
  counts (lowest rank first == lowest edge ids first when shards are id-ordered), and each rank
  compacts its own selection.  `DistributedTopQ` implements the protocol on top of the
  step-wise C-ABI entry points (sgs_topq_keys / _find / _hist / _compact); the local kernels
  are injected so that the protocol itself is covered by world_size-2 gloo tests on CPU.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


# ------------------------------------------------------------------------------------------
# destination-range edge sharding
# ------------------------------------------------------------------------------------------

def destination_ranges(in_degree, world):
    """Contiguous node ranges [n_r, n_{r+1}) balanced by in-edge count (a hub row is never split).
    Returns a list of world+1 boundaries."""
    n = in_degree.numel()
    csum = torch.cumsum(in_degree.to(torch.int64), 0)
    total = int(csum[-1]) if n > 0 else 0
    bounds = [0]
    for r in range(1, world):
        target = (total * r) // world
        b = int(torch.searchsorted(csum, torch.tensor([target], device=csum.device), right=False)[0])
        bounds.append(max(bounds[-1], min(b, n)))
    bounds.append(n)
    return bounds


def shard_by_destination(edge_index, num_nodes, world, rank):
    """Edge ids (ascending) whose destination lies in this rank's node range, and the ranges."""
    dst = edge_index[1]
    indeg = torch.bincount(dst, minlength=num_nodes)
    bounds = destination_ranges(indeg, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    ids = torch.nonzero((dst >= lo) & (dst < hi)).flatten()
    return ids, bounds


# ------------------------------------------------------------------------------------------
# distributed radix top-q
# ------------------------------------------------------------------------------------------

class CudaTopQOps:
    """Local steps of the select, backed by libsgs_b200 (see include/sgs_b200.h K2)."""

    def __init__(self):
        from . import _lib, ops
        self._lib, self._ops = _lib, ops

    def keys(self, p, prob, noise, mode, coef, S):
        ops, lib = self._ops, self._lib.lib()
        e = p.numel()
        keys = torch.empty(e, dtype=torch.int32, device=p.device)
        hist = torch.empty(self._lib.TOPQ_BINS, dtype=torch.int64, device=p.device)
        state = torch.empty(8, dtype=torch.int64, device=p.device)
        one_m, c = ops._coefs(coef)
        self._lib.check(lib.sgs_topq_keys(ops._p(p), ops._p(prob), ops._p(noise), e, one_m, c, mode, ops._p(S),
                                          ops._p(keys), ops._p(hist), ops._p(state), ops._stream()), "sgs_topq_keys")
        return keys, hist, state

    def find(self, hist, state, k_total, level):
        ops = self._ops
        self._lib.check(self._lib.lib().sgs_topq_find(ops._p(hist), ops._p(state), int(k_total), level, ops._stream()),
                        "sgs_topq_find")

    def hist(self, keys, hist, state, level):
        ops = self._ops
        self._lib.check(self._lib.lib().sgs_topq_hist(ops._p(keys), keys.numel(), ops._p(hist), ops._p(state), level,
                                                      ops._stream()), "sgs_topq_hist")

    def compact(self, keys, state, tie_skip, q_cap):
        ops, lib = self._ops, self._lib.lib()
        e = keys.numel()
        sel = torch.empty(max(q_cap, 1), dtype=torch.int32, device=keys.device)
        n_sel = torch.zeros(1, dtype=torch.int64, device=keys.device)
        ws = ops._ws(lib.sgs_topq_workspace_bytes(e), keys.device)
        self._lib.check(lib.sgs_topq_compact(ops._p(keys), e, ops._p(state), int(tie_skip), ops._p(sel), int(q_cap),
                                             None, ops._p(n_sel), ops._p(ws), ws.numel(), ops._stream()),
                        "sgs_topq_compact")
        return sel[: int(n_sel.item())]


class DistributedTopQ:
    """Global top-q over keys that live on different ranks; only histograms / counts move."""

    def __init__(self, local_ops=None, group=None):
        self.ops = local_ops if local_ops is not None else CudaTopQOps()
        self.group = group

    def _allreduce(self, t):
        if is_dist():
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def global_sum(self, p):
        """S = sum over all ranks of p, accumulated in fp64 and rounded once (identical on every rank)."""
        s = p.sum(dtype=torch.float64).reshape(1)
        self._allreduce(s)
        return s.to(torch.float32)

    def select(self, p, prob, noise, q_total, mode, coef=0.3, S=None):
        """Returns (sel: local ids of this rank's selected edges in ascending order, state)."""
        if S is None and mode != 2:
            S = self.global_sum(p)
        keys, hist, state = self.ops.keys(p, prob, noise, mode, coef, S)
        local_last = None
        for level in range(3):
            if level > 0:
                self.ops.hist(keys, hist, state, level)
            if level == 2:
                local_last = hist.clone()
            self._allreduce(hist)
            self.ops.find(hist, state, q_total, level)
        # threshold ties: ranks take them in rank order (shards are ordered by edge id)
        tau_bin = int(state[2].item()) & 511
        n_eq_local = local_last[tau_bin].reshape(1).clone()
        tie_skip = 0
        if is_dist():
            world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
            counts = [torch.zeros_like(n_eq_local) for _ in range(world)]
            dist.all_gather(counts, n_eq_local, group=self.group)
            tie_skip = int(sum(int(c.item()) for c in counts[:rank]))
        sel = self.ops.compact(keys, state, tie_skip, p.numel())
        return sel, state


# ------------------------------------------------------------------------------------------
# data-parallel helpers
# ------------------------------------------------------------------------------------------

def allreduce_gate(learned_correct, random_correct):
    """Sum the two correct-counts over ranks so every rank takes the same branch."""
    if not is_dist():
        return learned_correct, random_correct
    t = torch.stack([learned_correct, random_correct]).to(torch.float64)
    dist.all_reduce(t)
    return t[0], t[1]


def allreduce_grads(params, average=True):
    """One flat all-reduce of the gradients of `params` (those that have one)."""
    if not is_dist():
        return
    gs = [p.grad for p in params if p.grad is not None]
    if not gs:
        return
    flat = torch.cat([g.reshape(-1) for g in gs])
    dist.all_reduce(flat)
    if average:
        flat /= dist.get_world_size()
    off = 0
    for g in gs:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
