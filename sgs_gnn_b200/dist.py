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

    def __init__(self, persistent=False):
        """persistent: keep the key / compaction scratch in per-process arenas (one select in flight at a time,
        as in the sharded training step) instead of allocating it per call."""
        from . import _lib, ops
        self._lib, self._ops = _lib, ops
        self._persistent = persistent

    def keys(self, p, prob, noise, mode, coef, S):
        ops, lib = self._ops, self._lib.lib()
        e = p.numel()
        keys = ops._ws(4 * e, p.device, "topq_keys" if self._persistent else None).view(torch.int32)
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

    def compact(self, keys, state, tie_skip, q_cap, n_expected=None):
        ops, lib = self._ops, self._lib.lib()
        e = keys.numel()
        cap = int(q_cap if n_expected is None else n_expected)
        sel = ops._vec(max(cap, 1), torch.int32, keys.device)
        n_sel = torch.zeros(1, dtype=torch.int64, device=keys.device)
        ws = ops._ws(lib.sgs_topq_workspace_bytes(e), keys.device, "topq_ws" if self._persistent else None)
        self._lib.check(lib.sgs_topq_compact(ops._p(keys), e, ops._p(state), int(tie_skip), ops._p(sel), cap,
                                             None, ops._p(n_sel), ops._p(ws), ws.numel(), ops._stream()),
                        "sgs_topq_compact")
        # n_expected comes from the local digit histograms: no host sync needed to size the output
        return sel[:cap] if n_expected is not None else sel[: int(n_sel.item())]


class TopQResult:
    """sel: local ids (int32, ascending) of this rank's selected edges; state: the select's state
    vector (tau bits at [2]); S: the global normaliser; invalid: any rank saw nan/inf/negative input;
    n_global: number of edges selected over all ranks."""
    __slots__ = ("sel", "state", "S", "invalid", "n_global", "tau_bits")


class DistributedTopQ:
    """Global top-q over keys that live on different ranks; only histograms / counts move."""

    SHIFT = (20, 9, 0)
    TIE_GATHER_MAX = 1 << 16

    def __init__(self, local_ops=None, group=None):
        self.ops = local_ops if local_ops is not None else CudaTopQOps()
        self.group = group

    def _allreduce(self, t, op=None):
        if is_dist():
            if t.is_cuda and dist.get_backend(self.group) == "gloo":
                c = t.cpu()
                dist.all_reduce(c, op=op or dist.ReduceOp.SUM, group=self.group)
                t.copy_(c)
            else:
                dist.all_reduce(t, op=op or dist.ReduceOp.SUM, group=self.group)
        return t

    def _allgather(self, t, world):
        staged = t.is_cuda and dist.get_backend(self.group) == "gloo"
        src = t.cpu() if staged else t
        out = torch.empty(world * src.numel(), dtype=src.dtype, device=src.device)
        dist.all_gather_into_tensor(out, src.reshape(-1), group=self.group)
        out = out.view((world,) + tuple(src.shape))
        return out.to(t.device) if staged else out

    def global_sum(self, p):
        """S = sum over all ranks of p, accumulated in fp64 and rounded once (identical on every rank)."""
        if p.is_cuda:      # fp64-accumulating device reduction without a [E] fp64 temporary
            from . import ops
            s = ops.sum_f64(p)
        else:
            s = p.sum(dtype=torch.float64).reshape(1)
        self._allreduce(s)
        return s.to(torch.float32)

    def _tie_cutoff(self, tied_gid, take):
        """Smallest global edge id v such that `take` tied edges over all ranks have id <= v
        (binary search on v; each probe all-reduces one count)."""
        lo_v, hi_v = 0, (1 << 62)
        top = tied_gid.max().reshape(1) if tied_gid.numel() else torch.zeros(1, dtype=torch.int64,
                                                                             device=tied_gid.device)
        self._allreduce(top, dist.ReduceOp.MAX)
        hi_v = int(top.item())
        while lo_v < hi_v:
            mid = (lo_v + hi_v) // 2
            cnt = (tied_gid <= mid).sum().reshape(1)
            self._allreduce(cnt)
            if int(cnt.item()) >= take:
                hi_v = mid
            else:
                lo_v = mid + 1
        return lo_v

    def select_ex(self, p, prob, noise, q_total, mode, coef=0.3, S=None, gid=None):
        """gid: global edge ids of the local edges (ascending).  Threshold ties are broken by lowest
        GLOBAL edge id exactly as on one GPU; without gid, shards are assumed to be contiguous id
        ranges in rank order."""
        if S is None and mode != 2:
            S = self.global_sum(p)
        keys, hist, state = self.ops.keys(p, prob, noise, mode, coef, S)
        local = []
        for level in range(3):
            if level > 0:
                self.ops.hist(keys, hist, state, level)
            local.append(hist.clone())
            self._allreduce(hist)
            self.ops.find(hist, state, q_total, level)
        # one host read: tau, #ties to take, #ties overall, local ties, local #keys > tau, invalid flag
        dev = state.device
        tau_d = state[2] & 0x7FFFFFFF
        ar = torch.arange(local[0].numel(), device=dev)
        n_gt = sum((local[lv] * (ar > ((tau_d >> self.SHIFT[lv]) & (2047 if lv < 2 else 511)))).sum()
                   for lv in range(3))
        c_loc = local[2][(tau_d & 511)]
        bad = state[5:6].clone()
        self._allreduce(bad, dist.ReduceOp.MAX) if is_dist() else None
        host = torch.stack([state[2], state[4], state[6], c_loc, n_gt, bad[0], state[3]]).cpu()
        tau_bits, take, n_eq, c_r, n_gt_loc = (int(host[i]) for i in range(5))
        if n_eq == take or not is_dist():
            t_r = min(c_r, take)
        elif gid is None:
            world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
            counts = torch.zeros(world, dtype=torch.int64)
            counts[rank] = c_r
            self._allreduce(counts)
            before = int(counts[:rank].sum())
            t_r = max(0, min(c_r, take - before))
        else:
            # fp32 keys collide routinely at 10^8 edges: n_eq is a handful.  All-gather the tied global ids
            # (padded to n_eq) and cut at the take-th smallest -- one collective, one more host read.
            eq = keys == (tau_bits & 0x7FFFFFFF)
            try:      # the local tie count is already on the host: fixed-size nonzero, no extra sync
                tied = gid[torch.nonzero_static(eq, size=c_r).flatten()]
            except (RuntimeError, NotImplementedError, AttributeError):
                tied = gid[torch.nonzero(eq).flatten()]
            if n_eq <= self.TIE_GATHER_MAX:
                world = dist.get_world_size(self.group)
                mine = torch.full((n_eq,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=tied.device)
                mine[: tied.numel()] = tied
                every = self._allgather(mine, world)
                cut = torch.sort(every.flatten()).values[take - 1]
                t_r = int((tied <= cut).sum())
            else:
                cut = self._tie_cutoff(tied, take)
                t_r = int((tied <= cut).sum())
        sel = self.ops.compact(keys, state, take - t_r, p.numel(), n_expected=n_gt_loc + t_r)
        r = TopQResult()
        r.sel, r.state, r.S, r.invalid = sel, state, S, bool(host[5])
        r.n_global, r.tau_bits = int(host[6]) + take, tau_bits
        return r

    def select(self, p, prob, noise, q_total, mode, coef=0.3, S=None, gid=None):
        """Returns (sel: local ids of this rank's selected edges in ascending order, state)."""
        r = self.select_ex(p, prob, noise, q_total, mode, coef, S, gid)
        return r.sel, r.state


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
    """One flat all-reduce of the gradients of `params`.  The buffer layout is the same on every rank whatever its
    local has-grad pattern (a rank without a gradient contributes zeros, the has-grad bitmap rides along), so ranks
    can never issue mismatched collectives; a parameter without a gradient on EVERY rank keeps grad = None."""
    if not is_dist():
        return
    ps = [p for p in params if p.requires_grad]
    if not ps:
        return
    # the bitmap is assembled from two device scalars inside the same cat: no host->device copy (a pageable upload
    # synchronises the stream, a pinned one queues behind the loader's prefetch on the copy engine)
    dev = ps[0].device
    one, zero = torch.ones(1, dtype=torch.float32, device=dev), torch.zeros(1, dtype=torch.float32, device=dev)
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(torch.float32)
                      for p in ps] + [zero if p.grad is None else one for p in ps])
    staged = flat.is_cuda and dist.get_backend() == "gloo"
    if staged:
        c = flat.cpu()
        dist.all_reduce(c)
        flat.copy_(c)
    else:
        dist.all_reduce(flat)
    seen = flat[-len(ps):].cpu()
    if average:
        flat /= dist.get_world_size()
    off = 0
    for i, p in enumerate(ps):
        n = p.numel()
        if float(seen[i]) > 0.0:
            if p.grad is None:
                p.grad = flat[off:off + n].view_as(p).clone()
            else:
                p.grad.copy_(flat[off:off + n].view_as(p))
        off += n


def agree_on_path(has_train, big):
    """Data-parallel ranks decide control flow BEFORE any collective of the step (training_hybrid.py:29-30 `continue`
    on a batch without train nodes, :41 the E > q branch): one 4-int all-reduce tells every rank what the others
    would do.  Returns (any_has_train, all_has_train, any_big, all_big)."""
    t = torch.tensor([int(has_train), -int(has_train), int(big), -int(big)], dtype=torch.int64)
    if is_dist():
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = t.to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t = t.cpu()
    return bool(t[0] > 0), bool(-t[1] > 0), bool(t[2] > 0), bool(-t[3] > 0)
