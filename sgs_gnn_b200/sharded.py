"""One graph sharded across GPUs by destination-node range (BASELINE.json north_star).

Rank r owns the CSR-by-destination rows [n_r, n_{r+1}) and every edge pointing into them, with its
`prob`, noise and probability.  Node features, labels, masks and all weights are replicated; node
ids stay GLOBAL, so the single-GPU kernels run unchanged on the local edge list.  What crosses
NVLink per learned step:

  * after a destination-sharded SpMM whose output is gathered by arbitrary sources
    (scorer embeddings `out` [N,H], layer-2 pre-activations [N,D], logits [N,C]):
    an all-gather of the owned row slabs (`Comm.exchange_rows`);
  * in the backward, the by-source SpMM / scorer / loss produce partial sums for arbitrary rows:
    a reduce-scatter back to the owned slabs (`Comm.reduce_rows`);
  * gcn_norm: the owned entries of deg / dis / loop weight ([N,3]) are all-gathered, the
    edge-weight gradient all-reduces one [N] vector (SURVEY A.3's source-side sums);
  * the top-q sampler all-reduces only 3 digit histograms (16 KiB each) + a few scalars
    (`dist.DistributedTopQ`); the selected set is identical for any number of ranks;
  * loss numerators / counts (16 doubles) and, at the end, the ~0.5 M weight gradients, which
    every rank holds as partial sums over its rows / edges (all-reduce SUM, no averaging).

Dense X.W^T on replicated inputs is computed redundantly on every rank (it is ~0.2 ms of tensor
time at Reddit scale and saves an exchange); dense layers on sharded activations run on the owned
slab only.  Reference call sites: training_hybrid.py:39-147, model.py:102-122,155-164.
"""
from __future__ import annotations

import os
import weakref

import torch
import torch.distributed as dist

from . import dist as sdist
from . import ops, sampling
from ._lib import PREC_FP32, SAMPLE_RAW, SAMPLE_TRAIN, check, lib
from .ops import _p, _req, _stream, _timed


# ------------------------------------------------------------------------------------------
# collectives over row slabs
# ------------------------------------------------------------------------------------------

class PeerArena:
    """The symmetric arena of csrc/peer.cu (torch.distributed._symmetric_memory is the plumbing: same-size buffers on
    every GPU, all mapped into every process, plus a cross-GPU barrier on the stream).  Four [N, D]-sized regions:
    two alternating STAGES (destinations of the producers' peer stores of a slab all-gather) and two alternating
    PARTIALS (sources of the peer loads of a slab reduce-scatter).  Alternation makes ONE barrier per exchange enough:
    a region is rewritten two exchanges later, by when every rank has passed a barrier it could only reach after it
    had finished reading the region (single stream, program order)."""

    def __init__(self, comm, device):
        import torch.distributed._symmetric_memory as symm_mem
        self._sm = symm_mem
        self.comm = comm
        self.device = device
        self.cap = 0
        self.arena = self.hdl = None
        self._flip = {"stage": 0, "part": 0}

    def _ensure(self, nfloats):
        need = (int(nfloats) + 1023) // 1024 * 1024
        if need <= self.cap:
            return
        # collective (every rank asks for the same global [N, D] sizes in the same order)
        self.cap = need
        self.arena = self._sm.empty(4 * need, dtype=torch.float32, device=self.device)
        group = self.comm.group if self.comm.group is not None else dist.group.WORLD
        self.hdl = self._sm.rendezvous(self.arena, group)
        self.bases = self.hdl.buffer_ptrs_dev
        self.hdl.barrier(channel=0)

    def region(self, kind, n, d):
        """(tensor view [n, d] of the local arena, element offset) of the next `kind` region."""
        self._ensure(n * d)
        k = self._flip[kind]
        self._flip[kind] = k ^ 1
        off = ((0 if kind == "stage" else 2) + k) * self.cap
        return self.arena[off:off + n * d].view(n, d), off

    def barrier(self):
        self.hdl.barrier(channel=0)


class Comm:
    """Row-slab collectives.  On the GPU box (NCCL backend) the slab all-gathers / reduce-scatters of the sharded
    GCN layers run over NVLink PEER MEMORY with hand-written kernels (csrc/peer.cu, the SpMM epilogue's peer stores);
    NCCL keeps the small scalar / histogram / weight-gradient all-reduces.  In the tests (gloo) everything is staged
    through host memory."""

    def __init__(self, group=None):
        self.group = group
        if sdist.is_dist():
            self.world = dist.get_world_size(group)
            self.rank = dist.get_rank(group)
            self.backend = dist.get_backend(group)
        else:
            self.world, self.rank, self.backend = 1, 0, "none"
        self.staged = self.backend == "gloo"
        self.uneven_native = os.environ.get("SGS_SHARD_UNEVEN", "native") == "native"
        self.peer = None
        if self.world > 1 and self.backend == "nccl" and os.environ.get("SGS_PEER", "1") != "0":
            self.peer = PeerArena(self, torch.device("cuda", torch.cuda.current_device()))

    # ---- peer-memory slab exchange ----
    def peer_stage(self, n, d):
        """Destination of a fused slab all-gather: (stage view, elem_off) or None when peer memory is not in use."""
        if self.peer is None:
            return None
        return self.peer.region("stage", n, d)

    def peer_finish_exchange(self, full, bounds, stage, pushed):
        """Second half of a slab all-gather over peer memory.  pushed: the producer already stored this rank's rows
        into the peers' stages (SpMM epilogue); otherwise they are pushed here.  Then: barrier, copy the foreign rows
        out of the local stage."""
        st, off = stage
        lo, hi = bounds[self.rank], bounds[self.rank + 1]
        d = full.size(1)
        with _timed("comm_exchange"):
            if not pushed and hi > lo:
                check(lib().sgs_peer_push_rows(_p(full[lo:hi]), self.peer.bases, self.world, self.rank, off, lo,
                                               hi - lo, d, 0, _stream()), "sgs_peer_push_rows")
            self.peer.barrier()
            if lo > 0:
                full[:lo].copy_(st[:lo])
            if hi < full.size(0):
                full[hi:].copy_(st[hi:])
        return full

    def peer_reduce_rows(self, part, off, bounds):
        """Sum over ranks of the partial [N, D] every rank left in its PARTIAL region -> this rank's owned rows."""
        lo, hi = bounds[self.rank], bounds[self.rank + 1]
        d = part.size(1)
        out = torch.empty(hi - lo, d, dtype=torch.float32, device=part.device)
        with _timed("comm_reduce"):
            self.peer.barrier()
            if hi > lo:
                check(lib().sgs_peer_reduce_rows(self.peer.bases, self.world, self.rank, off, lo, hi - lo, d, _p(out),
                                                 _stream()), "sgs_peer_reduce_rows")
        return out

    def all_reduce(self, t):
        if self.world == 1:
            return t
        if self.peer is not None and t.dtype == torch.float32 and t.numel() >= 4096 and t.is_contiguous():
            # fp32 vectors (the edge-weight gradient's per-node sums, the flat weight gradients): every rank leaves its
            # copy in the arena and sums all `world` copies itself with peer loads -- a latency-bound NCCL ring of 8
            # becomes copy + barrier + one reduction kernel
            n = t.numel()
            part, off = self.peer.region("part", n, 1)
            with _timed("comm_allreduce"):
                part.view(-1).copy_(t.reshape(-1))
                self.peer.barrier()
                check(lib().sgs_peer_reduce_rows(self.peer.bases, self.world, self.rank, off, 0, n, 1,
                                                 _p(t), _stream()), "sgs_peer_reduce_rows")
            return t
        with _timed("comm_allreduce"):
            if self.staged and t.is_cuda:
                c = t.cpu()
                dist.all_reduce(c, group=self.group)
                t.copy_(c)
            else:
                dist.all_reduce(t, group=self.group)
        return t

    def exchange_rows(self, full, bounds):
        """In place: afterwards every rank's owned row slab of `full` holds its owner's values."""
        if self.world == 1:
            return full
        if full.dim() == 2 and full.dtype == torch.float32 and full.is_contiguous():
            stage = self.peer_stage(full.size(0), full.size(1))
            if stage is not None:
                return self.peer_finish_exchange(full, bounds, stage, pushed=False)
        sizes = [bounds[i + 1] - bounds[i] for i in range(self.world)]
        lo, hi = bounds[self.rank], bounds[self.rank + 1]
        even = len(set(sizes)) == 1
        with _timed("comm_exchange"):
            if self.staged or not (even or self.uneven_native):
                mx = max(sizes)
                pad = torch.zeros((mx,) + tuple(full.shape[1:]), dtype=full.dtype,
                                  device="cpu" if self.staged else full.device)
                pad[: hi - lo].copy_(full[lo:hi])
                outs = [torch.empty_like(pad) for _ in range(self.world)]
                dist.all_gather(outs, pad, group=self.group)
                for r in range(self.world):
                    if r != self.rank and sizes[r] > 0:
                        full[bounds[r]:bounds[r + 1]].copy_(outs[r][: sizes[r]])
            elif even:
                dist.all_gather_into_tensor(full, full[lo:hi], group=self.group)     # in place: no staging copy
            else:
                dist.all_gather(list(full.split(sizes)), full[lo:hi].clone(), group=self.group)
        return full

    def upload_replicated(self, host, dev):
        """Device copy of a host tensor that every rank holds (the replicated node features): each rank sends only
        its 1/world row slice over PCIe and the slices are all-gathered over NVLink -- the host link carries the
        tensor once per box instead of once per GPU."""
        n = host.size(0)
        if self.world == 1 or self.staged or n < self.world:
            return host.to(dev)
        chunk = (n + self.world - 1) // self.world
        lo, hi = min(self.rank * chunk, n), min((self.rank + 1) * chunk, n)
        full = torch.empty((self.world * chunk,) + tuple(host.shape[1:]), dtype=host.dtype, device=dev)
        mine = full[self.rank * chunk:(self.rank + 1) * chunk]
        mine[: hi - lo].copy_(host[lo:hi], non_blocking=True)
        with _timed("comm_exchange"):
            dist.all_gather_into_tensor(full, mine.clone(), group=self.group)
        if os.environ.get("SGS_CHECK_UPLOAD"):   # debug aid: the gathered copy must equal a plain full upload
            assert torch.equal(full[:n], host.to(dev)), "upload_replicated mismatch"
        return full[:n]

    def reduce_rows(self, full, bounds):
        """Sum of `full` over ranks, returned for this rank's owned rows only ([hi-lo, ...])."""
        lo, hi = bounds[self.rank], bounds[self.rank + 1]
        if self.world == 1:
            return full[lo:hi]
        if self.peer is not None and full.dim() == 2 and full.dtype == torch.float32:
            part, off = self.peer.region("part", full.size(0), full.size(1))
            with _timed("comm_reduce"):
                part.copy_(full)
            return self.peer_reduce_rows(part, off, bounds)
        sizes = [bounds[i + 1] - bounds[i] for i in range(self.world)]
        even = len(set(sizes)) == 1
        with _timed("comm_reduce"):
            if self.staged or not (even or self.uneven_native):
                c = full.cpu() if self.staged else full.clone()
                dist.all_reduce(c, group=self.group)
                return c[lo:hi].to(full.device).contiguous()
            out = torch.empty((hi - lo,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
            if even:
                dist.reduce_scatter_tensor(out, full, group=self.group)
            else:
                dist.reduce_scatter(out, list(full.split(sizes)), group=self.group)
        return out


# ------------------------------------------------------------------------------------------
# the shard of a batch / of an edge list
# ------------------------------------------------------------------------------------------

class LocalGraph:
    """The local edges of a destination-sharded edge list (global node ids)."""

    def __init__(self, graph, bounds, comm, gid=None, num_edges_global=None):
        self.graph = graph
        self.bounds = list(bounds)
        self.comm = comm
        self.lo, self.hi = self.bounds[comm.rank], self.bounds[comm.rank + 1]
        self.gid = gid                      # int64 [E_local] global edge ids (ascending) or None
        self.num_edges_global = num_edges_global
        self._norm_unw = None
        self._norm_w = None

    @property
    def num_nodes(self):
        return self.graph.num_nodes

    def subgraph(self, ids):
        gid = self.gid[ids.long()] if self.gid is not None else None
        return LocalGraph(self.graph.subgraph(ids, ascending=True), self.bounds, self.comm, gid)

    def norm(self, edge_weight=None):
        if edge_weight is None:
            if self._norm_unw is None:
                self._norm_unw = ShardedNorm(self, None)
            return self._norm_unw
        c = self._norm_w
        if c is not None and c[0] is edge_weight and c[1] == edge_weight._version:
            return c[2]
        w = _req(edge_weight.detach(), torch.float32, "edge_weight")
        if w.numel() != self.graph.num_edges:
            raise RuntimeError("edge_weight must have one entry per (local) edge")
        nrm = ShardedNorm(self, w)
        self._norm_w = (edge_weight, edge_weight._version, nrm)
        return nrm


class ShardedNorm:
    """gcn_norm of a destination-sharded edge list: the weighted in-degree of an owned
    row is complete locally; dis[src] of arbitrary sources comes from one [N,3] slab all-gather."""

    def __init__(self, lg, w):
        g = lg.graph
        n, m = g.num_nodes, g.num_edges
        dev = g.device
        rowptr, perm, nbr, _ = g.csr_dst
        deg = torch.empty(n, dtype=torch.float32, device=dev)
        dis = torch.empty_like(deg)
        loopw = torch.empty_like(deg)
        self.what_dst = ops._vec(max(m, 1), torch.float32, dev)
        check(lib().sgs_gcn_norm(_p(rowptr), _p(perm), _p(nbr), _p(w), m, n, _p(deg), _p(dis), _p(loopw),
                                 _p(self.what_dst), _stream()), "sgs_gcn_norm")
        if lg.comm.world > 1:
            stats = torch.stack([deg, dis, loopw], 1).contiguous()
            lg.comm.exchange_rows(stats, lg.bounds)
            deg, dis, loopw = (stats[:, i].contiguous() for i in range(3))
            check(lib().sgs_gcn_norm_apply(_p(rowptr), _p(perm), _p(nbr), _p(w), _p(dis), m, n, _p(self.what_dst),
                                           _stream()), "sgs_gcn_norm_apply")
        self.deg, self.dis, self.loopw = deg, dis, loopw
        self._what_src = None
        self._lg, self._w = weakref.ref(lg), w   # weak: lg -> norm -> lg would only be freed by the cyclic GC

    @property
    def what_src(self):
        if self._what_src is None:
            lg = self._lg()
            if lg is None:
                raise RuntimeError("the graph of this gcn_norm has been released")
            g = lg.graph
            rowptr, perm, nbr, _ = g.csr_src
            self._what_src = ops._vec(max(g.num_edges, 1), torch.float32, g.device)
            check(lib().sgs_gcn_norm_apply(_p(rowptr), _p(perm), _p(nbr), _p(self._w), _p(self.dis), g.num_edges,
                                           g.num_nodes, _p(self._what_src), _stream()), "sgs_gcn_norm_apply")
        return self._what_src


class ShardedBatch:
    """The per-rank shard of a batch: replicated x / y / masks, local edges + prob + the baseline
    draw's scores softmax(prob) (computed over the GLOBAL edge list before slicing, so every rank
    count yields bit-identical keys).  Duck-types the reference's batch (`.to(device)`)."""

    _TENSORS = ("x", "y", "train_mask", "val_mask", "test_mask", "edge_index", "prob", "scores", "gid",
                "train_mask_owned")

    def __init__(self, batch=None, comm=None):
        self.comm = comm
        self._local = None
        if batch is None:
            return
        n = batch.x.size(0)
        e = batch.edge_index.size(1)
        ids, bounds = sdist.shard_by_destination(batch.edge_index, n, comm.world, comm.rank)
        self.x, self.y = batch.x, batch.y
        self.train_mask = batch.train_mask
        self.val_mask = getattr(batch, "val_mask", None)
        self.test_mask = getattr(batch, "test_mask", None)
        self.num_classes = getattr(batch, "num_classes", None)
        self.num_edges_global = e
        self.bounds = bounds
        self.gid = ids
        self.edge_index = batch.edge_index[:, ids].contiguous()
        self.prob = batch.prob[ids].contiguous()
        self.scores = ops.softmax_f32(batch.prob)[ids].contiguous()
        lo, hi = bounds[comm.rank], bounds[comm.rank + 1]
        owned = torch.zeros(n, dtype=torch.bool, device=batch.x.device)
        owned[lo:hi] = True
        self.train_mask_owned = (batch.train_mask & owned).contiguous()
        self._sgs_has_train = bool(batch.train_mask.any())

    @property
    def local(self):
        if self._local is None:
            self._local = LocalGraph(ops.graph_of(self.edge_index, self.x.size(0)), self.bounds, self.comm,
                                     self.gid, self.num_edges_global)
        return self._local

    def _map(self, fn, fn_x=None):
        o = ShardedBatch(None, self.comm)
        for k in self._TENSORS:
            v = getattr(self, k, None)
            setattr(o, k, ((fn_x if (k == "x" and fn_x is not None) else fn)(v)) if v is not None else None)
        o.num_classes, o.num_edges_global, o.bounds = self.num_classes, self.num_edges_global, self.bounds
        o._sgs_has_train = self._sgs_has_train
        return o

    def to(self, device, non_blocking=False):
        dev = torch.device(device)
        if self.x.device == dev or (dev.type == "cuda" and dev.index is None and self.x.device.type == "cuda"):
            return self
        fn_x = None
        if dev.type == "cuda" and self.x.device.type == "cpu" and self.comm is not None and self.comm.world > 1:
            fn_x = lambda t: self.comm.upload_replicated(t, dev)   # replicated features: PCIe once per box
        return self._map(lambda t: t.to(dev, non_blocking=non_blocking), fn_x)

    _UPLOAD_AUX = ("_gid32", "_src_rowptr")

    def pin_memory(self):
        o = self._map(lambda t: t.pin_memory())
        for k in self._UPLOAD_AUX:
            if getattr(self, k, None) is not None:
                setattr(o, k, getattr(self, k).pin_memory())
        return o

    def compact(self):
        """Host form for `upload_async`: int32 edge_index (node ids fit 31 bits); the global edge ids as int32 when the
        global edge count fits (widened on the device); the source row of a source-sorted local edge list as its CSR
        row pointer (N + 1 ints, rebuilt on the device).  24 -> 16 bytes per local edge on the host link, and the
        device tensors are bit for bit those of `.to(device)`."""
        o = self._map_named(lambda k, t: t.to(torch.int32) if (k == "edge_index" and t.dtype != torch.int32) else t)
        ei = o.edge_index
        e, n = ei.size(1), o.x.size(0)
        if ei.device.type == "cpu" and e > 0:
            if o.num_edges_global < 2 ** 31:
                o._gid32 = o.gid.to(torch.int32)
            src = ei[0]
            if e > n + 1 and bool((src[1:] >= src[:-1]).all()) and int(src[0]) >= 0 and int(src[-1]) < n:
                nodes = torch.arange(n + 1, dtype=torch.int32)
                o._src_rowptr = torch.searchsorted(src, nodes, right=False).to(torch.int32)
        return o

    def _map_named(self, fn):
        o = ShardedBatch(None, self.comm)
        for k in self._TENSORS:
            v = getattr(self, k, None)
            setattr(o, k, fn(k, v) if v is not None else None)
        o.num_classes, o.num_edges_global, o.bounds = self.num_classes, self.num_edges_global, self.bounds
        o._sgs_has_train = self._sgs_has_train
        return o

    def upload_async(self, dev, stream):
        """loader.prefetch: H2D copies of this rank's pieces on `stream`; of the replicated features only this rank's
        1 / world row slice crosses PCIe (it lands in its slot of the full buffer, see finish_upload)."""
        from .loader import copy_fields_async
        w = self.comm.world if self.comm is not None else 1
        n = self.x.size(0)
        sliced = w > 1 and not self.comm.staged and n >= w
        g32, rp = getattr(self, "_gid32", None), getattr(self, "_src_rowptr", None)
        skip = {k for k, on in (("x", sliced), ("gid", g32 is not None), ("edge_index", rp is not None)) if on}
        fields = {k: getattr(self, k, None) for k in self._TENSORS if k not in skip}
        kw = copy_fields_async(fields, dev, stream)
        if g32 is not None:           # int32 over the link, widened by the copy itself
            gid = torch.empty(g32.numel(), dtype=torch.int64, device=dev)
            with torch.cuda.stream(stream):
                gid.copy_(g32.to(dev, non_blocking=True))
            kw["gid"] = gid
        if rp is not None:            # destination row + row pointer over the link, source row rebuilt
            e = self.edge_index.size(1)
            ei = torch.empty(2, e, dtype=torch.int32, device=dev)
            with torch.cuda.stream(stream):
                ei[1].copy_(self.edge_index[1], non_blocking=True)
                rp_d = rp.to(dev, non_blocking=True)
                deg = (rp_d[1:] - rp_d[:-1]).to(torch.int64)
                ei[0].copy_(torch.repeat_interleave(torch.arange(n, dtype=torch.int32, device=dev), deg,
                                                    output_size=e))
            kw["edge_index"] = ei
        o = self._map_named(lambda k, t: kw.get(k))
        o._x_pending = None
        if sliced:
            chunk = (n + w - 1) // w
            lo, hi = min(self.comm.rank * chunk, n), min((self.comm.rank + 1) * chunk, n)
            full = torch.empty((w * chunk,) + tuple(self.x.shape[1:]), dtype=self.x.dtype, device=dev)
            with torch.cuda.stream(stream):
                full[self.comm.rank * chunk:self.comm.rank * chunk + (hi - lo)].copy_(self.x[lo:hi], non_blocking=True)
            o._x_pending = (full, chunk, n)
        return o

    def finish_upload(self):
        """On the consumer's stream, after the copies have landed: the NVLink all-gather of the feature slices."""
        pend = getattr(self, "_x_pending", None)
        if pend is not None:
            full, chunk, n = pend
            mine = full[self.comm.rank * chunk:(self.comm.rank + 1) * chunk]
            with _timed("comm_exchange"):
                dist.all_gather_into_tensor(full, mine.clone(), group=self.comm.group)
            self.x = full[:n]
            self._x_pending = None
        return self

    def nbytes(self):
        return sum(getattr(self, k).numel() * getattr(self, k).element_size()
                   for k in self._TENSORS if getattr(self, k, None) is not None)

    def upload_nbytes(self):
        """Bytes this rank sends over the host link in `upload_async`: everything but the other ranks' slices of x,
        with the compact forms of `gid` / the source row when `compact()` prepared them."""
        w = self.comm.world if self.comm is not None else 1
        nb = self.nbytes()
        if getattr(self, "_gid32", None) is not None:
            nb -= self.gid.numel() * 4
        if getattr(self, "_src_rowptr", None) is not None:
            nb -= self.edge_index.size(1) * self.edge_index.element_size() - self._src_rowptr.numel() * 4
        if w > 1 and not self.comm.staged and self.x.size(0) >= w:
            chunk = (self.x.size(0) + w - 1) // w
            rows = max(0, min((self.comm.rank + 1) * chunk, self.x.size(0)) - self.comm.rank * chunk)
            nb -= (self.x.size(0) - rows) * self.x[0].numel() * self.x.element_size()
        return nb


# ------------------------------------------------------------------------------------------
# GCNConv on a destination-sharded edge list
# ------------------------------------------------------------------------------------------

class ShardedGCNConvFn(torch.autograd.Function):
    """out = act(A_hat (x W^T) + b) for the OWNED rows (other rows of the returned [N,D] tensor are
    unspecified unless exchange_out).  x_full: x is replicated (valid on every row) and the dense
    layer is computed redundantly; otherwise x is valid on the owned rows only, the dense layer runs
    on the slab and its output is all-gathered before the SpMM.  Incoming gradients need to be valid
    on the owned rows only (exchange_out: partial sums on all rows, reduced here)."""

    @staticmethod
    def forward(ctx, x, weight, bias, edge_weight, lg, relu, p_drop, seed, x_full, exchange_out):
        x = _req(x, torch.float32, "x")
        weight = _req(weight, torch.float32, "weight")
        bias = _req(bias, torch.float32, "bias")
        g = lg.graph
        n, lo, hi = g.num_nodes, lg.lo, lg.hi
        if x.size(0) != n:
            raise RuntimeError("x must have one row per node")
        if x_full and x.requires_grad:
            raise RuntimeError("a replicated dense input cannot require grad in sharded mode")
        norm = lg.norm(edge_weight)
        d, fin = weight.size(0), x.size(1)
        if x_full:
            h = ops.linear_nt(x, weight, static_x=True)
        else:
            h = torch.empty(n, d, dtype=torch.float32, device=x.device)
            if hi > lo:
                ops.linear_nt_into(x[lo:hi], weight, h[lo:hi])
            lg.comm.exchange_rows(h, lg.bounds)
        stage = lg.comm.peer_stage(n, d) if (exchange_out and lg.comm.world > 1) else None
        tab = ops.gather_table(h)
        out = ops.spmm(g.csr_dst, norm.what_dst, norm, h, bias, relu, p_drop, seed, table=tab,
                       rows=(lo, hi), peers=(lg.comm, stage) if stage is not None else None)
        if stage is not None:
            lg.comm.peer_finish_exchange(out, lg.bounds, stage, pushed=True)
        elif exchange_out:
            lg.comm.exchange_rows(out, lg.bounds)
        ctx.lg, ctx.norm, ctx.relu, ctx.p_drop, ctx.exchange_out = lg, norm, relu, p_drop, exchange_out
        ctx.has_w = edge_weight is not None
        ctx.tab = tab if ctx.has_w else None       # the edge-weight gradient gathers the same rows of h
        ctx.save_for_backward(x, weight, h if ctx.has_w else None, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, weight, h, out = ctx.saved_tensors
        lg, norm = ctx.lg, ctx.norm
        g = lg.graph
        comm, bounds, lo, hi = lg.comm, lg.bounds, lg.lo, lg.hi
        n, d = gout.shape
        ns = hi - lo
        fin = x.size(1)
        gout = _req(gout, torch.float32, "grad")
        g_slab = comm.reduce_rows(gout, bounds) if ctx.exchange_out else gout[lo:hi]
        g_full = torch.zeros(n, d, dtype=torch.float32, device=gout.device)
        if ns > 0:
            if ctx.relu:
                scale = 1.0 / (1.0 - ctx.p_drop) if ctx.p_drop > 0 else 1.0
                check(lib().sgs_act_bwd(_p(g_slab), _p(out[lo:hi]), ns * d, scale, _p(g_full[lo:hi]), _stream()),
                      "sgs_act_bwd")
            else:
                g_full[lo:hi].copy_(g_slab)
        need_x, need_w, need_b, need_ew = ctx.needs_input_grad[:4]
        db = dw = dx = dew = None
        if need_b:
            db = torch.zeros(d, dtype=torch.float32, device=gout.device)
            if ns > 0:
                check(lib().sgs_colsum(_p(g_full[lo:hi]), ns, d, _p(db), _stream()), "sgs_colsum")
        if need_w or need_x:
            tab_g = ops.gather_table(g_full, scaled=True)
            if comm.peer is not None and comm.world > 1:
                # partial sums for arbitrary rows, left in the symmetric arena and summed with peer loads
                part, off = comm.peer.region("part", n, d)
                ops.spmm(g.csr_src, norm.what_src, norm, g_full, table=tab_g, out=part)
                dh_slab = comm.peer_reduce_rows(part, off, bounds)
            else:
                dh = ops.spmm(g.csr_src, norm.what_src, norm, g_full, table=tab_g)
                dh_slab = comm.reduce_rows(dh, bounds)
            if need_w:
                dw = torch.zeros_like(weight)
                if ns > 0:
                    xs = x
                    if fin % 4 and not x.requires_grad:      # static features: the padded copy is 16-byte aligned
                        xs = ops._rows_aligned16(x, cache=True)[0]
                    ops.gemm_tn(dh_slab, xs[lo:hi], out=dw) if xs is x else \
                        ops.gemm(dh_slab, 1, d, xs[lo:hi], 1, xs.size(1), d, fin, ns, out=dw,
                                 precision=ops._state["gemm"] if (d % 4 == 0 and ops._state["gemm"] != PREC_FP32)
                                 else PREC_FP32)
            if need_x:
                dx = torch.zeros(n, fin, dtype=torch.float32, device=gout.device)
                if ns > 0:
                    if ops._state["gemm"] != PREC_FP32:   # dh W = dh (W^T)^T: NT form on tcgen05
                        ops.linear_nt_into(dh_slab, weight.t().contiguous(), dx[lo:hi])
                    else:
                        ops.gemm(dh_slab, d, 1, weight, 1, fin, ns, fin, d, out=dx[lo:hi], precision=PREC_FP32)
        if need_ew and ctx.has_w:
            m = g.num_edges
            dew = ops._vec(m, torch.float32, gout.device, zero=True)
            tmp = ops._ws(4 * (2 * m + n), gout.device, "edge_grad_tmp").view(torch.float32)
            tmp.zero_()
            rp_d, pm_d, nb_d, od_d = g.csr_dst
            rp_s, pm_s, _, _ = g.csr_src
            with _timed(f"edge_grad_d{d}"):
                if m > 0 and ctx.tab is not None:
                    check(lib().sgs_gcn_edge_grad_partial_h16(_p(rp_d), _p(pm_d), _p(nb_d), _p(norm.what_dst),
                                                              _p(od_d), _p(rp_s), _p(pm_s), _p(g_full),
                                                              _p(ctx.tab.data), _p(ctx.tab.scale), _p(norm.dis),
                                                              _p(norm.loopw), m, n, d, _p(tmp), _p(tmp[m:]),
                                                              _p(tmp[2 * m:]), _stream()),
                          "sgs_gcn_edge_grad_partial_h16")
                elif m > 0:
                    check(lib().sgs_gcn_edge_grad_partial(_p(rp_d), _p(pm_d), _p(nb_d), _p(norm.what_dst), _p(od_d),
                                                          _p(rp_s), _p(pm_s), _p(g_full), _p(h), _p(norm.dis),
                                                          _p(norm.loopw), m, n, d, _p(tmp), _p(tmp[m:]),
                                                          _p(tmp[2 * m:]), _stream()), "sgs_gcn_edge_grad_partial")
            ctx.tab = None
            comm.all_reduce(tmp[2 * m:])
            if m > 0:
                check(lib().sgs_gcn_edge_grad_final(_p(g.src), _p(g.dst), _p(tmp), _p(tmp[2 * m:]), _p(norm.dis),
                                                    _p(norm.deg), m, _p(dew), 0, _stream()),
                      "sgs_gcn_edge_grad_final")
        return dx, dw, db, dew, None, None, None, None, None, None


def sconv(conv, x, lg, edge_weight=None, relu=False, p_drop=0.0, seed=0, x_full=False, exchange_out=False):
    return ShardedGCNConvFn.apply(x, conv.lin.weight, conv.bias, edge_weight, lg, relu, float(p_drop), int(seed),
                                  bool(x_full), bool(exchange_out))


def embed(scorer, x, lg):
    """EdgeProbGCN pre-stage (model.py:106-111) -> `out` [N,H] valid on every rank."""
    if not hasattr(scorer, "gcn1"):
        # EdgeProbMLP: per-node relu(fcdim(x)) on replicated x; gradients are partial sums already
        return scorer.embed(x, lg.graph)
    h = sconv(scorer.gcn1, x, lg, None, True, scorer._drop(), ops.next_seed(), x_full=True)
    return sconv(scorer.gcn2, h, lg, None, True, 0.0, 0, exchange_out=True)


def gnn_forward(model, x, lg, edge_weight=None):
    """GNNModel.forward (model.py:155-164) -> logits [N,C] valid on every rank."""
    p_drop = float(model.dropout.p) if model.training else 0.0
    h = sconv(model.gcn1, x, lg, edge_weight, True, p_drop, ops.next_seed(), x_full=True)
    return sconv(model.gcn2, h, lg, edge_weight, False, 0.0, 0, exchange_out=True)


# ------------------------------------------------------------------------------------------
# straight-through weights with the global normaliser terms
# ------------------------------------------------------------------------------------------

class ShardedStraightThroughFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p_full, prob, sel, S, mode, coef, comm):
        p_full = _req(p_full, torch.float32, "edge_probs")
        p_sel, w = ops.gather_selected(p_full, prob, sel, mode, coef, S, straight_through=True)
        ctx.sel, ctx.mode, ctx.coef, ctx.e, ctx.comm = sel, mode, coef, p_full.numel(), comm
        ctx.save_for_backward(p_sel, w, S)
        return w

    @staticmethod
    def backward(ctx, g):
        import numpy as np
        p_sel, w, S = ctx.saved_tensors
        a = float(np.float32(1.0 - ctx.coef)) if ctx.mode == SAMPLE_TRAIN else 1.0
        s_eff = S + 1e-12
        gamma = torch.where((w > 0) & (w < 1), g, torch.zeros_like(g))
        tot = (gamma * p_sel * p_sel).sum().reshape(1)
        ctx.comm.all_reduce(tot)                       # the constant term sums over ALL selected edges
        const = (a / (s_eff * s_eff)) * tot
        sel_term = gamma * (1.0 + a * p_sel / s_eff)
        out = (-const).expand(ctx.e).contiguous()
        check(lib().sgs_scatter_selected(_p(sel_term.contiguous()), _p(ctx.sel), sel_term.numel(), _p(out),
                                         _stream()), "sgs_scatter_selected")
        return out, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------
# the learned step
# ------------------------------------------------------------------------------------------

def _check(flag_bad, n_sel, q):
    if flag_bad:
        raise RuntimeError("probability tensor contains either `inf`, `nan` or element < 0")
    if n_sel != q:
        raise RuntimeError(f"sampler selected {n_sel} edges, expected {q}")


def learned_step(pipeline, args, epoch, max_epoch, model, sb, criterion, q, backward_fn):
    """training_hybrid.py:39-147 / training_straight_through.py:36-134 on this rank's shard.
    `q` is the GLOBAL edge budget.  Returns (loss tensor identical on all ranks, update_edge_mlp);
    afterwards every parameter's .grad holds this rank's partial sum (see allreduce_partial_grads)."""
    comm = sb.comm
    lg = sb.local
    g_loc = lg.graph
    dev = sb.x.device
    scorer = model.edge_prob_mlp
    coef = args.degree_bias_coef
    topq = sdist.DistributedTopQ(sdist.CudaTopQOps(persistent=True), group=comm.group)
    e_loc = g_loc.num_edges
    tm_owned = sb.train_mask_owned.view(torch.uint8)    # CE / accuracy rows: train AND owned
    tm_full = sb.train_mask.view(torch.uint8)           # reg1 validity tests both endpoints on the full mask

    def draw_noise():
        inj = sampling._next_noise()
        if inj is not None:
            return inj[sb.gid].contiguous()
        # keyed by GLOBAL edge id with a seed every rank derives identically: the keys -- and with them the sampled
        # set -- are those of the unsharded draw for any number of ranks (SURVEY 8e's invariance claim)
        return ops.exponential(e_loc, dev, seed=ops.next_seed(), gid=sb.gid)

    lg_rand = None
    r = None
    if args.conditional or args.sparse_edge_mlp:
        r = topq.select_ex(sb.scores, None, draw_noise(), q, SAMPLE_RAW, 0.0, gid=sb.gid)
        _check(r.invalid, r.n_global, q)
        lg_rand = lg.subgraph(r.sel)

    out = embed(scorer, sb.x, lg_rand if lg_rand is not None else lg)
    seed_sc = ops.next_seed()
    p_drop = scorer._drop()
    fc1, fc2 = scorer.fc1, scorer.fc2
    with torch.no_grad():
        p_loc = ops.edge_score_forward(out.detach(), g_loc, fc1.weight, fc1.bias, fc2.weight.reshape(-1),
                                       fc2.bias.reshape(-1), None, p_drop, seed_sc)

    smp = topq.select_ex(p_loc, sb.prob, draw_noise(), q, SAMPLE_TRAIN, coef, S=sampling._next_S(), gid=sb.gid)
    _check(smp.invalid, smp.n_global, q)
    lg_s = lg.subgraph(smp.sel)
    if pipeline == "hybrid":
        p_sel = ops.gather_selected(p_loc, None, smp.sel, SAMPLE_RAW, 0.0, None)[0]
        p_s = scorer.score(out, g_loc, ids=smp.sel, precomputed=p_sel, seed=seed_sc)
    elif pipeline == "straight_through":
        p_full_g = scorer.score(out, g_loc, precomputed=p_loc, seed=seed_sc)
        p_s = ShardedStraightThroughFn.apply(p_full_g, sb.prob, smp.sel, smp.S, SAMPLE_TRAIN, coef, comm)
    else:
        raise ValueError(pipeline)

    learned_out = gnn_forward(model, sb.x, lg_s, p_s)

    update_edge_mlp = True
    with_edges = bool(args.reg1 or args.reg2)
    acc_l = ops.loss_forward(learned_out.detach(), sb.y, tm_full, lg_s.graph if with_edges else None,
                             p_s.detach() if with_edges else None, row_mask_u8=tm_owned)
    est_l = ops.edge_state_of(acc_l)   # acc_l is re-sliced from the all-reduced sums below: keep its edge state
    acc_r = random_out = None
    if args.conditional:
        random_out = gnn_forward(model, sb.x, lg_rand)
        acc_r = ops.loss_forward(random_out.detach(), sb.y, tm_full, row_mask_u8=tm_owned)
        acc = torch.cat([acc_l, acc_r])
    else:
        acc = torch.cat([acc_l, torch.zeros_like(acc_l)])
    comm.all_reduce(acc)
    acc_l, acc_r = acc[:8].contiguous(), acc[8:].contiguous()
    host = torch.cat([acc[2:3], acc[10:11]]).cpu()
    if args.conditional:
        update_edge_mlp = bool(host[0] > host[1])
        if getattr(args, "force_branch", None) == "learned":   # bench only: always time the full (learned-wins) step
            update_edge_mlp = True
        elif getattr(args, "force_branch", None) == "random":  # tests only: a random-wins step on demand
            update_edge_mlp = False
    if update_edge_mlp:
        loss = ops.fused_loss(learned_out, sb.y, tm_full, p_s if with_edges else None,
                              lg_s.graph if with_edges else None, args.regularizer1_coef, args.consist_reg_coef,
                              bool(args.reg1), bool(args.reg2), acc=acc_l, row_mask_u8=tm_owned, edge_state=est_l)
    else:
        loss = ops.fused_loss(random_out, sb.y, tm_full, acc=acc_r, reg1=False, reg2=False, row_mask_u8=tm_owned)
    backward_fn(loss)
    return loss, update_edge_mlp


def allreduce_partial_grads(params, comm):
    """Every rank holds partial sums (over its rows / edges) of the weight gradients: one flat all-reduce SUM.
    A parameter whose gradient is None on EVERY rank (e.g. edge_prob_mlp.gcn* on a random-wins step: they sit in
    optimizer_gnn through main.py:100's name filter but are outside that branch's autograd graph) keeps grad = None,
    so Adam skips it exactly as on one GPU / in the reference; the has-grad bitmap rides in the same buffer."""
    if comm.world == 1:
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
    comm.all_reduce(flat)
    # the bitmap is only read (one small D2H) when some local gradient is None; the usual learned-wins step has none
    seen = flat[-len(ps):].cpu() if any(p.grad is None for p in ps) else None
    off = 0
    for i, p in enumerate(ps):
        k = p.numel()
        if p.grad is not None:
            p.grad.copy_(flat[off:off + k].view_as(p))
        elif float(seen[i]) > 0.0:
            p.grad = flat[off:off + k].view_as(p).clone()
        off += k
