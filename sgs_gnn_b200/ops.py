"""Thin PyTorch-facing wrappers over the C ABI (include/sgs_b200.h): tensor validation,
workspace allocation with torch's caching allocator, and the autograd.Functions that give the
reference's modules their backward.  PyTorch is plumbing here (device memory, streams,
autograd bookkeeping); every arithmetic step is a libsgs_b200 kernel.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import itertools
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import (PREC_BF16, PREC_FP16, PREC_FP32, PREC_TF32, SAMPLE_RAW, SAMPLE_TEST, SAMPLE_TRAIN,
                   SPMM_ACCUM, SPMM_ADD_ROOT, SPMM_DROPOUT, SPMM_RELU, check, lib)

# ------------------------------------------------------------------------------------------
# configuration
# ------------------------------------------------------------------------------------------
_PRECISION = {"fp32": PREC_FP32, "bf16": PREC_BF16, "fp16": PREC_FP16, "tf32": PREC_TF32}
_state = {"gemm": PREC_FP32, "scorer": PREC_FP32, "gather": PREC_FP32}


def set_precision(gemm=None, scorer=None, gather=None):
    """Precision of the dense contractions: 'fp32' (CUDA cores, parity mode) or a tcgen05
    tensor-core mode ('bf16' / 'fp16' / 'tf32').
    gather: storage of the rows the wide (D >= 64) SpMM / SDDMM gather -- 'fp32' (the [N,D] tensor as it is) or
    'fp16' (a 16-bit copy with a power-of-two scale that stays L2-resident; fp32 accumulation)."""
    if gemm is not None:
        _state["gemm"] = _PRECISION[gemm] if isinstance(gemm, str) else int(gemm)
    if scorer is not None:
        _state["scorer"] = _PRECISION[scorer] if isinstance(scorer, str) else int(scorer)
    if gather is not None:
        g = _PRECISION[gather] if isinstance(gather, str) else int(gather)
        if g not in (PREC_FP32, PREC_FP16):
            raise ValueError("gather precision must be 'fp32' or 'fp16'")
        _state["gather"] = g


def scorer_supports(precision):
    """Edge-scorer precision modes of this build: 'fp32' (CUDA cores), 'tf32' (tcgen05 kind::tf32 GEMMs on the chunked
    fp32 pipeline: the tensor-core mode that keeps the fp32 1e-4 parity bar), 'fp16' / 'bf16' (fused tcgen05 kernels)."""
    return precision in _PRECISION


def get_precision():
    inv = {v: k for k, v in _PRECISION.items()}
    return {k: inv[v] for k, v in _state.items()}


_seed_counter = itertools.count(1)


def next_seed():
    """Counter-based seed for dropout masks / sampler noise, derived from torch's seed so that
    utils.fix_seeds() makes runs reproducible."""
    base = torch.initial_seed() & 0xFFFFFFFFFFFF
    return (base * 0x9E3779B1 + next(_seed_counter) * 0x85EBCA77) & 0xFFFFFFFFFFFFFFFF


def reset_seed_counter():
    global _seed_counter
    _seed_counter = itertools.count(1)


# ------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------

def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream():
    """The current CUDA stream's handle.  ~200 C-ABI calls per step ask for it: the raw getter skips the
    torch.cuda.Stream object the public call builds (several microseconds each on a host that is launch-bound at
    small shapes and at 8 GPUs)."""
    if _raw_stream is not None and _raw_device is not None:
        return C.c_void_p(_raw_stream(_raw_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t, dtype, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (sgs_gnn_b200 has no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


_scratch = {}


def _ws(nbytes, device, name=None):
    """Workspace for one C-ABI call.  With `name`, a persistent per-purpose arena that only ever grows (x1.25):
    the multi-GB scratch buffers of the scorer / CSR build / sampler then never go back through the caching
    allocator, whose best-fit reuse breaks down (cudaMalloc + cudaFree stalls inside the step) as soon as the
    request sizes wobble from step to step, e.g. with the per-rank edge counts of a sharded graph."""
    nbytes = max(int(nbytes), 256)
    if name is None:
        return torch.empty(nbytes, dtype=torch.uint8, device=device)
    key = (name, str(device))
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _scratch.pop(key, None)
        buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf[:nbytes]


_VEC_QUANTUM = 1 << 20


def _vec(n, dtype, device, zero=False):
    """1-D result buffer of n elements whose allocation is rounded up to a 1 Mi-element quantum, so that vectors
    whose length follows a sampled edge count hit the same cached blocks every step."""
    n = int(n)
    cap = max(n, 1)
    if cap > _VEC_QUANTUM:
        cap = (cap + _VEC_QUANTUM - 1) // _VEC_QUANTUM * _VEC_QUANTUM
    buf = (torch.zeros if zero else torch.empty)(cap, dtype=dtype, device=device)
    return buf[:n]


# optional per-kernel-family CUDA-event timing (bench.py's roofline leg); off by default
_timer = None


class KernelTimer:
    """Collects (start, end) CUDA events on the launching stream around each wrapped C-ABI call."""

    def __init__(self):
        self.events = {}

    def __enter__(self):
        global _timer
        _timer = self
        return self

    def __exit__(self, *exc):
        global _timer
        _timer = None

    def totals_ms(self):
        torch.cuda.synchronize()
        return {k: (sum(a.elapsed_time(b) for a, b in v), len(v)) for k, v in self.events.items()}


_nvtx = bool(__import__("os").environ.get("SGS_NVTX"))


def enable_nvtx(on=True):
    """NVTX ranges around every kernel family (the names bench.py's kernel_time_share uses) and around the
    reference's profiler segments (utils.GpuMemoryProfiler: edge_mlp_pre, edge_score, gnn_forward, backward), so an
    nsys / ncu timeline is readable.  Off by default (SGS_NVTX=1 turns it on at import)."""
    global _nvtx
    _nvtx = bool(on)


def nvtx_enabled():
    return _nvtx


def seg_begin(profiler, name):
    """Start of one of the reference's profiler segments (utils.py:13-80 names): NVTX range + memory profiler."""
    if _nvtx:
        torch.cuda.nvtx.range_push(name)
    if profiler is not None:
        profiler.begin(name)


def seg_end(profiler, name):
    if profiler is not None:
        profiler.end(name)
    if _nvtx:
        torch.cuda.nvtx.range_pop()


class _timed:
    __slots__ = ("name", "a")

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _nvtx:
            torch.cuda.nvtx.range_push(self.name)
        if _timer is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if _nvtx:
            torch.cuda.nvtx.range_pop()
        if _timer is not None:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            _timer.events.setdefault(self.name, []).append((self.a, b))


# ------------------------------------------------------------------------------------------
# graph structure
# ------------------------------------------------------------------------------------------

class GcnNorm:
    """deg / dis / loop weights and the normalised edge weights in both CSR orders."""

    __slots__ = ("deg", "dis", "loopw", "what_dst", "_what_src", "_graph", "_w")

    def __init__(self, graph, w):
        dev = graph.device
        n, m = graph.num_nodes, graph.num_edges
        rowptr, perm, nbr, _ = graph.csr_dst
        self.deg = torch.empty(n, dtype=torch.float32, device=dev)
        self.dis = torch.empty_like(self.deg)
        self.loopw = torch.empty_like(self.deg)
        self.what_dst = _vec(max(m, 1), torch.float32, dev)
        check(lib().sgs_gcn_norm(_p(rowptr), _p(perm), _p(nbr), _p(w), m, n, _p(self.deg), _p(self.dis),
                                 _p(self.loopw), _p(self.what_dst), _stream()), "sgs_gcn_norm")
        self._what_src = None
        self._graph = weakref.ref(graph)   # weak: graph -> norm -> graph would only be freed by the cyclic GC
        self._w = w

    @property
    def what_src(self):
        if self._what_src is None:
            g = self._graph()
            if g is None:
                raise RuntimeError("the graph of this gcn_norm has been released")
            rowptr, perm, nbr, _ = g.csr_src
            self._what_src = _vec(max(g.num_edges, 1), torch.float32, g.device)
            check(lib().sgs_gcn_norm_apply(_p(rowptr), _p(perm), _p(nbr), _p(self._w), _p(self.dis),
                                           g.num_edges, g.num_nodes, _p(self._what_src), _stream()),
                  "sgs_gcn_norm_apply")
        return self._what_src


class Graph:
    """An edge list narrowed to int32 with lazily built CSR-by-destination (forward SpMM) and
    CSR-by-source (backward SpMM) views, and cached gcn_norm results."""

    def __init__(self, src, dst, num_nodes, edge_index=None):
        self.src = src
        self.dst = dst
        self.num_nodes = int(num_nodes)
        self.num_edges = int(src.numel())
        self.device = src.device
        self.edge_index = edge_index
        self._csr_dst = None
        self._csr_src = None
        self._src_sorted = None   # is `src` non-decreasing?  (None: not checked yet)
        self._mean_w = None       # SAGEConv mean-aggregation weights in both CSR orders (lazily built)
        self._norm_unw = None
        self._norm_w = None  # (weakref to weight tensor, version, GcnNorm)
        self.oob_flag = None      # device int32 [1] set by sgs_edge_index_split when an id lies outside [0, N)
        self._keep = None         # owner of src / dst when they are views (int32 edge_index given as such)

    def take_oob_flag(self):
        """The not-yet-checked out-of-range flag of this edge list (device int32 [1]) or None; the caller folds it
        into a host read it performs anyway (end of the training step) and raises like PyG's device assert would."""
        f, self.oob_flag = self.oob_flag, None
        return f

    @staticmethod
    def from_edge_index(edge_index, num_nodes, validate=False):
        """edge_index: the reference's int64 [2, E] tensor, or the int32 [2, E] form of a compacted batch (its rows are
        used as they are: no narrowing pass, half the bytes)."""
        if isinstance(edge_index, torch.Tensor) and edge_index.dtype == torch.int32:
            ei = _req(edge_index, torch.int32, "edge_index")
            if ei.dim() != 2 or ei.size(0) != 2:
                raise RuntimeError("edge_index must have shape [2, E]")
            m = ei.size(1)
            src, dst = ei[0], ei[1]
            flag = torch.zeros(1, dtype=torch.int32, device=ei.device)
            check(lib().sgs_edge_index_check32(_p(src), _p(dst), m, int(num_nodes), _p(flag), _stream()),
                  "sgs_edge_index_check32")
            if validate and int(flag.item()) != 0:
                raise RuntimeError("edge_index contains node ids outside [0, num_nodes)")
            g = Graph(src, dst, num_nodes, None)
            g._keep = ei          # the rows are views of this tensor
            if not validate:
                g.oob_flag = flag
            return g
        ei = _req(edge_index, torch.int64, "edge_index")
        if ei.dim() != 2 or ei.size(0) != 2:
            raise RuntimeError("edge_index must have shape [2, E]")
        m = ei.size(1)
        src = torch.empty(m, dtype=torch.int32, device=ei.device)
        dst = torch.empty_like(src)
        flag = torch.zeros(1, dtype=torch.int32, device=ei.device)
        check(lib().sgs_edge_index_split(_p(ei), m, int(num_nodes), _p(src), _p(dst), _p(flag), _stream()),
              "sgs_edge_index_split")
        if validate and int(flag.item()) != 0:
            raise RuntimeError("edge_index contains node ids outside [0, num_nodes)")
        g = Graph(src, dst, num_nodes, ei)
        if not validate:
            g.oob_flag = flag
        return g

    @property
    def src_sorted(self):
        """True when the source column is non-decreasing (PyG edge lists are (src,dst)-sorted): checked on the
        device once per graph (one 4-byte read)."""
        if self._src_sorted is None:
            flag = torch.empty(1, dtype=torch.int32, device=self.device)
            check(lib().sgs_keys_unsorted(_p(self.src), self.num_edges, _p(flag), _stream()), "sgs_keys_unsorted")
            self._src_sorted = int(flag.item()) == 0
        return self._src_sorted

    def subgraph(self, ids, want_edge_index=False, ascending=False):
        """Edge-induced subgraph on edge ids (int32 [q]); optionally also the int64 [2,q] tensor.  `ascending`: the
        caller vouches that ids ascend (the sampler's compaction order) -- the subgraph of a source-sorted graph is
        then source-sorted too and its by-source CSR needs no sort."""
        q = int(ids.numel())
        src = _vec(q, torch.int32, self.device)
        dst = _vec(q, torch.int32, self.device)
        out = torch.empty(2, q, dtype=torch.int64, device=self.device) if want_edge_index else None
        if self.edge_index is not None:
            check(lib().sgs_edge_index_gather(_p(self.edge_index), self.num_edges, _p(ids), q, _p(out), _p(src),
                                              _p(dst), _stream()), "sgs_edge_index_gather")
        else:
            check(lib().sgs_edge_gather32(_p(self.src), _p(self.dst), _p(ids), q, _p(out), _p(src), _p(dst),
                                          _stream()), "sgs_edge_gather32")
        g = Graph(src, dst, self.num_nodes, out)
        if ascending and self.src_sorted:
            g._src_sorted = True
        return g

    def _build(self, key, other, presorted=False):
        n, m = self.num_nodes, self.num_edges
        rowptr = torch.empty(n + 1, dtype=torch.int32, device=self.device)
        perm = _vec(max(m, 1), torch.int32, self.device)
        nbr = _vec(max(m, 1), torch.int32, self.device)
        order = torch.empty(n + 1, dtype=torch.int32, device=self.device)   # [N] = number of hub rows
        nbytes = lib().sgs_csr_workspace_bytes(m, n)
        ws = _ws(nbytes, self.device, "csr_build")
        fn = lib().sgs_csr_build_sorted if presorted else lib().sgs_csr_build
        with _timed("csr_build"):
            check(fn(_p(key), _p(other), m, n, _p(rowptr), _p(perm), _p(nbr), _p(order), _p(ws), ws.numel(),
                     _stream()), "sgs_csr_build")
        return rowptr, perm, nbr, order

    @property
    def csr_dst(self):
        if self._csr_dst is None:
            self._csr_dst = self._build(self.dst, self.src)
        return self._csr_dst

    @property
    def csr_src(self):
        if self._csr_src is None:
            # only a hint that was already established is used here (no device read on this path)
            self._csr_src = self._build(self.src, self.dst, presorted=self._src_sorted is True)
        return self._csr_src

    def norm(self, edge_weight=None):
        if edge_weight is None:
            if self._norm_unw is None:
                self._norm_unw = GcnNorm(self, None)
            return self._norm_unw
        w = edge_weight.detach()
        c = self._norm_w
        if c is not None and c[0]() is edge_weight and c[1] == edge_weight._version:
            return c[2]
        w = _req(w, torch.float32, "edge_weight")
        if w.numel() != self.num_edges:
            raise RuntimeError("edge_weight must have one entry per edge")
        nrm = GcnNorm(self, w)
        self._norm_w = (weakref.ref(edge_weight), edge_weight._version, nrm)
        return nrm


_graph_cache = {}


def graph_of(edge_index, num_nodes):
    """Graph for an int64 [2,E] tensor, cached on the tensor's identity + version (the way PyG
    caches normalisation), evicted when the tensor dies."""
    if isinstance(edge_index, Graph):
        return edge_index
    key = id(edge_index)
    hit = _graph_cache.get(key)
    if hit is not None and hit[0]() is edge_index and hit[1] == edge_index._version and hit[2].num_nodes == num_nodes:
        return hit[2]
    g = Graph.from_edge_index(edge_index, num_nodes)
    try:
        ref = weakref.ref(edge_index, lambda _r, k=key, c=_graph_cache: c.pop(k, None))
        _graph_cache[key] = (ref, edge_index._version, g)
    except TypeError:
        pass
    return g


# ------------------------------------------------------------------------------------------
# dense contraction
# ------------------------------------------------------------------------------------------

def gemm(a, a_sm, a_sk, b, b_sn, b_sk, m, n, k, out=None, accumulate=False, precision=None):
    """out[m,n] (+)= sum_k A(m,k) B(n,k) with explicit element strides (include/sgs_b200.h K4)."""
    prec = _state["gemm"] if precision is None else precision
    if out is None:
        out = torch.empty(m, n, dtype=torch.float32, device=a.device)
    with _timed("gemm"):
        check(lib().sgs_gemm(_p(a), a_sm, a_sk, _p(b), b_sn, b_sk, _p(out), out.stride(0), m, n, k,
                             1 if accumulate else 0, prec, _stream()), "sgs_gemm")
    return out


_pad_cache = {}


def round_tf32(t, out=None):
    """t rounded to the nearest tf32 value (sgs_round_tf32): unbiased operands for the truncating kind::tf32 MMA."""
    t = _req(t, torch.float32, "tensor")
    if out is None:
        out = torch.empty_like(t)
    check(lib().sgs_round_tf32(_p(t), t.numel(), _p(out), _stream()), "sgs_round_tf32")
    return out


def _tf32_operand(t, static=False):
    """Operand of a kind::tf32 GEMM: (tensor, ld) with a 16-byte aligned row stride (TMA; odd row lengths such as
    F = 602 are padded) and values rounded to the nearest tf32.  Static inputs (node features) are prepared once."""
    k = t.size(1)
    if static:
        key = id(t)
        hit = _pad_cache.get(key)
        if hit is not None and hit[0]() is t and hit[1] == t._version:
            return hit[2], hit[2].size(1)
    if k % 4 == 0 and t.data_ptr() % 16 == 0 and t.is_contiguous():
        buf = round_tf32(t)
    else:
        ld = (k + 3) // 4 * 4
        buf = torch.zeros(t.size(0), ld, dtype=t.dtype, device=t.device)
        buf[:, :k].copy_(t)
        round_tf32(buf, out=buf)
    if static:
        try:
            _pad_cache[key] = (weakref.ref(t, lambda _r, kk=key, c=_pad_cache: c.pop(kk, None)), t._version, buf)
        except TypeError:
            pass
    return buf, buf.size(1)


def _rows_aligned16(t, cache=False):
    """(tensor, ld) with ld % 4 == 0 for TMA, tf32-rounded (see _tf32_operand)."""
    return _tf32_operand(t, static=cache)


_lin_cache = {}


def linear_nt(x, w, precision=None, static_x=False, remember=True):
    """x[M,K] @ w[N,K]^T.  static_x (node features): the product is remembered until `w` changes -- the learned and
    the random-baseline forward of one step apply the same gcn1 weight to the same features (training_hybrid.py:88,93),
    so the second projection (and its gather table) is free.  remember=False for a caller that writes into the
    result (the returned tensor is then private to it)."""
    x = _req(x, torch.float32, "x")
    w = _req(w, torch.float32, "weight")
    prec = _state["gemm"] if precision is None else precision
    m, k, n = x.size(0), x.size(1), w.size(0)
    key = None
    if static_x and remember:
        key = (id(x), id(w))
        hit = _lin_cache.get(key)
        if hit is not None and hit[0]() is x and hit[1]() is w and hit[2] == (x._version, w._version, prec):
            return hit[3]
    if prec == PREC_TF32:
        xa, lda = _tf32_operand(x, static=static_x)
        wa, ldb = _tf32_operand(w)
        out = gemm(xa, lda, 1, wa, ldb, 1, m, n, k, precision=prec)
    else:
        out = gemm(x, k, 1, w, k, 1, m, n, k, precision=prec)
    if key is not None:
        try:
            drop = lambda _r, kk=key, c=_lin_cache: c.pop(kk, None)   # noqa: E731
            _lin_cache[key] = (weakref.ref(x, drop), weakref.ref(w, drop), (x._version, w._version, prec), out)
        except TypeError:
            pass
    return out


def linear_nt_into(x, w, out, precision=None):
    """out[M,N] = x[M,K] @ w[N,K]^T into a caller-provided (row-contiguous) buffer, e.g. a row slab."""
    prec = _state["gemm"] if precision is None else precision
    m, k, n = x.size(0), x.size(1), w.size(0)
    if prec != PREC_FP32:
        xa, lda = _tf32_operand(x)
        wa, ldb = _tf32_operand(w)
        return gemm(xa, lda, 1, wa, ldb, 1, m, n, k, out=out, precision=PREC_TF32)
    return gemm(x, k, 1, w, k, 1, m, n, k, out=out, precision=prec)


def gemm_tn(a, b, out=None, static_b=False, precision=None):
    """out[M,N] = a[K,M]^T @ b[K,N] (weight gradients dW = dh^T x: a reduction over the K node rows).
    Tensor-core mode: tcgen05 kind::tf32 with MN-major operands straight from the row-major tensors, K split
    over the CTAs; needs 16-byte aligned row strides (odd widths fall back to the fp32 split-K kernel, a static
    `b` such as the node features is padded once).  Operands are rounded to the nearest tf32 first."""
    prec = _state["gemm"] if precision is None else precision
    k, m = a.shape
    n = b.size(1)
    if prec != PREC_FP32 and m % 4 == 0 and a.data_ptr() % 16 == 0:
        if n % 4 == 0 and b.data_ptr() % 16 == 0 or static_b:
            bb, ldb = _tf32_operand(b, static=static_b)
            aa = round_tf32(a if a.is_contiguous() else a.contiguous())
            return gemm(aa, 1, m, bb, 1, ldb, m, n, k, out=out, precision=PREC_TF32)
    return gemm(a, 1, m, b, 1, n, m, n, k, out=out, precision=PREC_FP32)


class GatherTable:
    """fp16 copy of an [N, D] fp32 tensor whose rows a SpMM / SDDMM gathers (sgs_table_f16): `data` int16-typed
    storage [N, D], `scale` device float[4] = {S, 1/S, scratch, -}."""
    __slots__ = ("data", "scale")

    def __init__(self, data, scale):
        self.data, self.scale = data, scale


def gather_table(h, scaled=False):
    """The fp16 gather table of `h` when the gather precision is 'fp16' and the width qualifies (D % 8 == 0,
    64 <= D <= 512: narrower tables are L2-resident in fp32 already), else None.  scaled: gradient tables get a
    power-of-two scale from max|h| (fp16 range), activations are stored as they are."""
    if _state["gather"] != PREC_FP16:
        return None
    n, d = h.shape
    if d % 8 or d < 64 or d > 512 or (n * d) % 8:
        return None
    h = _req(h, torch.float32, "h")
    memo = getattr(h, "_sgs_table", None)      # a remembered projection (linear_nt cache) keeps its table
    if memo is not None and memo[0] == (h._version, bool(scaled)):
        return memo[1]
    data = torch.empty(n, d, dtype=torch.int16, device=h.device)
    scale = torch.empty(4, dtype=torch.float32, device=h.device)
    with _timed("table_f16"):
        check(lib().sgs_table_f16(_p(h), n, d, 1 if scaled else 0, _p(data), _p(scale), _stream()), "sgs_table_f16")
    tab = GatherTable(data, scale)
    try:
        h._sgs_table = ((h._version, bool(scaled)), tab)
    except Exception:
        pass
    return tab


def spmm(csr, what, norm, h, bias=None, relu=False, p_drop=0.0, seed=0, out=None, accumulate=False, add_root=False,
         table=None, rows=None, peers=None):
    """K3b.  `table`: the GatherTable of `h` (fp16 gather mode) -- the neighbour rows and the self term then come
    from it instead of from `h`.  Sharded form: `rows` = (lo, hi) computes only these rows; `peers` = (comm, (stage
    view, elem_off)) also stores every finished row into the stage of every peer GPU (sgs_spmm_sharded)."""
    rowptr, _perm, nbr, order = csr
    n, d = h.shape
    if out is None:
        out = torch.empty(n, d, dtype=torch.float32, device=h.device)
    flags = ((SPMM_RELU if relu else 0) | (SPMM_DROPOUT if p_drop > 0 else 0) | (SPMM_ACCUM if accumulate else 0) |
             (SPMM_ADD_ROOT if add_root else 0))
    dis = _p(norm.dis) if norm is not None else None
    loopw = _p(norm.loopw) if norm is not None else None
    if rows is not None or peers is not None:
        lo, hi = rows if rows is not None else (0, n)
        bases, world, rank, off = None, 1, 0, 0
        if peers is not None:
            comm, (_st, off) = peers
            bases, world, rank = comm.peer.bases, comm.world, comm.rank
        with _timed(f"spmm_d{d}"):
            check(lib().sgs_spmm_sharded(_p(rowptr), _p(nbr), _p(what), _p(order), dis, loopw,
                                         None if table is not None else _p(h),
                                         _p(table.data) if table is not None else None,
                                         _p(table.scale) if table is not None else None, n, d, _p(bias), _p(out),
                                         flags, float(p_drop), int(seed), int(lo), int(hi), bases, world, rank,
                                         int(off), _stream()), "sgs_spmm_sharded")
        return out
    if table is not None:
        with _timed(f"spmm_d{d}"):
            check(lib().sgs_spmm_h16(_p(rowptr), _p(nbr), _p(what), _p(order), dis, loopw, _p(table.data),
                                     _p(table.scale), n, d, _p(bias), _p(out), flags, float(p_drop), int(seed),
                                     _stream()), "sgs_spmm_h16")
        return out
    with _timed(f"spmm_d{d}"):
        check(lib().sgs_spmm(_p(rowptr), _p(nbr), _p(what), _p(order), dis, loopw, _p(h), n, d, _p(bias), _p(out),
                             flags, float(p_drop), int(seed), _stream()), "sgs_spmm")
    return out


class GCNConvFn(torch.autograd.Function):
    """GCNConv (PyG 2.3.1 semantics) with optional fused ReLU + dropout epilogue:
        out = dropout(relu(A_hat (x W^T) + b)).
    Backward: SpMM over the by-source CSR, SDDMM-based edge-weight gradient (A.3), dense dW/dx."""

    @staticmethod
    def forward(ctx, x, weight, bias, edge_weight, graph, relu, p_drop, seed):
        x = _req(x, torch.float32, "x")
        weight = _req(weight, torch.float32, "weight")
        bias = _req(bias, torch.float32, "bias")
        if x.size(0) != graph.num_nodes:
            raise RuntimeError("x must have one row per node")
        norm = graph.norm(edge_weight)
        h = linear_nt(x, weight, static_x=not x.requires_grad)
        tab = gather_table(h)
        out = spmm(graph.csr_dst, norm.what_dst, norm, h, bias, relu, p_drop, seed, table=tab)
        ctx.graph, ctx.norm, ctx.relu, ctx.p_drop = graph, norm, relu, p_drop
        ctx.has_w = edge_weight is not None
        ctx.tab = tab if ctx.has_w else None      # the edge-weight gradient gathers the same rows again
        ctx.save_for_backward(x, weight, h if (ctx.has_w and tab is None) else None, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, weight, h, out = ctx.saved_tensors
        graph, norm = ctx.graph, ctx.norm
        n, d = gout.shape
        gout = _req(gout, torch.float32, "grad")
        if ctx.relu:
            g = torch.empty_like(gout)
            scale = 1.0 / (1.0 - ctx.p_drop) if ctx.p_drop > 0 else 1.0
            check(lib().sgs_act_bwd(_p(gout), _p(out), gout.numel(), scale, _p(g), _stream()), "sgs_act_bwd")
        else:
            g = gout
        need_x, need_w, need_b, need_ew = ctx.needs_input_grad[:4]
        db = dw = dx = dew = None
        if need_b:
            db = torch.empty(d, dtype=torch.float32, device=g.device)
            check(lib().sgs_colsum(_p(g), n, d, _p(db), _stream()), "sgs_colsum")
        if need_w or need_x:
            dh = spmm(graph.csr_src, norm.what_src, norm, g, table=gather_table(g, scaled=True))
            fin = x.size(1)
            if need_w:  # dW[d, fin] = dh^T x
                dw = gemm_tn(dh, x, static_b=not x.requires_grad)
            if need_x:  # dx[n, fin] = dh W
                if _state["gemm"] != PREC_FP32:
                    # NT form on tcgen05 against the (tiny) transposed weight: dh W = dh (W^T)^T
                    dx = linear_nt(dh, weight.t().contiguous(), precision=PREC_TF32)
                else:
                    dx = gemm(dh, d, 1, weight, 1, fin, n, fin, d, precision=PREC_FP32)
        if need_ew and ctx.has_w:
            m = graph.num_edges
            dew = _vec(m, torch.float32, g.device)
            tmp = _ws(4 * (2 * m + n), g.device, "edge_grad_tmp").view(torch.float32)
            rp_d, pm_d, nb_d, od_d = graph.csr_dst
            rp_s, pm_s, _, _ = graph.csr_src
            with _timed(f"edge_grad_d{d}"):
                if ctx.tab is not None:
                    check(lib().sgs_gcn_edge_grad_h16(_p(rp_d), _p(pm_d), _p(nb_d), _p(norm.what_dst), _p(od_d),
                                                      _p(rp_s), _p(pm_s), _p(graph.src), _p(graph.dst), _p(g),
                                                      _p(ctx.tab.data), _p(ctx.tab.scale), _p(norm.dis), _p(norm.deg),
                                                      _p(norm.loopw), m, n, d, _p(tmp), _p(tmp[m:]), _p(tmp[2 * m:]),
                                                      _p(dew), 0, _stream()), "sgs_gcn_edge_grad_h16")
                else:
                    check(lib().sgs_gcn_edge_grad(_p(rp_d), _p(pm_d), _p(nb_d), _p(norm.what_dst), _p(od_d), _p(rp_s),
                                                  _p(pm_s),
                                                  _p(graph.src), _p(graph.dst), _p(g), _p(h), _p(norm.dis),
                                                  _p(norm.deg), _p(norm.loopw), m, n, d, _p(tmp), _p(tmp[m:]),
                                                  _p(tmp[2 * m:]), _p(dew), 0, _stream()), "sgs_gcn_edge_grad")
            ctx.tab = None
        return dx, dw, db, dew, None, None, None, None


class GCNConvPairFn(torch.autograd.Function):
    """Two unweighted GCNConv layers over the SAME graph and the SAME static input in one pass:
        out_a = dropout(relu(A_hat (x Wa^T) + ba)),   out_b = dropout(relu(A_hat (x Wb^T) + bb)).
    One stacked projection x [Wa; Wb]^T, one fp16 gather table [N, 2D], ONE SpMM sweep (sgs_spmm_h16_pair): the
    gather kernels are bound by the rate of row gathers, not by their bytes, so the pair costs about one layer.
    Used for the scorer's gcn1 and the random baseline's gcn1, which both see the random subgraph
    (model.py:107 / :159, training_hybrid.py:45-48,93).  Backward: per half, the plain GCNConv backward."""

    @staticmethod
    def forward(ctx, x, wa, ba, wb, bb, graph, relu, p_drop, seed):
        x = _req(x, torch.float32, "x")
        wa, wb = _req(wa, torch.float32, "weight"), _req(wb, torch.float32, "weight")
        ctx.set_materialize_grads(False)     # an unused half (e.g. the random baseline on a learned-wins step) costs nothing
        norm = graph.norm(None)
        d = wa.size(0)
        h = linear_nt(x, torch.cat([wa, wb], 0), static_x=True)
        tab = gather_table(h)
        if tab is None:
            raise RuntimeError("gcn_conv_pair needs the fp16 gather mode and a width with 2 * D % 16 == 0, 2 * D <= 512")
        n = x.size(0)
        out_a = torch.empty(n, d, dtype=torch.float32, device=x.device)
        out_b = torch.empty_like(out_a)
        rowptr, _perm, nbr, order = graph.csr_dst
        flags = (SPMM_RELU if relu else 0) | (SPMM_DROPOUT if p_drop > 0 else 0)
        bias = torch.cat([_req(ba, torch.float32, "bias"), _req(bb, torch.float32, "bias")])
        with _timed(f"spmm_d{2 * d}_pair"):
            check(lib().sgs_spmm_h16_pair(_p(rowptr), _p(nbr), _p(norm.what_dst), _p(order), _p(norm.dis),
                                          _p(norm.loopw), _p(tab.data), _p(tab.scale), n, 2 * d, _p(bias), _p(out_a),
                                          _p(out_b), flags, float(p_drop), int(seed), _stream()), "sgs_spmm_h16_pair")
        ctx.graph, ctx.norm, ctx.relu, ctx.p_drop = graph, norm, relu, p_drop
        ctx.save_for_backward(x, out_a if relu else None, out_b if relu else None)
        return out_a, out_b

    @staticmethod
    def backward(ctx, ga, gb):
        x, out_a, out_b = ctx.saved_tensors
        graph, norm = ctx.graph, ctx.norm
        res = []
        for gout, out, (need_w, need_b) in ((ga, out_a, ctx.needs_input_grad[1:3]), (gb, out_b, ctx.needs_input_grad[3:5])):
            dw = db = None
            if gout is not None and (need_w or need_b):
                n, d = gout.shape
                gout = _req(gout, torch.float32, "grad")
                g = gout
                if ctx.relu:
                    g = torch.empty_like(gout)
                    scale = 1.0 / (1.0 - ctx.p_drop) if ctx.p_drop > 0 else 1.0
                    check(lib().sgs_act_bwd(_p(gout), _p(out), gout.numel(), scale, _p(g), _stream()), "sgs_act_bwd")
                if need_b:
                    db = torch.empty(d, dtype=torch.float32, device=g.device)
                    check(lib().sgs_colsum(_p(g), n, d, _p(db), _stream()), "sgs_colsum")
                if need_w:
                    dh = spmm(graph.csr_src, norm.what_src, norm, g, table=gather_table(g, scaled=True))
                    dw = gemm_tn(dh, x, static_b=True)
            res += [dw, db]
        return (None, res[0], res[1], res[2], res[3], None, None, None, None)


def gcn_conv_pair_available(x, wa, wb):
    """The one-sweep pair applies in the fp16 gather mode, for a static input and equal widths with 2 D <= 512."""
    d = wa.size(0)
    if __import__("os").environ.get("SGS_NO_PAIR"):      # A/B switch
        return False
    return (_state["gather"] == PREC_FP16 and not x.requires_grad and wa.shape == wb.shape and (2 * d) % 16 == 0
            and 128 <= 2 * d <= 512 and (x.size(0) * 2 * d) % 8 == 0)


def gcn_conv_pair(x, wa, ba, wb, bb, graph, relu=False, p_drop=0.0, seed=0):
    return GCNConvPairFn.apply(x, wa, ba, wb, bb, graph, relu, float(p_drop), int(seed))


def gcn_conv(x, weight, bias, graph, edge_weight=None, relu=False, p_drop=0.0, seed=0):
    return GCNConvFn.apply(x, weight, bias, edge_weight, graph, relu, float(p_drop), int(seed))


class SAGEConvFn(torch.autograd.Function):
    """PyG 2.3.1 SAGEConv defaults (mean aggregation + root weight) with the fused ReLU + dropout epilogue
    EdgeProbSAGE applies (model.py:50,63,66):   out = dropout(relu(mean_{j->i}(x_j) W_l^T + b_l + x_i W_r^T)).
    The projections commute with the mean, so both run first on the tcgen05 GEMM (F -> H) and the SpMM (K3b, no
    self term, edge weight 1/in-degree) gathers H-wide rows and accumulates onto the root term.
    Backward: the same SpMM over the by-source CSR, dW by the TN GEMM."""

    @staticmethod
    def _mean_weights(graph):
        c = graph._mean_w
        if c is None:
            rp_d, pm_d, _, _ = graph.csr_dst
            rp_s, _, nb_s, _ = graph.csr_src
            indeg = (rp_d[1:] - rp_d[:-1]).clamp(min=1).to(torch.float32)
            inv = 1.0 / indeg
            n = graph.num_nodes
            rows_d = torch.repeat_interleave(torch.arange(n, device=graph.device), (rp_d[1:] - rp_d[:-1]).long(),
                                             output_size=graph.num_edges)
            w_dst = inv[rows_d].contiguous()                      # by-destination CSR order: 1 / indeg(row)
            w_src = inv[nb_s[: graph.num_edges].long()].contiguous()   # by-source CSR order: 1 / indeg(dst of the edge)
            c = graph._mean_w = (w_dst, w_src)
        return c

    @staticmethod
    def forward(ctx, x, w_l, b_l, w_r, graph, relu, p_drop, seed):
        x = _req(x, torch.float32, "x")
        if x.size(0) != graph.num_nodes:
            raise RuntimeError("x must have one row per node")
        w_dst, _ = SAGEConvFn._mean_weights(graph)
        static = not x.requires_grad
        h_l = linear_nt(x, _req(w_l, torch.float32, "lin_l.weight"), static_x=static)
        out = linear_nt(x, _req(w_r, torch.float32, "lin_r.weight"), static_x=static, remember=False)   # root term,
        # accumulated into by the SpMM below
        spmm(graph.csr_dst, w_dst, None, h_l, _req(b_l, torch.float32, "lin_l.bias"), relu, p_drop, seed, out=out,
             add_root=True)
        ctx.graph, ctx.relu, ctx.p_drop = graph, relu, p_drop
        ctx.save_for_backward(x, w_l, w_r, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, w_l, w_r, out = ctx.saved_tensors
        graph = ctx.graph
        n, d = gout.shape
        gout = _req(gout, torch.float32, "grad")
        if ctx.relu:
            g = torch.empty_like(gout)
            scale = 1.0 / (1.0 - ctx.p_drop) if ctx.p_drop > 0 else 1.0
            check(lib().sgs_act_bwd(_p(gout), _p(out), gout.numel(), scale, _p(g), _stream()), "sgs_act_bwd")
        else:
            g = gout
        need_x, need_wl, need_bl, need_wr = ctx.needs_input_grad[:4]
        dx = dwl = dbl = dwr = None
        if need_bl:
            dbl = torch.empty(d, dtype=torch.float32, device=g.device)
            check(lib().sgs_colsum(_p(g), n, d, _p(dbl), _stream()), "sgs_colsum")
        if need_wr:
            dwr = gemm_tn(g, x, static_b=not x.requires_grad)
        if need_wl or need_x:
            _, w_src = SAGEConvFn._mean_weights(graph)
            dh = spmm(graph.csr_src, w_src, None, g)
            if need_wl:
                dwl = gemm_tn(dh, x, static_b=not x.requires_grad)
            if need_x:
                prec = PREC_TF32 if _state["gemm"] != PREC_FP32 else PREC_FP32
                dx = linear_nt(dh, w_l.t().contiguous(), precision=prec) + linear_nt(g, w_r.t().contiguous(),
                                                                                     precision=prec)
        return dx, dwl, dbl, dwr, None, None, None, None


def sage_conv(x, w_l, b_l, w_r, graph, relu=False, p_drop=0.0, seed=0):
    return SAGEConvFn.apply(x, w_l, b_l, w_r, graph, relu, float(p_drop), int(seed))


# ------------------------------------------------------------------------------------------
# edge scorer
# ------------------------------------------------------------------------------------------

def edge_score_forward(out, graph, w1, b1, w2, b2, ids=None, p_drop=0.0, seed=0, precision=None):
    """p[n] for edges `ids` (int32) or all edges of `graph` (no autograd)."""
    prec = _state["scorer"] if precision is None else precision
    out = _req(out, torch.float32, "out")
    n_nodes, h = out.shape
    n = graph.num_edges if ids is None else int(ids.numel())
    p = _vec(n, torch.float32, out.device)
    nbytes = lib().sgs_edge_score_workspace_bytes(n, n_nodes, h, prec, 0)
    ws = _ws(nbytes, out.device, "edge_score_fwd")
    with _timed("edge_score_fwd"):
        check(lib().sgs_edge_score_fwd(_p(out), n_nodes, h, _p(graph.src), _p(graph.dst), _p(ids), n, _p(w1),
                                       _p(b1), _p(w2), _p(b2), float(p_drop), int(seed), _p(p), _p(ws), ws.numel(),
                                       prec, _stream()), "sgs_edge_score_fwd")
    return p


class EdgeScoreFn(torch.autograd.Function):
    """_edge_score (model.py:115-122) over all edges or an id subset.  `precomputed` lets the
    hybrid pipeline reuse the no-grad full-graph probabilities for the forward value while the
    backward still runs (recompute + gradient) over just these edges."""

    @staticmethod
    def forward(ctx, out, w1, b1, w2, b2, graph, ids, p_drop, seed, precomputed, precision):
        ctx.w2_shape, ctx.b2_shape = w2.shape, b2.shape
        w1 = _req(w1, torch.float32, "fc1.weight")
        b1 = _req(b1, torch.float32, "fc1.bias")
        w2 = _req(w2.reshape(-1), torch.float32, "fc2.weight")
        b2 = _req(b2.reshape(-1), torch.float32, "fc2.bias")
        out = _req(out, torch.float32, "out")
        if precomputed is None:
            p = edge_score_forward(out, graph, w1, b1, w2, b2, ids, p_drop, seed, precision)
        else:
            p = precomputed
        ctx.graph, ctx.ids, ctx.p_drop, ctx.seed = graph, ids, p_drop, seed
        ctx.prec = _state["scorer"] if precision is None else precision
        ctx.save_for_backward(out, w1, b1, w2, b2, p)
        return p

    @staticmethod
    def backward(ctx, dp):
        out, w1, b1, w2, b2, p_fwd = ctx.saved_tensors
        graph, ids = ctx.graph, ctx.ids
        dp = _req(dp, torch.float32, "grad")
        n_nodes, h = out.shape
        n = dp.numel()
        dev = out.device
        d_out = torch.zeros_like(out)
        dw1 = torch.zeros_like(w1)
        small = torch.zeros(2 * h + 1, dtype=torch.float32, device=dev)
        nbytes = lib().sgs_edge_score_workspace_bytes(n, n_nodes, h, ctx.prec, 1)
        ws = _ws(nbytes, dev, "edge_score_bwd")
        with _timed("edge_score_bwd"):
            check(lib().sgs_edge_score_bwd(_p(out), n_nodes, h, _p(graph.src), _p(graph.dst), _p(ids), n, _p(w1),
                                           _p(b1), _p(w2), _p(b2), float(ctx.p_drop), int(ctx.seed), _p(p_fwd),
                                           _p(dp), _p(d_out), _p(dw1), _p(small), _p(small[h:]),
                                           _p(small[2 * h:]), _p(ws), ws.numel(), ctx.prec, _stream()),
                      "sgs_edge_score_bwd")
        db1, dw2, db2 = small[:h], small[h:2 * h].reshape(ctx.w2_shape), small[2 * h:].reshape(ctx.b2_shape)
        return d_out, dw1, db1, dw2, db2, None, None, None, None, None, None


def edge_score(out, w1, b1, w2, b2, graph, ids=None, p_drop=0.0, seed=0, precomputed=None, precision=None):
    return EdgeScoreFn.apply(out, w1, b1, w2, b2, graph, ids, float(p_drop), int(seed), precomputed, precision)


# ------------------------------------------------------------------------------------------
# sampler
# ------------------------------------------------------------------------------------------

def sum_f32(p):
    p = _req(p, torch.float32, "p")
    s = torch.empty(1, dtype=torch.float32, device=p.device)
    ws = _ws(8192, p.device)
    check(lib().sgs_sum_f32(_p(p), p.numel(), _p(s), _p(ws), ws.numel(), _stream()), "sgs_sum_f32")
    return s


def sum_f64(p):
    """sum(p) as a device float64 [1]: the fp64 block partials of sgs_sum_f32 (1024 doubles) summed in fp64 -- the
    multi-GPU sampler all-reduces this before rounding once to fp32 (no [E] fp64 temporary)."""
    p = _req(p, torch.float32, "p")
    s = torch.empty(1, dtype=torch.float32, device=p.device)
    ws = _ws(8192, p.device)
    check(lib().sgs_sum_f32(_p(p), p.numel(), _p(s), _p(ws), ws.numel(), _stream()), "sgs_sum_f32")
    return ws.view(torch.float64)[:1024].sum().reshape(1)


def softmax_f32(x):
    x = _req(x, torch.float32, "x")
    out = torch.empty_like(x)
    ws = _ws(16384, x.device)
    check(lib().sgs_softmax_f32(_p(x), x.numel(), _p(out), _p(ws), ws.numel(), _stream()), "sgs_softmax_f32")
    return out


def exponential(n, device, seed=None, gid=None):
    """Exp(1) noise from the counter-based generator keyed (seed, edge id).  `gid` (int64 [n], global edge ids of a
    shard): element i is the value the contiguous draw gives edge gid[i], whatever the number of shards."""
    out = torch.empty(n, dtype=torch.float32, device=device)
    seed = next_seed() if seed is None else int(seed)
    if gid is not None:
        gid = _req(gid, torch.int64, "gid")
        check(lib().sgs_exponential_ids_f32(_p(out), _p(gid), n, seed, _stream()), "sgs_exponential_ids_f32")
    else:
        check(lib().sgs_exponential_f32(_p(out), n, seed, _stream()), "sgs_exponential_f32")
    return out


def _coefs(coef):
    # exactly as Python evaluates sampling.py:95: (1 - coef) in double, then each scalar -> fp32
    return float(np.float32(1.0 - float(coef))), float(np.float32(float(coef)))


class TopQ:
    """Result of one draw: sel (int32 [q], ascending edge ids), optional mask (uint8 [E]) and the
    device state vector (tau bits, counts, invalid-input flag)."""
    __slots__ = ("sel", "mask", "state", "S", "mode", "coef", "q", "E")

    def check_valid(self):
        st = self.state.cpu()
        if int(st[5]) != 0:
            raise RuntimeError("probability tensor contains either `inf`, `nan` or element < 0")
        if int(st[7]) != self.q:
            raise RuntimeError(f"sampler selected {int(st[7])} edges, expected {self.q}")
        return st

    @property
    def tau(self):
        bits = int(self.state[2].item()) & 0xFFFFFFFF
        return float(np.array([bits], dtype=np.uint32).view(np.float32)[0])


def sample_topq(p, prob, q, mode=SAMPLE_TRAIN, coef=0.3, noise=None, S=None, want_mask=False, seed=None,
                validate=True):
    """Top-q of key = s/noise (include/sgs_b200.h K2).  noise: injected Exp(1) tensor, or None to
    draw it on device.  S: injected normaliser (device float tensor [1]) or None to reduce p."""
    p = _req(p, torch.float32, "edge_probs")
    e = p.numel()
    q = int(q)
    if q > e:
        raise RuntimeError("cannot sample n_sample > prob_dist.size(-1) samples without replacement")
    if q < 1:
        raise RuntimeError("cannot sample n_sample <= 0 samples")
    dev = p.device
    if mode == SAMPLE_TRAIN:
        prob = _req(prob, torch.float32, "batch.prob")
        if prob.numel() != e:
            raise RuntimeError(f"batch.prob has {prob.numel()} entries but edge_probs has {e}")
    else:
        prob = None
    if noise is None:
        noise = exponential(e, dev, seed)
    else:
        noise = _req(noise, torch.float32, "noise")
        if noise.numel() != e:
            raise RuntimeError("noise must have one entry per edge")
    if mode != SAMPLE_RAW and S is None:
        S = sum_f32(p)
    one_m, c = _coefs(coef)
    keys = _ws(4 * e, dev, "topq_keys").view(torch.int32)
    sel = _vec(q, torch.int32, dev)
    mask = torch.empty(e, dtype=torch.uint8, device=dev) if want_mask else None
    state = torch.empty(8, dtype=torch.int64, device=dev)
    nbytes = lib().sgs_topq_workspace_bytes(e) + _lib.TOPQ_BINS * 8
    ws = _ws(nbytes, dev, "topq_ws")
    with _timed("sample_topq"):
        check(lib().sgs_sample_topq(_p(p), _p(prob), _p(noise), e, q, one_m, c, mode, _p(S), _p(keys), _p(sel),
                                    _p(mask), _p(state), _p(ws), ws.numel(), _stream()), "sgs_sample_topq")
    r = TopQ()
    r.sel, r.mask, r.state, r.S, r.mode, r.coef, r.q, r.E = sel, mask, state, S, mode, coef, q, e
    if validate:
        r.check_valid()
    return r


def gather_selected(p, prob, sel, mode, coef, S, straight_through=False):
    q = sel.numel()
    one_m, c = _coefs(coef)
    p_sel = _vec(q, torch.float32, p.device)
    w_st = _vec(q, torch.float32, p.device) if straight_through else None
    check(lib().sgs_gather_selected(_p(p), _p(prob), _p(sel), q, one_m, c, mode, _p(S), _p(p_sel), _p(w_st),
                                    _stream()), "sgs_gather_selected")
    return p_sel, w_st


class GatherSelectedFn(torch.autograd.Function):
    """p_full[mask] (training_hybrid.py:86): gather forward, scatter into zeros[E] backward."""

    @staticmethod
    def forward(ctx, p_full, sel):
        p_full = _req(p_full, torch.float32, "edge_probs")
        ctx.sel, ctx.e = sel, p_full.numel()
        return gather_selected(p_full, None, sel, SAMPLE_RAW, 0.0, None)[0]

    @staticmethod
    def backward(ctx, g):
        g = _req(g, torch.float32, "grad")
        out = torch.zeros(ctx.e, dtype=torch.float32, device=g.device)
        check(lib().sgs_scatter_selected(_p(g), _p(ctx.sel), g.numel(), _p(out), _stream()), "sgs_scatter_selected")
        return out, None


class StraightThroughWeightsFn(torch.autograd.Function):
    """(p * st)[mask].clamp(0,1) with st = (one_hot - s).detach() + s (sampling.py:137-155) and the
    dense backward of SURVEY A.4:
        dL/dp_j = gamma_j (st_j + a p_j / S') - (a / S'^2) sum_i gamma_i p_i^2."""

    @staticmethod
    def forward(ctx, p_full, prob, sel, S, mode, coef):
        p_full = _req(p_full, torch.float32, "edge_probs")
        p_sel, w = gather_selected(p_full, prob, sel, mode, coef, S, straight_through=True)
        ctx.sel, ctx.mode, ctx.coef, ctx.e = sel, mode, coef, p_full.numel()
        ctx.save_for_backward(p_sel, w, S)
        return w

    @staticmethod
    def backward(ctx, g):
        p_sel, w, S = ctx.saved_tensors
        a = float(np.float32(1.0 - ctx.coef)) if ctx.mode == SAMPLE_TRAIN else 1.0
        s_eff = S + 1e-12
        gamma = torch.where((w > 0) & (w < 1), g, torch.zeros_like(g))
        const = (a / (s_eff * s_eff)) * (gamma * p_sel * p_sel).sum()
        sel_term = gamma * (1.0 + a * p_sel / s_eff)
        out = (-const).expand(ctx.e).contiguous()
        check(lib().sgs_scatter_selected(_p(sel_term.contiguous()), _p(ctx.sel), sel_term.numel(), _p(out),
                                         _stream()), "sgs_scatter_selected")
        return out, None, None, None, None, None


# ------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------

class LossStats:
    """Device accumulator of sgs_loss_fwd: {ce_sum, n_train, correct, bce_sum, n_valid, sum_label,
    mse_sum, q} as float64[8]."""
    __slots__ = ("acc",)


class EdgeLossState:
    """What the fused edge pass (sgs_loss_fwd_fused) leaves behind for the backward: the unscaled gradients of the
    reg1 / reg2 terms w.r.t. the sampled edge probabilities and the logits."""
    __slots__ = ("dlog_e", "u1", "u2", "q")

    def __init__(self, dlog_e, u1, u2, q):
        self.dlog_e, self.u1, self.u2, self.q = dlog_e, u1, u2, q


_FUSED_MAX_CLASSES = 64


def edge_state_of(acc):
    """The EdgeLossState the fused forward attached to its accumulator tensor (None for the two-kernel path).  A
    caller that re-slices / all-reduces `acc` (sharded step) keeps the state and passes it to fused_loss itself."""
    return getattr(acc, "_sgs_edge_state", None)


def loss_forward(logits, y, train_mask_u8, sub=None, p_s=None, row_mask_u8=None):
    logits = _req(logits, torch.float32, "logits")
    n, c = logits.shape
    acc = torch.empty(8, dtype=torch.float64, device=logits.device)
    with_edges = sub is not None
    q = sub.num_edges if with_edges else 0
    if with_edges and q > 0 and c <= _FUSED_MAX_CLASSES:
        # learned step: one sweep over the sampled edges computes the loss sums and the unscaled edge-term gradients
        p_s = _req(p_s, torch.float32, "edge_probs")
        dlog_e = torch.empty_like(logits)
        u1, u2 = _vec(q, torch.float32, logits.device), _vec(q, torch.float32, logits.device)
        code = torch.empty(n, dtype=torch.int32, device=logits.device)
        with _timed("loss_fwd"):
            check(lib().sgs_loss_fwd_fused(_p(logits), n, c, _p(y), _p(train_mask_u8), _p(row_mask_u8), _p(sub.src),
                                           _p(sub.dst), _p(p_s), q, _p(acc), _p(dlog_e), _p(u1), _p(u2), _p(code),
                                           _stream()), "sgs_loss_fwd_fused")
        acc._sgs_edge_state = EdgeLossState(dlog_e, u1, u2, q)
        return acc
    check(lib().sgs_loss_fwd(_p(logits), n, c, _p(y), _p(train_mask_u8), _p(row_mask_u8),
                             _p(sub.src) if with_edges else None,
                             _p(sub.dst) if with_edges else None, _p(p_s) if with_edges else None, q,
                             1 if with_edges else 0, _p(acc), _stream()), "sgs_loss_fwd")
    return acc


class FusedLossFn(torch.autograd.Function):
    """CE(logits[train], y[train]) [+ c1*BCE_reg1 + c2*MSE_reg2 over the sampled edges]
    (training_hybrid.py:103-132).  `acc` may be a precomputed sgs_loss_fwd accumulator."""

    @staticmethod
    def forward(ctx, logits, p_s, y, train_mask_u8, sub, c0, c1, c2, reg1, reg2, acc, row_mask_u8=None,
                edge_state=None):
        logits = _req(logits, torch.float32, "logits")
        with_edges = sub is not None and p_s is not None and (reg1 or reg2)
        if with_edges:
            p_s = _req(p_s, torch.float32, "edge_probs")
        if acc is None:
            acc = loss_forward(logits, y, train_mask_u8, sub if with_edges else None, p_s if with_edges else None,
                               row_mask_u8)
        if edge_state is None and with_edges:
            edge_state = edge_state_of(acc)
        ctx.edge_state = edge_state if with_edges else None
        loss = torch.empty(1, dtype=torch.float32, device=logits.device)
        check(lib().sgs_loss_finish(_p(acc), c0, c1, c2, 1 if (reg1 and with_edges) else 0,
                                    1 if (reg2 and with_edges) else 0, _p(loss), _stream()), "sgs_loss_finish")
        ctx.sub, ctx.c0, ctx.c1, ctx.c2, ctx.reg1, ctx.reg2, ctx.with_edges = sub, c0, c1, c2, reg1, reg2, with_edges
        ctx.save_for_backward(logits, p_s if with_edges else None, y, train_mask_u8, acc, row_mask_u8)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        logits, p_s, y, tm, acc, rm = ctx.saved_tensors
        n, c = logits.shape
        sub = ctx.sub
        g = g.reshape(1).to(torch.float32).contiguous()
        dlogits = torch.empty_like(logits)
        dp = _vec(p_s.numel(), torch.float32, p_s.device) if ctx.with_edges else None
        q = sub.num_edges if ctx.with_edges else 0
        est = ctx.edge_state
        if est is not None:
            with _timed("loss_bwd"):
                check(lib().sgs_loss_bwd_fused(_p(logits), n, c, _p(y), _p(tm), _p(rm), est.q, _p(acc), ctx.c0, ctx.c1,
                                               ctx.c2, 1 if ctx.reg1 else 0, 1 if ctx.reg2 else 0, _p(g),
                                               _p(est.dlog_e), _p(est.u1), _p(est.u2), _p(dlogits), _p(dp),
                                               _stream()), "sgs_loss_bwd_fused")
            ctx.edge_state = None
            return dlogits, dp, None, None, None, None, None, None, None, None, None, None, None
        with _timed("loss_bwd"):
          check(lib().sgs_loss_bwd(_p(logits), n, c, _p(y), _p(tm), _p(rm), _p(sub.src) if ctx.with_edges else None,
                                 _p(sub.dst) if ctx.with_edges else None, _p(p_s), q, 1 if ctx.with_edges else 0,
                                 _p(acc), ctx.c0, ctx.c1, ctx.c2, 1 if ctx.reg1 else 0, 1 if ctx.reg2 else 0, _p(g),
                                 _p(dlogits), _p(dp), _stream()), "sgs_loss_bwd")
        return dlogits, dp, None, None, None, None, None, None, None, None, None, None, None


def fused_loss(logits, y, train_mask_u8, p_s=None, sub=None, c1=1.0, c2=0.5, reg1=True, reg2=True, acc=None,
               c0=1.0, row_mask_u8=None, edge_state=None):
    return FusedLossFn.apply(logits, p_s, y, train_mask_u8, sub, float(c0), float(c1), float(c2), bool(reg1),
                             bool(reg2), acc, row_mask_u8, edge_state)
