"""Seeded synthetic graphs of the shapes BASELINE.json names (there is no network
for the real datasets).  Device-agnostic torch code: the bench builds Reddit-shaped
graphs directly in HBM, the CPU tests build small ones on the host.

Follows SURVEY.md section 8(d): labels ~ U{0..C-1}; undirected simple graph with
exactly E directed edges (both directions, no self loops, coalesced, sorted by
(src, dst) like PyG datasets); heavy-tailed degrees; edge homophily h in the
spirit of `generate_synthetic` in the reference's Dataset.ipynb (cell 31);
x ~ (mu_y + N(0, I)) * 0.5; degree prior `prob` per datasets.py:141-156.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

SHAPES = {
    # name: (N, E, F, C, train_frac, val_frac, homophily)
    "smallcora": (2708, 10556, 1433, 7, 0.2, 0.4, 0.81),
    "reddit": (232965, 114615892, 602, 41, 0.66, 0.10, 0.76),
    "amazon-ratings": (24492, 93050, 300, 5, 0.2, 0.4, 0.38),
    "arxiv-year": (169343, 1166243, 128, 5, 0.2, 0.4, 0.22),
    "ogbn-products": (2449029, 61859140, 100, 47, 0.08, 0.02, 0.81),
}


class Batch:
    """Duck type of the PyG `Data` batch the reference's train() consumes
    (SURVEY 8b): x, y, edge_index, train/val/test masks, prob, .to(device)."""

    _FIELDS = ("x", "y", "edge_index", "train_mask", "val_mask", "test_mask", "prob")

    def __init__(self, **kw):
        for k in self._FIELDS:
            setattr(self, k, kw.get(k))
        self.num_classes = kw.get("num_classes")

    @property
    def num_nodes(self):
        return self.x.size(0)

    @property
    def num_edges(self):
        return self.edge_index.size(1)

    def to(self, device, non_blocking=False):
        dev = torch.device(device)
        if self.x.device == dev or (dev.type == "cuda" and dev.index is None
                                    and self.x.device.type == "cuda"):
            return self
        kw = {k: getattr(self, k).to(dev, non_blocking=non_blocking)
              for k in self._FIELDS if getattr(self, k) is not None}
        return Batch(num_classes=self.num_classes, **kw)

    def pin_memory(self):
        kw = {k: getattr(self, k).pin_memory() for k in self._FIELDS if getattr(self, k) is not None}
        out = Batch(num_classes=self.num_classes, **kw)
        if getattr(self, "_src_rowptr", None) is not None:
            out._src_rowptr = self._src_rowptr.pin_memory()
            out._dst_row = out.edge_index[1]           # (a view of the pinned edge_index)
        return out

    def compact(self, rowptr=True):
        """Same batch with `edge_index` narrowed to int32 [2, E] (node ids fit 31 bits): half the bytes on the host
        link; ops.Graph uses the int32 rows as they are (the reference's int64 form stays supported everywhere).
        rowptr: when the edge list is sorted by source (coalesced PyG datasets are), `upload_async` sends the source
        row as its CSR row pointer (N + 1 ints instead of E) and rebuilds it on the device, bit for bit."""
        kw = {k: getattr(self, k) for k in self._FIELDS if getattr(self, k) is not None}
        if kw["edge_index"].dtype != torch.int32:
            if self.num_nodes >= 2 ** 31:
                raise RuntimeError("node ids do not fit int32")
            kw["edge_index"] = kw["edge_index"].to(torch.int32)
        out = Batch(num_classes=self.num_classes, **kw)
        ei = out.edge_index
        e, n = ei.size(1), self.num_nodes
        if rowptr and ei.device.type == "cpu" and 0 < e < 2 ** 31 and e > n + 1:
            src = ei[0]
            if bool((src[1:] >= src[:-1]).all()) and int(src[0]) >= 0 and int(src[-1]) < n:
                nodes = torch.arange(n + 1, dtype=torch.int32)
                out._src_rowptr = torch.searchsorted(src, nodes, right=False).to(torch.int32)
                out._dst_row = ei[1]
        return out

    def upload_async(self, dev, stream):
        """Device copy whose H2D transfers are enqueued on `stream` (see loader.prefetch)."""
        from .loader import copy_fields_async
        rp = getattr(self, "_src_rowptr", None)
        fields = {k: getattr(self, k) for k in self._FIELDS}
        if rp is not None:
            fields["edge_index"] = None
        kw = copy_fields_async(fields, dev, stream)
        if rp is not None:
            e, n = self.edge_index.size(1), self.num_nodes
            ei = torch.empty(2, e, dtype=torch.int32, device=dev)
            with torch.cuda.stream(stream):
                ei[1].copy_(self._dst_row, non_blocking=True)
                rp_d = rp.to(dev, non_blocking=True)
                deg = (rp_d[1:] - rp_d[:-1]).to(torch.int64)
                ei[0].copy_(torch.repeat_interleave(torch.arange(n, dtype=torch.int32, device=dev), deg,
                                                    output_size=e))
            kw["edge_index"] = ei
        return Batch(num_classes=self.num_classes, **{k: v for k, v in kw.items() if v is not None})

    def nbytes(self):
        return sum(getattr(self, k).numel() * getattr(self, k).element_size()
                   for k in self._FIELDS if getattr(self, k) is not None)

    def upload_nbytes(self):
        """Bytes `upload_async` sends over the host link."""
        nb = self.nbytes()
        rp = getattr(self, "_src_rowptr", None)
        if rp is not None:
            nb += rp.numel() * rp.element_size() - self.edge_index[0].numel() * self.edge_index.element_size()
        return nb


def _weighted_pick(cdf, u):
    return torch.searchsorted(cdf, u).clamp_(max=cdf.numel() - 1)


def degree_prior(edge_index, num_nodes):
    """`data.prob` (datasets.py:141-156) with torch ops on any device."""
    row, col = edge_index[0], edge_index[1]
    e = row.numel()
    colcount = torch.bincount(col, minlength=num_nodes)
    rowcount = torch.bincount(row, minlength=num_nodes)
    deg_in = 1.0 / colcount
    deg_out = 1.0 / rowcount
    prob = (1.0 / deg_in[row]) + (1.0 / deg_out[col])
    prob = 1.0 / (prob + 1e-10)
    return F.softmax(prob * e ** -0.5, dim=0)


def make_edges(n, e, y, homophily, seed, device, exponent=2.2, undirected=True):
    """Exactly `e` directed edges, coalesced, no self loops, sorted by (src,dst)."""
    g = torch.Generator(device=device).manual_seed(seed)
    c = int(y.max()) + 1
    # heavy-tailed node weights  w ~ u^(-1/(exponent-1)), clipped
    u = torch.rand(n, generator=g, device=device).clamp_(min=1e-6)
    wt = u.pow(-1.0 / (exponent - 1.0)).clamp_(max=float(n) ** 0.5 * 4)
    cdf_all = torch.cumsum(wt.double(), 0)
    cdf_all /= cdf_all[-1].clone()
    # per-class cdf over class-sorted nodes
    order = torch.argsort(y, stable=True)
    wt_sorted = wt[order].double()
    cls_count = torch.bincount(y, minlength=c)
    cls_off = torch.cumsum(cls_count, 0) - cls_count
    csum = torch.cumsum(wt_sorted, 0)
    cls_base = torch.cat([csum.new_zeros(1), csum])[cls_off]          # mass before class
    cls_mass = torch.cat([csum.new_zeros(1), csum])[cls_off + cls_count] - cls_base

    target_pairs = (e + 1) // 2 if undirected else e
    have = None
    want = target_pairs
    rounds = 0
    while True:
        m = int(want * 1.15) + 1024
        a = _weighted_pick(cdf_all, torch.rand(m, generator=g, device=device, dtype=torch.float64))
        b_any = _weighted_pick(cdf_all, torch.rand(m, generator=g, device=device, dtype=torch.float64))
        ca = y[a]
        pos = cls_base[ca] + torch.rand(m, generator=g, device=device, dtype=torch.float64) * cls_mass[ca]
        b_same = order[torch.searchsorted(csum, pos).clamp_(max=n - 1)]
        same = torch.rand(m, generator=g, device=device) < homophily
        b = torch.where(same, b_same, b_any)
        keep = a != b
        a, b = a[keep], b[keep]
        if undirected:
            lo, hi = torch.minimum(a, b), torch.maximum(a, b)
            key = lo * n + hi
        else:
            key = a * n + b
        key = torch.unique(key if have is None else torch.cat([have, key]))
        rounds += 1
        if key.numel() >= target_pairs or rounds > 20:
            break
        have = key
        want = target_pairs - key.numel()
    if key.numel() < target_pairs:
        raise RuntimeError("graph too dense for the generator")
    if key.numel() > target_pairs:
        sel = torch.randperm(key.numel(), generator=g, device=device)[:target_pairs]
        key = key[sel]
    if undirected:
        lo, hi = key // n, key % n
        key = torch.cat([lo * n + hi, hi * n + lo])
        if key.numel() > e:            # odd e: drop one direction of one pair
            key = key[:e]
    key = torch.sort(key).values
    return torch.stack([key // n, key % n])


def make_graph(shape="smallcora", seed=42, device="cpu", scale=1.0, undirected=None,
               n=None, e=None, f=None, c=None, homophily=None, train_frac=None, val_frac=None):
    """Build a `Batch`.  `shape` names a row of SHAPES; `scale` shrinks N and E
    together (used for bounded CPU-baseline samples); explicit n/e/f/c override."""
    if shape is not None:
        n0, e0, f0, c0, tf0, vf0, h0 = SHAPES[shape]
    else:
        n0, e0, f0, c0, tf0, vf0, h0 = n, e, f, c, 0.2, 0.4, 0.5
    n = int(n if n is not None else max(8, round(n0 * scale)))
    e = int(e if e is not None else max(8, round(e0 * scale)))
    f = int(f if f is not None else f0)
    c = int(c if c is not None else c0)
    h = homophily if homophily is not None else h0
    tf = train_frac if train_frac is not None else tf0
    vf = val_frac if val_frac is not None else vf0
    if undirected is None:
        undirected = e % 2 == 0
    device = torch.device(device)
    g = torch.Generator(device=device).manual_seed(seed)
    y = torch.randint(0, c, (n,), generator=g, device=device)
    ei = make_edges(n, e, y, h, seed + 1, device, undirected=undirected)
    mu = torch.randn(c, f, generator=g, device=device)
    x = torch.randn(n, f, generator=g, device=device)
    x += mu[y]
    x *= 0.5
    perm = torch.randperm(n, generator=g, device=device)
    n_tr, n_va = int(n * tf), int(n * vf)
    tm = torch.zeros(n, dtype=torch.bool, device=device)
    vm = torch.zeros_like(tm)
    sm = torch.zeros_like(tm)
    tm[perm[:n_tr]] = True
    vm[perm[n_tr:n_tr + n_va]] = True
    sm[perm[n_tr + n_va:]] = True
    prob = degree_prior(ei, n)
    return Batch(x=x, y=y, edge_index=ei, train_mask=tm, val_mask=vm, test_mask=sm, prob=prob,
                 num_classes=c)


def make_clusters(shape, parts, seed=42, device="cpu", edges_per_part=None):
    """METIS-free "virtual clusters" (SURVEY 8f rank 3): `parts` independent graphs that together have the shape's
    edge count, standing in for the reference's ClusterData / ClusterLoader mini-batches (main.py:41-67: graphs with
    >= metis_threshold edges are cut into ceil(E / threshold) parts and inter-cluster edges dropped).  Every cluster
    is a `Batch` of its own (local node ids, its own degree prior); its node count is the shape's share of nodes, but
    never so small that edges_per_part distinct edges would not fit comfortably."""
    n0, e0 = SHAPES[shape][:2]
    e_p = int(edges_per_part if edges_per_part is not None else e0 // parts)
    e_p -= e_p % 2
    n_p = max(n0 // parts, int(8 * e_p ** 0.5) + 1)
    return [make_graph(shape, seed=seed + 7 * k, device=device, n=n_p, e=e_p) for k in range(parts)]
