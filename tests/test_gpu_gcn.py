"""GPU parity: K3 (gcn_norm, SpMM fwd/bwd, edge-weight gradient) and K4 (dense contraction) through
the C ABI vs the CPU oracle.  Tolerance: 1e-4 relative (north_star), measured against max-norm."""
import numpy as np
import pytest
import torch

from conftest import load_golden, t
from oracle import extended as ox

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def relerr(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.mark.parametrize("m,n,k", [(64, 64, 16), (300, 41, 256), (1000, 256, 602), (7, 5, 3), (256, 512, 5000)])
def test_gemm_forms(dev, m, n, k):
    from sgs_gnn_b200 import ops
    g = torch.Generator().manual_seed(m + n + k)
    a = torch.randn(m, k, generator=g)
    b = torch.randn(n, k, generator=g)
    want = (a.double() @ b.double().t()).float()
    ad, bd = a.to(dev), b.to(dev)
    assert relerr(ops.linear_nt(ad, bd).cpu(), want) < 1e-5
    # TN: C = A^T B with A [k,m], B [k,n]
    at, bt = a.t().contiguous().to(dev), b.t().contiguous().to(dev)
    c = ops.gemm(at, 1, m, bt, 1, n, m, n, k)
    assert relerr(c.cpu(), want) < 1e-5
    # NN with accumulate
    c0 = torch.ones(m, n, device=dev)
    c = ops.gemm(ad, k, 1, bt, 1, n, m, n, k, out=c0, accumulate=True)
    assert relerr(c.cpu(), want + 1.0) < 1e-5


def _rand_graph(n, m, seed, self_loops=False):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, m), generator=g)
    if not self_loops:
        ei[1] = torch.where(ei[0] == ei[1], (ei[1] + 1) % n, ei[1])
    return ei


@pytest.mark.parametrize("d", [41, 256, 64, 7])
@pytest.mark.parametrize("weighted", [False, True])
def test_gcn_conv_forward_backward(dev, d, weighted):
    from sgs_gnn_b200 import ops
    n, m, f = 500, 6000, 48
    ei = _rand_graph(n, m, d + weighted, self_loops=True)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, f, generator=g)
    W = torch.randn(d, f, generator=g) * 0.2
    b = torch.randn(d, generator=g) * 0.1
    w = torch.rand(m, generator=g) if weighted else None
    G = torch.randn(n, d, generator=g)

    xr, Wr, br = x.clone().requires_grad_(True), W.clone().requires_grad_(True), b.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True) if weighted else None
    out_ref = torch.relu(ox.gcn_conv(xr, Wr, br, ei, wr))
    grads_ref = torch.autograd.grad((out_ref * G).sum(), [xr, Wr, br] + ([wr] if weighted else []))

    xd, Wd, bd = (v.to(dev).requires_grad_(True) for v in (x, W, b))
    wd = w.to(dev).requires_grad_(True) if weighted else None
    graph = ops.graph_of(ei.to(dev), n)
    out = ops.gcn_conv(xd, Wd, bd, graph, wd, relu=True)
    assert relerr(out.detach().cpu(), out_ref.detach()) < RTOL
    grads = torch.autograd.grad((out * G.to(dev)).sum(), [xd, Wd, bd] + ([wd] if weighted else []))
    for name, a, r in zip(("dx", "dW", "db", "dw"), grads, grads_ref):
        if name == "dw":  # input self-loop edges: gradient not propagated (documented)
            keep = ei[0] != ei[1]
            assert relerr(a.cpu()[keep], r[keep]) < RTOL, name
        else:
            assert relerr(a.cpu(), r) < RTOL, name


@pytest.mark.parametrize("d", [41, 256])
def test_gcn_conv_hub_rows(dev, d):
    """Power-law hubs: rows with more than SGS_HEAVY_ROW_DEG (512) in- or out-edges take the block-cooperative
    path of the SpMM / SDDMM kernels; two hubs (one in, one out, 3000 and 1500 edges) among ordinary rows."""
    from sgs_gnn_b200 import ops
    n, f = 4000, 32
    g = torch.Generator().manual_seed(17 + d)
    base = _rand_graph(n, 8000, 3)
    hub_in = torch.stack([torch.randperm(n - 1, generator=g)[:3000] + 1, torch.zeros(3000, dtype=torch.int64)])
    hub_out = torch.stack([torch.full((1500,), 7, dtype=torch.int64), torch.randperm(n - 8, generator=g)[:1500] + 8])
    ei = torch.cat([base, hub_in, hub_out], 1)
    ei = ei[:, torch.randperm(ei.size(1), generator=g)]
    m = ei.size(1)
    x = torch.randn(n, f, generator=g)
    W = torch.randn(d, f, generator=g) * 0.2
    b = torch.randn(d, generator=g) * 0.1
    w = torch.rand(m, generator=g)
    G = torch.randn(n, d, generator=g)
    xr, Wr, br, wr = (v.clone().requires_grad_(True) for v in (x, W, b, w))
    out_ref = torch.relu(ox.gcn_conv(xr, Wr, br, ei, wr))
    grads_ref = torch.autograd.grad((out_ref * G).sum(), [xr, Wr, br, wr])
    xd, Wd, bd, wd = (v.to(dev).requires_grad_(True) for v in (x, W, b, w))
    graph = ops.graph_of(ei.to(dev), n)
    assert int(graph.csr_dst[3][n]) >= 1 and int(graph.csr_src[3][n]) >= 1      # hub rows were detected
    out = ops.gcn_conv(xd, Wd, bd, graph, wd, relu=True)
    assert relerr(out.detach().cpu(), out_ref.detach()) < RTOL
    grads = torch.autograd.grad((out * G.to(dev)).sum(), [xd, Wd, bd, wd])
    keep = ei[0] != ei[1]
    for name, a, r in zip(("dx", "dW", "db", "dw"), grads, grads_ref):
        if name == "dw":
            assert relerr(a.cpu()[keep], r[keep]) < RTOL, name
        else:
            assert relerr(a.cpu(), r) < RTOL, name


def test_gcn_norm_edge_cases(dev):
    from sgs_gnn_b200 import ops
    # isolated nodes, duplicate edges, an input self loop carrying a weight
    n = 6
    ei = torch.tensor([[0, 0, 1, 2, 2, 4], [1, 1, 0, 2, 3, 3]])
    w = torch.tensor([0.5, 0.25, 1.0, 0.3, 0.7, 0.9])
    row2, col2, w_hat, deg, dis, loop_w = ox.gcn_norm(ei, w, n)
    graph = ops.graph_of(ei.to(dev), n)
    nrm = graph.norm(w.to(dev))
    assert torch.allclose(nrm.deg.cpu(), deg, rtol=1e-6) and torch.allclose(nrm.dis.cpu(), dis, rtol=1e-6)
    assert torch.allclose(nrm.loopw.cpu(), loop_w)
    x = torch.eye(n)
    out = ops.gcn_conv(x.to(dev), torch.eye(n, device=dev), torch.zeros(n, device=dev), graph, w.to(dev))
    want = ox.gcn_conv(x, torch.eye(n), torch.zeros(n), ei, w)
    assert torch.allclose(out.cpu(), want, atol=1e-6)
    # empty edge set: out = x W^T + b
    e0 = torch.zeros(2, 0, dtype=torch.int64, device=dev)
    out = ops.gcn_conv(x.to(dev), torch.eye(n, device=dev), torch.ones(n, device=dev), ops.graph_of(e0, n), None)
    assert torch.allclose(out.cpu(), x + 1.0)


def test_fused_dropout_matches_host_mirror(dev):
    from sgs_gnn_b200 import ops, rng
    n, m, f, d = 300, 2000, 16, 64
    ei = _rand_graph(n, m, 1)
    g = torch.Generator().manual_seed(2)
    x, W, b = torch.randn(n, f, generator=g), torch.randn(d, f, generator=g), torch.zeros(d)
    graph = ops.graph_of(ei.to(dev), n)
    seed, p = 424242, 0.3
    out = ops.gcn_conv(x.to(dev), W.to(dev), b.to(dev), graph, None, relu=True, p_drop=p, seed=seed).cpu()
    keep = torch.from_numpy(rng.keep_mask(seed, np.arange(n), d, p))
    want = torch.relu(ox.gcn_conv(x, W, b, ei)) * keep / (1 - p)
    assert relerr(out, want) < RTOL
    assert abs(float(keep.float().mean()) - 0.7) < 0.02


def test_golden_gnn_forward(dev):
    from sgs_gnn_b200.model import GNNModel
    z = load_golden("forward_small.npz")
    sd = {k[3:]: t(v) for k, v in z.items() if k.startswith("sd.")}
    model = GNNModel(24, 32, 5, 0.3, "GCN")
    model.load_state_dict(sd)
    model = model.to(dev).eval()

    class D:
        x = t(z["x"], dev)

    rei = t(z["rand_edge_index"], dev)
    with torch.no_grad():
        lw = model(D, rei, t(z["w"], dev))
        lu = model(D, rei)
    assert relerr(lw.cpu(), t(z["logits_weighted"])) < RTOL
    assert relerr(lu.cpu(), t(z["logits_unweighted"])) < RTOL


def test_presorted_by_source_csr_equals_sorted_build(dev):
    """Ascending-id subgraphs of a (src,dst)-sorted edge list skip the by-source sort (sgs_csr_build_sorted):
    rowptr / perm / nbr / row order must equal the radix-sort build bit for bit; an unsorted parent keeps the sort."""
    from sgs_gnn_b200 import ops
    g = torch.Generator().manual_seed(11)
    n, e = 3000, 40000
    ei = torch.randint(0, n, (2, e), generator=g)
    key = ei[0] * n + ei[1]
    ei = ei[:, torch.argsort(key, stable=True)].contiguous()
    full = ops.graph_of(ei.to(dev), n)
    assert full.src_sorted
    ids = torch.sort(torch.randperm(e, generator=g)[:9000]).values.int().to(dev)
    fast = full.subgraph(ids, ascending=True)
    assert fast._src_sorted is True
    slow = ops.Graph(fast.src.clone(), fast.dst.clone(), n)
    for a, b in zip(fast.csr_src, slow.csr_src):
        assert torch.equal(a, b)
    for a, b in zip(fast.csr_dst, slow.csr_dst):
        assert torch.equal(a, b)
    shuffled = ops.graph_of(ei[:, torch.randperm(e, generator=g)].contiguous().to(dev), n)
    assert not shuffled.src_sorted
    assert shuffled.subgraph(ids, ascending=True)._src_sorted is None


def test_fp16_gather_tables_match_fp32_path(dev):
    """gather precision 'fp16' (D = 256 SpMM / SDDMM read an L2-resident fp16 copy of the gathered rows, fp32
    accumulation): forward (with the fused ReLU + dropout epilogue) within 1e-3 of the fp32 path relative to max|out|
    (stated bound: fp16 rows carry 11 significant bits); gradients of the linear layer (no ReLU: a pre-activation
    within 1e-3 of zero may flip its gate between the two modes, which is not what this test measures) within 2e-3."""
    from sgs_gnn_b200 import ops, synth
    b = synth.make_graph(None, seed=8, n=4000, e=120000, f=64, c=4).to(dev)
    g = ops.graph_of(b.edge_index, 4000)
    gen = torch.Generator(device=dev).manual_seed(2)
    w = torch.randn(256, 64, generator=gen, device=dev) * 0.1
    bias = torch.randn(256, generator=gen, device=dev) * 0.1
    ew = torch.rand(g.num_edges, generator=gen, device=dev)
    gout = torch.randn(4000, 256, generator=gen, device=dev)
    res = {}
    before = ops.get_precision()
    try:
        for mode in ("fp32", "fp16"):
            ops.set_precision(gather=mode)
            with torch.no_grad():
                act = ops.gcn_conv(b.x, w, bias, g, ew, relu=True, p_drop=0.2, seed=5)
            x = b.x.clone().requires_grad_(True)
            wt = w.clone().requires_grad_(True)
            bt = bias.clone().requires_grad_(True)
            e = ew.clone().requires_grad_(True)
            out = ops.gcn_conv(x, wt, bt, g, e)
            out.backward(gout)
            res[mode] = (act, out.detach(), x.grad, wt.grad, bt.grad, e.grad)
    finally:
        ops.set_precision(**before)
    rel = lambda a, c: float((a - c).abs().max() / (c.abs().max() + 1e-30))   # noqa: E731
    assert rel(res["fp16"][0], res["fp32"][0]) < 1e-3
    assert rel(res["fp16"][1], res["fp32"][1]) < 1e-3
    for i in range(2, 6):
        assert rel(res["fp16"][i], res["fp32"][i]) < 2e-3, i


def test_degree_prior_kernel_matches_reference_formula(dev):
    """datasets.add_degree (datasets.py:141-156) on the device -- sgs_degree_scores + sgs_softmax_f32 --
    against the oracle's restatement of the reference formula."""
    from types import SimpleNamespace
    from oracle import extended as ox
    from sgs_gnn_b200 import datasets, synth
    b = synth.make_graph("arxiv-year", seed=4, scale=0.05)
    want = ox.degree_prior(b.edge_index, b.num_nodes)
    data = SimpleNamespace(x=b.x.to(dev), edge_index=b.edge_index.to(dev), num_nodes=b.num_nodes)
    datasets.add_degree(data)
    got = data.prob.cpu()
    assert got.shape == want.shape and abs(float(got.sum()) - 1.0) < 1e-4
    # the kernels accumulate the softmax normaliser in fp64, torch's CPU softmax in fp32: a uniform ~1e-5 relative
    # offset of every entry (measured 1.5e-5 at E = 58 k), nothing edge-specific
    rel = (got - want).abs() / want
    assert float(rel.max()) < 1e-4 and float(rel.max() - rel.min()) < 2e-6


def test_gcn_conv_pair_equals_two_layers(dev):
    """ops.gcn_conv_pair (two GCN layers over one graph in ONE gather sweep) against two separate gcn_conv calls:
    forward bit-identical sums (same table values, same accumulation order), gradients of either half equal to the
    plain layer's."""
    from sgs_gnn_b200 import ops, synth
    b = synth.make_graph(None, seed=12, n=3000, e=90000, f=48, c=4).to(dev)
    g = ops.graph_of(b.edge_index, 3000)
    gen = torch.Generator(device=dev).manual_seed(4)
    ws = [torch.randn(256, 48, generator=gen, device=dev) * 0.1 for _ in range(2)]
    bs = [torch.randn(256, generator=gen, device=dev) * 0.1 for _ in range(2)]
    gouts = [torch.randn(3000, 256, generator=gen, device=dev) for _ in range(2)]
    before = ops.get_precision()
    try:
        ops.set_precision(gather="fp16")
        assert ops.gcn_conv_pair_available(b.x, ws[0], ws[1])
        pa = [t_.clone().requires_grad_(True) for t_ in (ws[0], bs[0], ws[1], bs[1])]
        oa, ob = ops.gcn_conv_pair(b.x, pa[0], pa[1], pa[2], pa[3], g, True, 0.0, 0)
        ref = []
        for i in range(2):
            w_, b_ = ws[i].clone().requires_grad_(True), bs[i].clone().requires_grad_(True)
            o = ops.gcn_conv(b.x, w_, b_, g, None, relu=True)
            o.backward(gouts[i])
            ref.append((o.detach(), w_.grad, b_.grad))
        assert float((oa - ref[0][0]).abs().max()) <= 1e-6 * float(ref[0][0].abs().max())
        assert float((ob - ref[1][0]).abs().max()) <= 1e-6 * float(ref[1][0].abs().max())
        # only the second half is back-propagated (a random-wins step): the first half's parameters keep grad None
        ob.backward(gouts[1])
        assert pa[0].grad is None and pa[1].grad is None
        rel = lambda a, c: float((a - c).abs().max() / (c.abs().max() + 1e-30))   # noqa: E731
        assert rel(pa[2].grad, ref[1][1]) < 1e-4 and rel(pa[3].grad, ref[1][2]) < 1e-4
    finally:
        ops.set_precision(**before)
