"""GPU: tensor-core (tcgen05) kernels against the fp32 CUDA-core parity path and the CPU oracle.
Stated bounds for 16-bit GEMM operands (north_star: "stated looser bound for bf16 GEMM inputs"):
    bf16 operands: |p - p_fp32| <= 4e-3 absolute (p in (0,1)), fp16 operands: <= 5e-4."""
import numpy as np
import pytest
import torch

from oracle import extended as ox

pytestmark = pytest.mark.gpu

BOUND = {"bf16": 4e-3, "fp16": 5e-4}


def _setup(n_nodes, e, h, seed):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n_nodes, (2, e), generator=g)
    out = torch.relu(torch.randn(n_nodes, h, generator=g))
    W1 = (torch.rand(h, 2 * h, generator=g) - 0.5) * (2 / (2 * h) ** 0.5)
    b1 = (torch.rand(h, generator=g) - 0.5) * 0.1
    w2 = (torch.rand(h, generator=g) - 0.5) * (2 / h ** 0.5)
    b2 = torch.tensor([0.05])
    return ei, out, W1, b1, w2, b2


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("h,e", [(256, 128), (256, 5000), (128, 777), (64, 3000), (256, 40001)])
def test_scorer_tc_forward_matches_fp32_path(dev, prec, h, e):
    from sgs_gnn_b200 import ops
    n_nodes = 700
    ei, out, W1, b1, w2, b2 = _setup(n_nodes, e, h, h + e)
    graph = ops.graph_of(ei.to(dev), n_nodes)
    args = [t.to(dev) for t in (out, W1, b1, w2, b2)]
    for p_drop in (0.0, 0.3):
        ref = ops.edge_score_forward(args[0], graph, *args[1:], None, p_drop, 17, ops.PREC_FP32)
        got = ops.edge_score_forward(args[0], graph, *args[1:], None, p_drop, 17, ops._PRECISION[prec])
        err = float((got - ref).abs().max())
        assert err <= BOUND[prec], (prec, h, e, p_drop, err)
    # oracle cross-check (no dropout)
    want = ox.edge_score(out, ei, W1, b1, w2.reshape(1, -1), b2, training=False).squeeze(-1)
    got = ops.edge_score_forward(args[0], graph, *args[1:], None, 0.0, 0, ops._PRECISION[prec]).cpu()
    assert float((got - want).abs().max()) <= BOUND[prec]


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_scorer_tc_forward_id_subset(dev, prec):
    from sgs_gnn_b200 import ops
    n_nodes, e, h = 500, 9000, 256
    ei, out, W1, b1, w2, b2 = _setup(n_nodes, e, h, 5)
    graph = ops.graph_of(ei.to(dev), n_nodes)
    args = [t.to(dev) for t in (out, W1, b1, w2, b2)]
    ids = torch.sort(torch.randperm(e)[:1234]).values.int().to(dev)
    full = ops.edge_score_forward(args[0], graph, *args[1:], None, 0.3, 99, ops._PRECISION[prec])
    sub = ops.edge_score_forward(args[0], graph, *args[1:], ids, 0.3, 99, ops._PRECISION[prec])
    # same edges, same dropout mask (keyed by edge id): identical up to accumulation order in TMEM
    assert float((sub - full[ids.long()]).abs().max()) <= 1e-6


# Gradients, 16-bit operands.  Two checks:
#  (1) against an fp32 emulation of the SAME quantised computation (operands rounded to 16 bit exactly where
#      the kernels round them, fp32 accumulation): tight bound -- validates the tcgen05 kernels themselves;
#  (2) against the fp32 CUDA-core parity path: looser stated bound, dominated by ReLU-mask flips of hidden
#      units whose pre-activation is within rounding distance of zero (dw2 / db2, which are insensitive to
#      flips, show the pure arithmetic error: ~3e-3 bf16, ~4e-4 fp16).
L2BOUND_VS_FP32 = {"bf16": 8e-2, "fp16": 4e-2}
L2BOUND_VS_EMUL = {"bf16": 3e-3, "fp16": 3e-3}
TORCH_T = {"bf16": torch.bfloat16, "fp16": torch.float16}


def _l2rel(a, r):
    a, r = a.flatten().double().cpu(), r.flatten().double().cpu()
    return float((a - r).norm() / (r.norm() + 1e-300))


def _emulated_grads(prec, out, ei, W1, b1, w2, b2, gup, keep, p_drop):
    """CPU emulation (fp64 accumulation) of the quantised computation the TC backward kernels perform
    (csrc/edge_score_bwd_tc.cu): operands are rounded to 16 bit exactly where the kernels round them --
    the node-embedding table, the features [x*y | x-y], W1 (recompute), the gate gradient
    G = S*dz/(1-p)*[pre>0]*keep and the BF operand diag(w2).W1 -- and dW1 / db1 / dw2 are derived from
    P = G^T F and g = colsum(G) as the BW epilogue does."""
    dt = TORCH_T[prec]
    q = lambda t: t.float().to(dt).double()
    h = W1.shape[0]
    w2v = w2.reshape(-1).double()
    absmax = float(gup.abs().max())
    S = 2.0 ** np.floor(np.log2(1024.0 / absmax))
    scale = 1.0 / (1.0 - p_drop) if keep is not None else 1.0
    oq = q(out)
    x, y = oq[ei[0]], oq[ei[1]]
    F = torch.cat([q(x * y), q(x - y)], 1)
    pre = F @ q(W1).t() + b1.double()
    mask = (pre > 0).double()
    if keep is not None:
        mask = mask * keep.double()
    hid = pre * mask * scale
    p = torch.sigmoid(hid @ w2v + b2.double().reshape(()))
    dz = gup.double() * p * (1 - p)
    G = q((S * dz * scale).float()[:, None] * mask.float())
    dF = (G @ q(w2v.float()[:, None] * W1)) / S
    dF1, dF2 = dF[:, :h], dF[:, h:]
    d_out = torch.zeros_like(oq)
    d_out.index_add_(0, ei[0], dF1 * y + dF2)
    d_out.index_add_(0, ei[1], dF1 * x - dF2)
    P = G.t() @ F / S
    g = G.sum(0) / S
    dW1 = w2v[:, None] * P
    db1 = w2v * g
    dw2 = ((W1.double() * P).sum(1) + b1.double() * g).reshape(1, -1)
    db2 = dz.sum().reshape(1)
    return d_out, dW1, db1, dw2, db2


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("h,e,subset,p_drop,n_nodes", [
    (256, 128, False, 0.0, 600), (256, 5000, False, 0.3, 600), (128, 3333, True, 0.3, 600), (256, 20011, True, 0.0, 600),
    # around the CTA-pair threshold of BA (n >= 256 edges) and single-edge lists
    (256, 255, False, 0.3, 600), (256, 256, False, 0.0, 600), (256, 257 * 3, True, 0.3, 600), (256, 1, False, 0.0, 600),
    (128, 1, False, 0.3, 600),
    # more nodes than one destination bucket holds (2^15 rows at H = 256, 2^16 at H = 128): the stable
    # destination-range partition of the edge list is exercised
    (256, 30000, True, 0.3, 100000), (128, 9000, False, 0.0, 140000)])
def test_scorer_tc_backward(dev, prec, h, e, subset, p_drop, n_nodes):
    from sgs_gnn_b200 import ops, rng
    ei, out, W1, b1, w2, b2 = _setup(n_nodes, e, h, h + e + 1)
    graph = ops.graph_of(ei.to(dev), n_nodes)
    g = torch.Generator().manual_seed(3)
    ids_c = torch.sort(torch.randperm(e, generator=g)[: max(1, e // 3)]).values if subset else torch.arange(e)
    ids = ids_c.int().to(dev) if subset else None
    n = ids_c.numel()
    gup = torch.randn(n, generator=g) * 1e-6               # realistic tiny upstream gradients (mean over q)
    seed = 1234
    grads = {}
    for mode in ("fp32", prec):
        leaves = [t.to(dev).clone().requires_grad_(True) for t in (out, W1, b1, w2.reshape(1, -1), b2)]
        p = ops.edge_score(*leaves, graph, ids, p_drop, seed, None, ops._PRECISION[mode])
        grads[mode] = torch.autograd.grad((p * gup.to(dev)).sum(), leaves)
    keep = torch.from_numpy(rng.keep_mask(seed, ids_c.numpy(), h, p_drop)).float() if p_drop > 0 else None
    emul = _emulated_grads(prec, out, ei[:, ids_c], W1, b1, w2, b2, gup, keep, p_drop)
    for name, a, r, em in zip(("d_out", "dW1", "db1", "dw2", "db2"), grads[prec], grads["fp32"], emul):
        assert _l2rel(a, em) <= L2BOUND_VS_EMUL[prec], ("vs emulation", name, prec, h, e, _l2rel(a, em))
        assert _l2rel(a, r) <= L2BOUND_VS_FP32[prec], ("vs fp32 path", name, prec, h, e, _l2rel(a, r))


@pytest.mark.parametrize("m,n,k", [(128, 128, 32), (1000, 256, 602), (333, 41, 256), (5000, 256, 256), (129, 7, 33)])
def test_gemm_tf32_tma(dev, m, n, k):
    """K4 tensor-core GEMM (tcgen05 kind::tf32 + TMA).  Stated bound for tf32 operands: 2e-3 of max |C|."""
    from sgs_gnn_b200 import ops
    g = torch.Generator().manual_seed(m + n + k)
    a = torch.randn(m, k, generator=g)
    b = torch.randn(n, k, generator=g)
    want = (a.double() @ b.double().t()).float()
    got = ops.linear_nt(a.to(dev), b.to(dev), precision=ops.PREC_TF32).cpu()
    err = float((got - want).abs().max() / want.abs().max())
    assert got.shape == (m, n) and err < 2e-3, err
    ref32 = ops.linear_nt(a.to(dev), b.to(dev), precision=ops.PREC_FP32).cpu()
    assert float((ref32 - want).abs().max() / want.abs().max()) < 1e-5


@pytest.mark.parametrize("k,m,n", [(4096, 256, 602), (1000, 128, 128), (70000, 256, 256), (513, 64, 36), (33, 8, 4)])
def test_gemm_tf32_tn_splitk(dev, k, m, n):
    """K4 weight-gradient form dW = A^T B (A [K,M], B [K,N]): tcgen05 kind::tf32 with MN-major operands and the
    K range split over CTAs.  Stated bound for tf32 operands: 2e-3 of max |C|."""
    from sgs_gnn_b200 import ops
    g = torch.Generator().manual_seed(k + m + n)
    a = torch.randn(k, m, generator=g)
    b = torch.randn(k, n, generator=g)
    want = (a.double().t() @ b.double()).float()
    got = ops.gemm_tn(a.to(dev), b.to(dev), static_b=True, precision=ops.PREC_TF32).cpu()
    err = float((got - want).abs().max() / want.abs().max())
    assert got.shape == (m, n) and err < 2e-3, err
    ref32 = ops.gemm_tn(a.to(dev), b.to(dev), precision=ops.PREC_FP32).cpu()
    assert float((ref32 - want).abs().max() / want.abs().max()) < 1e-5
    # accumulate into an existing C
    c0 = torch.ones(m, n, device=dev)
    got2 = ops.gemm(a.to(dev), 1, m, ops._rows_aligned16(b.to(dev))[0], 1, (n + 3) // 4 * 4, m, n, k, out=c0,
                    accumulate=True, precision=ops.PREC_TF32).cpu()
    assert float((got2 - want - 1.0).abs().max() / want.abs().max()) < 2e-3


def test_tf32_scorer_mode_keeps_the_fp32_bar(dev):
    """precision 'tf32': the chunked fp32 pipeline with its contractions on the tcgen05 kind::tf32 GEMM -- the
    tensor-core mode that keeps the fp32 parity bar (north_star: 1e-4 relative)."""
    from sgs_gnn_b200 import ops, synth
    b = synth.make_graph(None, seed=21, n=3000, e=80000, f=16, c=4).to(dev)
    h = 256
    g = torch.Generator(device=dev).manual_seed(5)
    out = torch.relu(torch.randn(3000, h, generator=g, device=dev)) * 0.5
    w1 = (torch.rand(h, 2 * h, generator=g, device=dev) - 0.5) * (2.0 / (2 * h) ** 0.5)
    b1 = (torch.rand(h, generator=g, device=dev) - 0.5) * 0.1
    w2 = (torch.rand(h, generator=g, device=dev) - 0.5) * (2.0 / h ** 0.5)
    b2 = torch.zeros(1, device=dev)
    graph = ops.graph_of(b.edge_index, 3000)
    dp = torch.randn(graph.num_edges, generator=g, device=dev)
    res = {}
    for prec in ("fp32", "tf32"):
        o = out.clone().requires_grad_(True)
        ws = [t_.clone().requires_grad_(True) for t_ in (w1, b1, w2, b2)]
        p = ops.edge_score(o, ws[0], ws[1], ws[2], ws[3], graph, precision=ops._PRECISION[prec])
        p.backward(dp)
        res[prec] = (p.detach(), o.grad, [t_.grad for t_ in ws])
    p32, p19 = res["fp32"][0], res["tf32"][0]
    assert float(((p32 - p19).abs() / p32.abs().clamp_min(1e-6)).max()) < 1e-4
    # gradients, L2-relative: a hidden pre-activation within ~1e-5 of zero may flip its ReLU gate between the two
    # modes (a whole dz * w2_j term appears / disappears), so the max norm is not the right yardstick
    rel = lambda a, b_: float((a - b_).norm() / (b_.norm() + 1e-30))   # noqa: E731
    # (measured r02: 1.4e-2 on d_out with ~160 flipped gates among 80 000 x 256, each worth a whole dz * w2_j term)
    assert rel(res["tf32"][1], res["fp32"][1]) < 3e-2
    for ga, gb in zip(res["tf32"][2], res["fp32"][2]):
        assert rel(ga, gb) < 3e-2
