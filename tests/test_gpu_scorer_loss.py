"""GPU parity: K1/K1b edge scorer and K5 fused losses through the C ABI vs oracle + golden fixtures."""
import numpy as np
import pytest
import torch

from conftest import load_golden, t
from oracle import extended as ox

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def relerr(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def test_golden_scorer_probabilities(dev):
    from sgs_gnn_b200.model import GNNModel
    z = load_golden("forward_small.npz")
    sd = {k[3:]: t(v) for k, v in z.items() if k.startswith("sd.")}
    model = GNNModel(24, 32, 5, 0.3, "GCN")
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    x, ei, rei = t(z["x"], dev), t(z["edge_index"], dev), t(z["rand_edge_index"], dev)
    with torch.no_grad():
        ps = model.edge_prob_mlp(x, ei, rei)
        pf = model.edge_prob_mlp(x, ei, None, use_checkpoint=True)
    assert ps.shape == (ei.size(1), 1)
    assert relerr(ps.squeeze().cpu(), t(z["p_sparse"])) < RTOL
    assert relerr(pf.squeeze().cpu(), t(z["p_fullgraph"])) < RTOL


@pytest.mark.parametrize("h,p_drop,subset", [(32, 0.0, False), (256, 0.0, True), (64, 0.3, False), (128, 0.3, True)])
def test_scorer_forward_backward_vs_oracle(dev, h, p_drop, subset):
    from sgs_gnn_b200 import ops, rng
    n, e = 300, 4000
    g = torch.Generator().manual_seed(h)
    ei = torch.randint(0, n, (2, e), generator=g)
    out = torch.relu(torch.randn(n, h, generator=g))
    W1 = (torch.rand(h, 2 * h, generator=g) - 0.5) * (2 / (2 * h) ** 0.5)
    b1 = (torch.rand(h, generator=g) - 0.5) * 0.1
    w2 = (torch.rand(1, h, generator=g) - 0.5) * (2 / h ** 0.5)
    b2 = torch.tensor([0.05])
    ids = torch.sort(torch.randperm(e, generator=g)[:900]).values if subset else torch.arange(e)
    seed = 99
    keep = torch.from_numpy(rng.keep_mask(seed, ids.numpy(), h, p_drop)) if p_drop > 0 else None
    gup = torch.randn(ids.numel(), generator=g)

    ro, rW1, rb1, rw2, rb2 = (v.clone().requires_grad_(True) for v in (out, W1, b1, w2, b2))
    p_ref = ox.edge_score(ro, ei[:, ids], rW1, rb1, rw2, rb2, p_drop, keep, True).squeeze(-1)
    g_ref = torch.autograd.grad((p_ref * gup).sum(), [ro, rW1, rb1, rw2, rb2])

    graph = ops.graph_of(ei.to(dev), n)
    do, dW1, db1, dw2, db2 = (v.to(dev).requires_grad_(True) for v in (out, W1, b1, w2, b2))
    ids_d = ids.to(dev).int() if subset else None
    p = ops.edge_score(do, dW1, db1, dw2, db2, graph, ids_d, p_drop, seed)
    assert relerr(p.detach().cpu(), p_ref.detach()) < RTOL
    grads = torch.autograd.grad((p * gup.to(dev)).sum(), [do, dW1, db1, dw2, db2])
    for name, a, r in zip(("d_out", "dW1", "db1", "dw2", "db2"), grads, g_ref):
        assert relerr(a.cpu().reshape(r.shape), r) < RTOL, name


def test_golden_losses_and_grads(dev):
    from sgs_gnn_b200 import ops, utils
    z = load_golden("losses_small.npz")
    logits = t(z["logits"], dev).requires_grad_(True)
    p_s = t(z["p_s"], dev).requires_grad_(True)
    s_ei, y, tm = t(z["s_ei"], dev), t(z["y"], dev), t(z["train_mask"], dev)
    sub = ops.graph_of(s_ei, logits.size(0))
    loss = ops.fused_loss(logits, y, tm.view(torch.uint8), p_s, sub, 1.0, 0.5, True, True)
    assert abs(loss.item() - float(z["total"])) < 1e-5 * abs(float(z["total"]))
    gl, gp = torch.autograd.grad(loss, [logits, p_s])
    assert relerr(gl.cpu(), t(z["grad_logits"])) < RTOL
    assert relerr(gp.cpu(), t(z["grad_p"])) < RTOL
    acc = ops.loss_forward(logits.detach(), y, tm.view(torch.uint8), sub, p_s.detach()).cpu()
    assert int(acc[4]) == int(z["n_valid"]) and float(acc[5]) == float(z["sum_label"])
    assert abs(float(acc[0] / acc[1]) - float(z["ce"])) < 1e-5
    assert abs(float(acc[3] / acc[4]) - float(z["bce"])) < 1e-5
    assert abs(float(acc[6] / acc[7]) - float(z["cons"])) < 1e-5
    assert abs(utils.calculate_f1(logits.detach(), y, tm) - float(z["f1"])) < 1e-9
    # stand-alone consistency_loss (utils.py:187-211) with gradients to both arguments
    l2 = utils.consistency_loss(p_s, s_ei, logits)
    assert abs(l2.item() - float(z["cons"])) < 1e-6
    ref_l = t(z["logits"]).requires_grad_(True)
    ref_p = t(z["p_s"]).requires_grad_(True)
    rg = torch.autograd.grad(ox.consistency_loss(ref_p, t(z["s_ei"]), ref_l), [ref_l, ref_p])
    og = torch.autograd.grad(l2, [logits, p_s])
    assert relerr(og[0].cpu(), rg[0]) < RTOL and relerr(og[1].cpu(), rg[1]) < RTOL


def test_reg1_skipped_when_at_most_one_positive_label(dev):
    from sgs_gnn_b200 import ops
    n, c = 10, 3
    logits = torch.randn(n, c, device=dev, requires_grad=True)
    y = torch.arange(n, device=dev) % c
    tm = torch.ones(n, dtype=torch.bool, device=dev)
    s_ei = torch.tensor([[0, 1, 2], [3, 5, 4]], device=dev)     # labels: same(0,3), diff, diff -> sum = 1
    p_s = torch.tensor([0.3, 0.6, 0.8], device=dev, requires_grad=True)
    loss = ops.fused_loss(logits, y, tm.view(torch.uint8), p_s, ops.graph_of(s_ei, n), 1.0, 0.5, True, True)
    want = ox.hybrid_loss(logits.detach().cpu(), p_s.detach().cpu(), s_ei.cpu(), y.cpu(), tm.cpu())
    assert abs(loss.item() - float(want)) < 1e-5


@pytest.mark.parametrize("c,q,n", [(3, 1, 40), (16, 15, 50), (17, 16, 64), (33, 1000, 300), (41, 4097, 700), (49, 333, 90),
                                   (64, 2049, 400), (70, 500, 100)])
def test_fused_edge_loss_class_counts_vs_oracle(dev, c, q, n):
    """One-sweep edge loss (sgs_loss_fwd_fused / _bwd_fused; C <= 64) and the two-kernel path (C = 70) against the
    fp64 torch restatement of training_hybrid.py:103-132, for every columns-per-lane template and ragged q."""
    from sgs_gnn_b200 import ops
    g = torch.Generator().manual_seed(c * 1000 + q)
    logits_c = torch.randn(n, c, generator=g)
    y_c = torch.randint(0, c, (n,), generator=g)
    tm_c = torch.rand(n, generator=g) < 0.7
    src = torch.sort(torch.randint(0, n, (q,), generator=g)).values          # ascending sources, like a sampled list
    dst = torch.randint(0, n, (q,), generator=g)
    s_ei_c = torch.stack([src, dst])
    p_c = torch.rand(q, generator=g) * 0.98 + 0.01
    ref_l = logits_c.double().requires_grad_(True)
    ref_p = p_c.double().requires_grad_(True)
    want = ox.hybrid_loss(ref_l, ref_p, s_ei_c, y_c, tm_c, True, True, 0.7, 0.4)
    rg = torch.autograd.grad(want, [ref_l, ref_p])
    logits = logits_c.to(dev).requires_grad_(True)
    p_s = p_c.to(dev).requires_grad_(True)
    loss = ops.fused_loss(logits, y_c.to(dev), tm_c.to(dev).view(torch.uint8), p_s, ops.graph_of(s_ei_c.to(dev), n),
                          0.7, 0.4, True, True)
    assert abs(loss.item() - float(want.detach())) <= 2e-5 * max(1.0, abs(float(want.detach())))
    gl, gp = torch.autograd.grad(loss, [logits, p_s])
    assert relerr(gl.cpu().double(), rg[0]) < RTOL
    assert relerr(gp.cpu().double(), rg[1]) < RTOL


def test_golden_sage_scorer(dev):
    """EdgeProbSAGE drop-in (mean-aggregation SpMM + root term inside the fused epilogue) against the fixture produced
    by the reference's EdgeProbSAGE: probabilities and all parameter gradients."""
    from sgs_gnn_b200.model import GNNModel
    z = load_golden("sage_small.npz")
    sd = {k[3:]: t(v) for k, v in z.items() if k.startswith("sd.")}
    model = GNNModel(20, int(z["hidden"]), 4, 0.3, "GSAGE")
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    x, ei, rei, gup = t(z["x"], dev), t(z["edge_index"], dev), t(z["rand_edge_index"], dev), t(z["gup"], dev)
    for tag, sub in (("full", None), ("sparse", rei)):
        model.zero_grad()
        p = model.edge_prob_mlp(x, ei, sub).squeeze()
        assert relerr(p.detach().cpu(), t(z["p_" + tag])) < RTOL
        (p * gup).sum().backward()
        for k, v in model.edge_prob_mlp.named_parameters():
            assert relerr(v.grad.cpu(), t(z[f"grad_{tag}.{k}"])) < 5 * RTOL, (tag, k)


def test_hybrid_epoch_with_sage_scorer_trains(dev):
    """training_hybrid.train end to end with --edge_mlp_type GSAGE (dropout on): finite decreasing loss."""
    from types import SimpleNamespace
    import torch.nn as nn
    from sgs_gnn_b200 import synth, training_hybrid
    from sgs_gnn_b200.model import GNNModel
    torch.manual_seed(0)
    b = synth.make_graph(None, seed=17, n=600, e=7000, f=24, c=4).to(dev)
    model = GNNModel(24, 64, 4, 0.3, "GSAGE").to(dev)
    og = torch.optim.Adam([p for n, p in model.named_parameters() if "gcn" in n], lr=1e-2)
    oe = torch.optim.Adam([p for n, p in model.named_parameters() if "edge_prob_mlp" in n], lr=1e-2)
    oa = torch.optim.Adam(model.parameters(), lr=1e-2)
    args = SimpleNamespace(device=dev, mode="learned", hybrid_checkpoint=False, conditional=False, sparse_edge_mlp=True,
                           t_init=0.7, t_min=0.5, degree_bias_coef=0.3, reg1=True, reg2=True, regularizer1_coef=1.0,
                           consist_reg_coef=0.5, pipeline="hybrid")
    losses = [training_hybrid.train(args, ep, 20, model, og, oe, oa, nn.CrossEntropyLoss(), [b], q=1400,
                                    alternate_frequency=0)[0] for ep in range(1, 9)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]


def test_golden_mlp_scorer(dev):
    """EdgeProbMLP drop-in (per-node relu(fcdim(X)) then the fused edge scorer; identical to the reference's per-edge
    projection whenever dropout is off) against the fixture produced by the reference class; the
    random_sampled_edge_index form, whose [q] output the reference itself cannot consume, raises."""
    from sgs_gnn_b200.model import GNNModel
    z = load_golden("mlp_small.npz")
    sd = {k[3:]: t(v) for k, v in z.items() if k.startswith("sd.")}
    model = GNNModel(18, int(z["hidden"]), 4, 0.3, "MLP")
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    x, ei, gup = t(z["x"], dev), t(z["edge_index"], dev), t(z["gup"], dev)
    p = model.edge_prob_mlp(x, ei, None).squeeze()
    assert relerr(p.detach().cpu(), t(z["p_full"])) < RTOL
    (p * gup).sum().backward()
    for k, v in model.edge_prob_mlp.named_parameters():
        assert relerr(v.grad.cpu(), t(z[f"grad.{k}"])) < 5 * RTOL, k
    with pytest.raises(RuntimeError):
        model.edge_prob_mlp(x, ei, ei[:, :10])
