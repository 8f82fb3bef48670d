"""CPU: the extended oracle (oracle/extended.py) against the committed golden fixtures that
tests/golden/make_golden.py produced from the REFERENCE'S OWN modules."""
import numpy as np
import torch

from conftest import FixtureBatch, load_golden, t
from oracle import extended as ox


def _unpack(bits, n):
    return torch.from_numpy(np.unpackbits(bits)[:n].astype(bool))


def test_sampler_matches_reference_masks():
    for name in ("sampler_small.npz", "sampler_mid.npz"):
        z = load_golden(name)
        p, prob, q = t(z["p"]), t(z["prob"]), int(z["q"])
        e = p.numel()
        for k, istest in (("train", False), ("test", True)):
            s = ox.sample_topq(p, prob, q, t(z[f"noise_{k}"]), 0.3, istest)
            assert torch.equal(s.mask, _unpack(z[f"mask_{k}"], e))
            assert torch.equal(s.weights, t(z[f"weights_{k}"]))
            assert s.sel.numel() == q and bool((s.sel[1:] > s.sel[:-1]).all())
        idx, _ = ox.random_draw(prob, q, t(z["noise_rand"]))
        assert torch.equal(torch.sort(idx).values, t(z["rand_idx_sorted"]))


def test_topq_tie_break_is_lowest_index():
    keys = torch.tensor([1.0, 2.0, 2.0, 2.0, 3.0, 2.0])
    sel, tau, n_gt = ox.topq_select(keys, 3)
    assert sel.tolist() == [1, 2, 4] and tau == 2.0 and n_gt == 1


def test_forward_matches_reference():
    z = load_golden("forward_small.npz")
    b = FixtureBatch(z)
    params = {k[3:]: t(v) for k, v in z.items() if k.startswith("sd.")}
    rei, w = t(z["rand_edge_index"]), t(z["w"])
    with torch.no_grad():
        p_sparse = ox.edge_prob_gcn(params, b.x, b.edge_index, rei, training=False).squeeze()
        p_full = ox.edge_prob_gcn(params, b.x, b.edge_index, None, training=False, chunk=777).squeeze()
        lw = ox.gnn_forward(params, b.x, rei, w, training=False)
        lu = ox.gnn_forward(params, b.x, rei, None, training=False)
    assert torch.allclose(p_sparse, t(z["p_sparse"]), atol=1e-6)
    assert torch.allclose(p_full, t(z["p_fullgraph"]), atol=1e-6)
    assert torch.allclose(lw, t(z["logits_weighted"]), atol=1e-5)
    assert torch.allclose(lu, t(z["logits_unweighted"]), atol=1e-5)


def test_losses_match_reference():
    z = load_golden("losses_small.npz")
    logits = t(z["logits"]).requires_grad_(True)
    p_s = t(z["p_s"]).requires_grad_(True)
    s_ei, y, tm = t(z["s_ei"]), t(z["y"]), t(z["train_mask"])
    total = ox.hybrid_loss(logits, p_s, s_ei, y, tm)
    assert abs(float(total) - float(z["total"])) < 1e-6
    gl, gp = torch.autograd.grad(total, [logits, p_s])
    assert torch.allclose(gl, t(z["grad_logits"]), atol=1e-7)
    assert torch.allclose(gp, t(z["grad_p"]), atol=1e-7)
    c, n = ox.accuracy_counts(logits.detach(), y, tm)
    assert abs(c / n - float(z["f1"])) < 1e-9
    _, n_valid, sum_label = ox.reg1_loss(p_s.detach(), s_ei, y, tm)
    assert n_valid == int(z["n_valid"]) and sum_label == float(z["sum_label"])


def test_first_step_matches_reference_training():
    for name, pipe in (("step_hybrid.npz", "hybrid"), ("step_st.npz", "straight_through"),
                       ("step_two_pass.npz", "two_pass")):
        z = load_golden(name)
        b = FixtureBatch(z)
        params = {k[4:]: t(v).clone().requires_grad_(True) for k, v in z.items() if k.startswith("sd0.")}
        noises = t(z["noises"])
        st = ox.learned_step(params, b, int(z["q"]), noises[0][0], noises[0][1], pipeline=pipe)
        assert abs(st.loss - float(z["losses"][0])) < 1e-5
        assert (st.branch == "learned") == bool(z["learned_wins"][0])
        assert torch.equal(st.sel, t(z["oracle_sel0"]))


def test_ensemble_evaluate_matches_reference():
    """evaluate.py:6-173 (learned mode): the oracle with the reference's own Exp(1) draws injected reproduces the
    reference's F1 triple, for the 3-member ensemble and for the single-member `evaluate`."""
    z = load_golden("eval_small.npz")

    class B(FixtureBatch):
        pass

    b = B(z)
    b.val_mask, b.test_mask = t(z["val_mask"]), t(z["test_mask"])
    params = {k[3:]: t(v) for k, v in z.items() if k.startswith("sd.")}
    with torch.no_grad():
        f1, mean_logits = ox.ensemble_evaluate(params, b, int(z["q"]), list(t(z["noises"])))
        f1_one, _ = ox.ensemble_evaluate(params, b, int(z["q"]), [t(z["noise_one"])])
    assert np.allclose(f1, z["f1_ensemble"]) and np.allclose(f1_one, z["f1_single"])
    assert torch.allclose(mean_logits, t(z["mean_logits"]), atol=1e-6)


def test_edge_weight_grad_formula():
    torch.manual_seed(0)
    n, m, f, d = 40, 300, 6, 5
    ei = torch.randint(0, n, (2, m))
    x = torch.randn(n, f, dtype=torch.float64)
    w = torch.rand(m, dtype=torch.float64, requires_grad=True)
    W = torch.randn(d, f, dtype=torch.float64)
    G = torch.randn(n, d, dtype=torch.float64)
    out = ox.gcn_conv(x, W, torch.zeros(d, dtype=torch.float64), ei, w)
    (ga,) = torch.autograd.grad((out * G).sum(), [w])
    keep = ei[0] != ei[1]
    closed = ox.gcn_edge_weight_grad(x, W, ei, w.detach(), G)
    assert torch.allclose(ga[keep], closed[keep], atol=1e-12)


def test_add_degree_dropin_matches_oracle():
    from types import SimpleNamespace
    from sgs_gnn_b200 import datasets, synth
    b = synth.make_graph("amazon-ratings", seed=5, scale=0.2)
    import pytest
    d = SimpleNamespace(edge_index=b.edge_index, x=b.x, num_nodes=b.num_nodes)
    if torch.cuda.is_available():      # host data: uploaded, computed by the kernels, copied back
        datasets.add_degree(d)
        want = ox.degree_prior(b.edge_index, b.num_nodes)
        assert d.prob.device.type == "cpu" and float(((d.prob - want).abs() / want).max()) < 1e-4
    else:                              # no GPU: fails loudly, there is no CPU implementation
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            datasets.add_degree(d)
    d.edge_index = b.edge_index.flip(1)
    with pytest.raises(RuntimeError, match="sorted"):
        datasets.add_degree(d)


def test_degree_prior_is_a_distribution():
    from sgs_gnn_b200 import synth
    b = synth.make_graph("smallcora", seed=3)
    assert torch.equal(ox.degree_prior(b.edge_index, b.num_nodes), b.prob)
    assert abs(float(b.prob.sum()) - 1.0) < 1e-4


def test_sage_scorer_matches_reference():
    """oracle/extended.edge_prob_sage against the reference's EdgeProbSAGE (model.py:47-89) fixture: probabilities
    and parameter gradients, full-graph and sparse message passing."""
    z = load_golden("sage_small.npz")
    x, ei, rei, gup = t(z["x"]), t(z["edge_index"]), t(z["rand_edge_index"]), t(z["gup"])
    for tag, sub in (("full", None), ("sparse", rei)):
        params = {k[3:]: t(v).clone().requires_grad_(k.startswith("sd.edge_prob_mlp.")) for k, v in z.items()
                  if k.startswith("sd.")}
        p = ox.edge_prob_sage(params, x, ei, sub, training=False).squeeze()
        assert torch.allclose(p.detach(), t(z["p_" + tag]), atol=1e-6)
        (p * gup).sum().backward()
        for k, v in params.items():
            if k.startswith("edge_prob_mlp."):
                want = t(z[f"grad_{tag}.{k[len('edge_prob_mlp.'):]}"])
                assert torch.allclose(v.grad, want, rtol=1e-4, atol=1e-6), (tag, k)


def test_mlp_scorer_matches_reference():
    """oracle/extended.edge_prob_mlp against the reference's EdgeProbMLP (model.py:8-45) fixture."""
    z = load_golden("mlp_small.npz")
    x, ei, gup = t(z["x"]), t(z["edge_index"]), t(z["gup"])
    params = {k[3:]: t(v).clone().requires_grad_(k.startswith("sd.edge_prob_mlp.")) for k, v in z.items()
              if k.startswith("sd.")}
    p = ox.edge_prob_mlp(params, x, ei, training=False).squeeze()
    assert torch.allclose(p.detach(), t(z["p_full"]), atol=1e-6)
    (p * gup).sum().backward()
    for k, v in params.items():
        if k.startswith("edge_prob_mlp."):
            assert torch.allclose(v.grad, t(z[f"grad.{k[len('edge_prob_mlp.'):]}"]), rtol=1e-4, atol=1e-6), k
