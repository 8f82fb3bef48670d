"""Generates tests/golden/*.npz by running the REFERENCE'S OWN modules (imported
unmodified from /root/reference through oracle/shim) on seeded synthetic inputs.

Run in the builder container only:   python tests/golden/make_golden.py
The fixtures are committed; nothing on the GPU box reads /root/reference.

What the fixtures pin (SURVEY.md section 8c -- the reference ships no golden vectors):
  sampler_*.npz   sampling.gumbel_softmax_sampling mask/weights for given (p, prob)
                  with the Exp(1) noise torch.multinomial drew (train and istest).
  scorer_*.npz    model.EdgeProbGCN forward probabilities (eval mode => no dropout).
  sage_*.npz      model.EdgeProbSAGE forward probabilities + parameter gradients (eval mode).
  mlp_*.npz       model.EdgeProbMLP forward probabilities + parameter gradients (eval mode).
  gnn_*.npz       model.GNNModel forward logits, weighted and unweighted.
  losses_*.npz    utils.consistency_loss, reg1 BCE block, CE -- values + grads.
  step_*.npz      training_hybrid.train / training_straight_through.train over
                  several epochs with drop_rate 0: per-epoch loss, branch counters,
                  the injected noise tensors, initial and final state_dict.
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader, extended as ox  # noqa: E402
from sgs_gnn_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ref = ref_loader.load()


def npz(name, **kw):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in kw.items()})
    print(f"wrote {name}: {os.path.getsize(path) / 1024:.1f} KiB")


def make_args(**kw):
    a = dict(device="cpu", mode="learned", hybrid_checkpoint=False, conditional=True,
             sparse_edge_mlp=True, t_init=0.7, t_min=0.5, degree_bias_coef=0.3, reg1=True,
             reg2=True, regularizer1_coef=1.0, consist_reg_coef=0.5, pipeline="hybrid")
    a.update(kw)
    return SimpleNamespace(**a)


def build_ref_model(f, h, c, drop, seed):
    torch.manual_seed(seed)
    model = ref.model.GNNModel(f, h, c, drop, "GCN")
    opt_gnn = torch.optim.Adam([p for n, p in model.named_parameters() if "gcn" in n], lr=1e-3)
    opt_edge = torch.optim.Adam([p for n, p in model.named_parameters() if "edge_prob_mlp" in n], lr=1e-3)
    opt_all = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-4)
    return model, opt_gnn, opt_edge, opt_all


def golden_sampler():
    for tag, shape, scale, seed in (("small", "smallcora", 1.0, 3), ("mid", "arxiv-year", 0.2, 4)):
        b = synth.make_graph(shape, seed=seed, scale=scale, f=8)
        e = b.num_edges
        g = torch.Generator().manual_seed(seed)
        p = torch.sigmoid(torch.randn(e, generator=g))
        q = int(e * 0.2)
        out = {"p": p, "prob": b.prob, "q": q}
        for istest in (False, True):
            torch.manual_seed(100 + seed)
            mask, w = ref.sampling.gumbel_softmax_sampling(b, p, b.edge_index, q=q, degree_bias_coef=0.3,
                                                           istest=istest)
            torch.manual_seed(100 + seed)
            noise = ox.exponential_noise(e)
            s = ox.sample_topq(p, b.prob, q, noise, 0.3, istest)
            assert torch.equal(mask, s.mask) and torch.equal(w, s.weights)
            k = "test" if istest else "train"
            out[f"noise_{k}"] = noise
            out[f"mask_{k}"] = np.packbits(mask.numpy())
            out[f"weights_{k}"] = w
            out[f"S_{k}"] = p.sum()
        # random-baseline draw (training_hybrid.py:45-48)
        torch.manual_seed(200 + seed)
        idx = torch.multinomial(torch.softmax(b.prob, -1), q, replacement=False)
        torch.manual_seed(200 + seed)
        noise = ox.exponential_noise(e)
        idx2, _ = ox.random_draw(b.prob, q, noise)
        assert torch.equal(torch.sort(idx).values, torch.sort(idx2).values)
        out["noise_rand"] = noise
        out["rand_idx_sorted"] = torch.sort(idx).values
        npz(f"sampler_{tag}.npz", **out)


def golden_forward():
    b = synth.make_graph(None, seed=11, n=400, e=3200, f=24, c=5)
    h = 32
    model, *_ = build_ref_model(24, h, 5, 0.3, 21)
    model.eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    torch.manual_seed(5)
    q = 640
    ridx = torch.multinomial(torch.softmax(b.prob, -1), q, replacement=False)
    rei = b.edge_index[:, ridx]
    with torch.no_grad():
        p_sparse = model.edge_prob_mlp(b.x, b.edge_index, rei).squeeze()
        p_fullg = model.edge_prob_mlp(b.x, b.edge_index, None).squeeze()
        w = torch.rand(q)
        logits_w = model(b, rei, w)
        logits_u = model(b, rei)
    params = {k: v for k, v in sd.items()}
    with torch.no_grad():
        o1 = ox.edge_prob_gcn(params, b.x, b.edge_index, rei, training=False).squeeze()
        o2 = ox.gnn_forward(params, b.x, rei, w, training=False)
    assert torch.allclose(o1, p_sparse, atol=1e-6) and torch.allclose(o2, logits_w, atol=1e-5)
    npz("forward_small.npz", x=b.x, y=b.y, edge_index=b.edge_index, train_mask=b.train_mask, prob=b.prob,
        rand_edge_index=rei, w=w, p_sparse=p_sparse, p_fullgraph=p_fullg, logits_weighted=logits_w,
        logits_unweighted=logits_u, **{"sd." + k: v for k, v in sd.items()})

    # losses + grads (training_hybrid.py:103-132)
    torch.manual_seed(9)
    logits = torch.randn(400, 5, requires_grad=True)
    p_s = torch.rand(q).clamp(0.01, 0.99).requires_grad_(True)
    crit = nn.CrossEntropyLoss()
    ce = crit(logits[b.train_mask], b.y[b.train_mask])
    labels = torch.full((q,), -1, dtype=torch.long)
    tr_idx = torch.nonzero(b.train_mask).squeeze()
    src, dst = rei[0], rei[1]
    tem = torch.isin(src, tr_idx) & torch.isin(dst, tr_idx)
    same = b.y[src] == b.y[dst]
    labels[tem & same] = 1
    labels[tem & ~same] = 0
    valid = labels != -1
    assert labels[valid].sum().item() > 1
    bce = torch.nn.functional.binary_cross_entropy(p_s[valid], labels[valid].float())
    cons = ref.utils.consistency_loss(p_s, rei, logits)
    total = ce + 1.0 * bce + 0.5 * cons
    gl, gp = torch.autograd.grad(total, [logits, p_s])
    lc = ref.utils.calculate_f1(logits.detach(), b.y, b.train_mask)
    npz("losses_small.npz", logits=logits, p_s=p_s, s_ei=rei, y=b.y, train_mask=b.train_mask, ce=ce, bce=bce,
        cons=cons, total=total, grad_logits=gl, grad_p=gp, f1=lc, n_valid=int(valid.sum()),
        sum_label=float(labels[valid].sum()))


def golden_sage():
    """model.EdgeProbSAGE (model.py:47-89) forward probabilities + parameter gradients, eval-mode dropout, on the full
    graph and with a random message-passing subgraph; the reference class runs on the shim's SAGEConv."""
    b = synth.make_graph(None, seed=13, n=350, e=2800, f=20, c=4)
    h = 32
    torch.manual_seed(23)
    model = ref.model.GNNModel(20, h, 4, 0.3, "GSAGE")
    model.eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    assert [k for k in sd if k.startswith("edge_prob_mlp.gcn1")] == [
        "edge_prob_mlp.gcn1.lin_l.weight", "edge_prob_mlp.gcn1.lin_l.bias", "edge_prob_mlp.gcn1.lin_r.weight"]
    torch.manual_seed(6)
    ridx = torch.multinomial(torch.softmax(b.prob, -1), 560, replacement=False)
    rei = b.edge_index[:, ridx]
    gup = torch.randn(b.edge_index.size(1))
    out = {}
    for tag, sub in (("full", None), ("sparse", rei)):
        model.zero_grad()
        p = model.edge_prob_mlp(b.x, b.edge_index, sub).squeeze()
        (p * gup).sum().backward()
        out["p_" + tag] = p.detach()
        for k, v in model.edge_prob_mlp.named_parameters():
            out[f"grad_{tag}.{k}"] = v.grad.clone()
        with torch.no_grad():
            o = ox.edge_prob_sage(sd, b.x, b.edge_index, sub, training=False).squeeze()
        assert torch.allclose(o, p.detach(), atol=1e-6)
    npz("sage_small.npz", x=b.x, edge_index=b.edge_index, prob=b.prob, rand_edge_index=rei, gup=gup, hidden=h, **out,
        **{"sd." + k: v for k, v in sd.items()})


def golden_mlp():
    """model.EdgeProbMLP (model.py:8-45) with random_sampled_edge_index=None (the only shape-consistent form, SURVEY
    a3): forward probabilities + parameter gradients, eval-mode dropout."""
    b = synth.make_graph(None, seed=14, n=330, e=2600, f=18, c=4)
    h = 32
    torch.manual_seed(24)
    model = ref.model.GNNModel(18, h, 4, 0.3, "MLP")
    model.eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    gup = torch.randn(b.edge_index.size(1))
    p = model.edge_prob_mlp(b.x, b.edge_index, None).squeeze()
    (p * gup).sum().backward()
    out = {"p_full": p.detach()}
    for k, v in model.edge_prob_mlp.named_parameters():
        out[f"grad.{k}"] = v.grad.clone()
    with torch.no_grad():
        o = ox.edge_prob_mlp(sd, b.x, b.edge_index, training=False).squeeze()
    assert torch.allclose(o, p.detach(), atol=1e-6)
    npz("mlp_small.npz", x=b.x, edge_index=b.edge_index, gup=gup, hidden=h, **out,
        **{"sd." + k: v for k, v in sd.items()})


def golden_step(pipeline, tag, n, e, f, c, h, epochs, seed, record=False):
    """record: also store, per epoch, what the reference's training loop handed to / got back from
    gumbel_softmax_sampling (the full edge probabilities and the selected-edge mask), so that reduced-precision
    modes can report per-epoch probability error and sampled-set overlap against the reference."""
    b = synth.make_graph(None, seed=seed, n=n, e=e, f=f, c=c, homophily=0.7)
    q = int(e * 0.2)
    model, opt_gnn, opt_edge, opt_all = build_ref_model(f, h, c, 0.0, seed + 1)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    args = make_args(pipeline=pipeline)
    crit = nn.CrossEntropyLoss()
    mod = {"hybrid": ref.training_hybrid, "straight_through": ref.training_straight_through,
           "two_pass": ref.training_two_pass}[pipeline]
    losses, branch, noises = [], [], []
    rec_p, rec_mask = [], []
    if record:
        inner = mod.gumbel_softmax_sampling

        def spy(batch, edge_probs, *a, **kw):
            mask, w = inner(batch, edge_probs, *a, **kw)
            rec_p.append(edge_probs.detach().clone())
            rec_mask.append(np.packbits(mask.numpy()))
            return mask, w
        mod.gumbel_softmax_sampling = spy
    for ep in range(epochs):
        torch.manual_seed(1000 + ep)
        noises.append(torch.stack([ox.exponential_noise(e), ox.exponential_noise(e)]))
        torch.manual_seed(1000 + ep)
        loss, _, n_cond, n_tot = mod.train(args, ep, epochs, model, opt_gnn, opt_edge, opt_all, crit, [b],
                                           q=q, alternate_frequency=0)
        losses.append(loss)
        branch.append(n_cond)
    sd1 = model.state_dict()
    extra = {}
    if record:
        mod.gumbel_softmax_sampling = inner
        extra = {"ref_p_full": torch.stack(rec_p), "ref_masks": np.stack(rec_mask)}
    # cross-check the extended oracle on the first step (same noise, dropout off)
    params = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    st = ox.learned_step(params, b, q, noises[0][0], noises[0][1], pipeline=pipeline)
    assert abs(st.loss - losses[0]) < 1e-5 * max(1, abs(losses[0])), (st.loss, losses[0])
    assert (st.branch == "learned") == bool(branch[0])
    print(pipeline, "losses", [round(x, 5) for x in losses], "learned-wins", branch)
    npz(f"step_{tag}.npz", x=b.x, y=b.y, edge_index=b.edge_index, train_mask=b.train_mask, prob=b.prob, q=q,
        hidden=h, noises=torch.stack(noises), losses=np.array(losses), learned_wins=np.array(branch),
        oracle_sel0=st.sel, oracle_pfull0=st.p_full, **extra,
        **{"sd0." + k: v for k, v in sd0.items()}, **{"sd1." + k: v for k, v in sd1.items()})


def golden_baseline_modes():
    """training_hybrid.train in the baseline modes `full` and `edge` (training_hybrid.py:149-180) with `optimizer`
    = Adam(lr 1e-3, weight_decay 5e-4) over all parameters (main.py:123): 6-epoch trajectories.  `edge` draws
    torch.multinomial(softmax(prob), q) once per epoch; the Exp(1) tensors it drew are recorded."""
    n, e, f, c, h, epochs = 300, 2400, 20, 4, 32, 6
    for mode, seed in (("full", 91), ("edge", 93)):
        b = synth.make_graph(None, seed=seed, n=n, e=e, f=f, c=c, homophily=0.7)
        q = int(e * 0.2)
        model, opt_gnn, opt_edge, opt_all = build_ref_model(f, h, c, 0.0, seed + 1)
        sd0 = {k: v.clone() for k, v in model.state_dict().items()}
        args = make_args(mode=mode)
        crit = nn.CrossEntropyLoss()
        losses, noises = [], []
        for ep in range(epochs):
            torch.manual_seed(2000 + ep)
            noises.append(ox.exponential_noise(e))
            torch.manual_seed(2000 + ep)
            loss, _, n_cond, n_tot = ref.training_hybrid.train(args, ep, epochs, model, opt_gnn, opt_edge, opt_all, crit,
                                                               [b], q=q, alternate_frequency=0)
            losses.append(loss)
        print(mode, "losses", [round(x, 5) for x in losses])
        npz(f"step_mode_{mode}.npz", x=b.x, y=b.y, edge_index=b.edge_index, train_mask=b.train_mask, prob=b.prob, q=q,
            hidden=h, noises=torch.stack(noises), losses=np.array(losses),
            **{"sd0." + k: v for k, v in sd0.items()}, **{"sd1." + k: v for k, v in model.state_dict().items()})


def golden_eval():
    """evaluate.evaluate / evaluate.ensemble_evaluate (learned mode) on a model in eval mode; the Exp(1) noise
    torch.multinomial draws for each ensemble member is re-generated from the same seed and stored."""
    n, e, f, c, h, members = 500, 5000, 16, 4, 32, 3
    b = synth.make_graph(None, seed=51, n=n, e=e, f=f, c=c, homophily=0.7)
    q = int(e * 0.2)
    model, *_ = build_ref_model(f, h, c, 0.3, 52)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    args = make_args(num_samples_eval=members)
    torch.manual_seed(77)
    f1_ens = ref.evaluate.ensemble_evaluate(args, model, [b], torch.device("cpu"), q=q, mode="learned")
    torch.manual_seed(77)
    noises = torch.stack([ox.exponential_noise(e) for _ in range(members)])
    torch.manual_seed(78)
    f1_one = ref.evaluate.evaluate(args, model, [b], torch.device("cpu"), q=q, mode="learned")
    torch.manual_seed(78)
    noise_one = ox.exponential_noise(e)
    params = {k: v for k, v in sd.items()}
    with torch.no_grad():
        f1_o, mean_logits = ox.ensemble_evaluate(params, b, q, list(noises))
        f1_o1, logits_one = ox.ensemble_evaluate(params, b, q, [noise_one])
    assert np.allclose(f1_o, f1_ens) and np.allclose(f1_o1, f1_one), (f1_o, f1_ens, f1_o1, f1_one)
    print("eval f1 (ensemble, single):", f1_ens, f1_one)
    npz("eval_small.npz", x=b.x, y=b.y, edge_index=b.edge_index, train_mask=b.train_mask, val_mask=b.val_mask,
        test_mask=b.test_mask, prob=b.prob, q=q, hidden=h, members=members, noises=noises, noise_one=noise_one,
        f1_ensemble=np.array(f1_ens), f1_single=np.array(f1_one), mean_logits=mean_logits, logits_single=logits_one,
        **{"sd." + k: v for k, v in sd.items()})


if __name__ == "__main__":
    torch.set_num_threads(4)
    which = sys.argv[1:] or ["sampler", "forward", "hybrid", "st", "two_pass", "eval", "sage", "mlp", "hybrid_h256",
                               "st_h256", "modes"]
    if "mlp" in which:
        golden_mlp()
    if "sage" in which:
        golden_sage()
    if "sampler" in which:
        golden_sampler()
    if "forward" in which:
        golden_forward()
    if "hybrid" in which:
        golden_step("hybrid", "hybrid", 300, 2400, 20, 4, 32, 6, 31)
    if "st" in which:
        golden_step("straight_through", "st", 300, 2400, 20, 4, 32, 6, 41)
    if "two_pass" in which:
        golden_step("two_pass", "two_pass", 300, 2400, 20, 4, 32, 6, 61)
    if "hybrid_h256" in which:     # the benched tensor-core widths (H = 256): reduced-precision trajectory tests
        golden_step("hybrid", "hybrid_h256", 1200, 20000, 48, 5, 256, 6, 71, record=True)
    if "st_h256" in which:
        golden_step("straight_through", "st_h256", 1200, 20000, 48, 5, 256, 6, 81, record=True)
    if "modes" in which:
        golden_baseline_modes()
    if "eval" in which:
        golden_eval()
