"""CPU, builder container only: the extended oracle against the reference's own modules run live
(skipped where /root/reference does not exist, e.g. on the GPU box)."""
import pytest
import torch

from oracle import extended as ox
from oracle import ref_loader
from sgs_gnn_b200 import synth

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference not present")


def test_multinomial_is_topk_of_p_over_exponential():
    ref = ref_loader.load()
    b = synth.make_graph(None, seed=2, n=200, e=3000, f=4, c=3)
    p = torch.rand(3000)
    for istest in (False, True):
        torch.manual_seed(11)
        mask, w = ref.sampling.gumbel_softmax_sampling(b, p, b.edge_index, q=700, istest=istest)
        torch.manual_seed(11)
        s = ox.sample_topq(p, b.prob, 700, ox.exponential_noise(3000), 0.3, istest)
        assert torch.equal(mask, s.mask) and torch.equal(w, s.weights)


def test_state_dict_keys_and_optimizer_groups():
    ref = ref_loader.load()
    from sgs_gnn_b200.model import GNNModel
    torch.manual_seed(3)
    a = ref.model.GNNModel(12, 16, 3, 0.3, "GCN")
    torch.manual_seed(3)
    b = GNNModel(12, 16, 3, 0.3, "GCN")
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert sa[k].shape == sb[k].shape and torch.equal(sa[k], sb[k]), k
    gcn_a = [n for n, _ in a.named_parameters() if "gcn" in n]
    gcn_b = [n for n, _ in b.named_parameters() if "gcn" in n]
    assert gcn_a == gcn_b and any(n.startswith("edge_prob_mlp.gcn") for n in gcn_b)


def test_oracle_step_with_dropout_uses_same_rng_stream():
    ref = ref_loader.load()
    import torch.nn as nn
    from types import SimpleNamespace
    b = synth.make_graph(None, seed=8, n=150, e=1200, f=10, c=3)
    torch.manual_seed(4)
    model = ref.model.GNNModel(10, 16, 3, 0.3, "GCN")
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    args = SimpleNamespace(device="cpu", mode="learned", hybrid_checkpoint=False, conditional=True,
                           sparse_edge_mlp=True, t_init=0.7, t_min=0.5, degree_bias_coef=0.3, reg1=True, reg2=True,
                           regularizer1_coef=1.0, consist_reg_coef=0.5)
    og = torch.optim.Adam([p for n, p in model.named_parameters() if "gcn" in n], lr=1e-3)
    oe = torch.optim.Adam([p for n, p in model.named_parameters() if "edge_prob_mlp" in n], lr=1e-3)
    oa = torch.optim.Adam(model.parameters(), lr=1e-3)
    torch.manual_seed(99)
    loss, *_ = ref.training_hybrid.train(args, 1, 10, model, og, oe, oa, nn.CrossEntropyLoss(), [b], q=240,
                                         alternate_frequency=0)
    # oracle: same generator, masks drawn by F.dropout in the same order
    torch.manual_seed(99)
    params = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    noise_r = ox.exponential_noise(1200)
    # replay: the oracle draws dropout masks itself between the two multinomial draws
    x, ei = b.x, b.edge_index
    ridx, _ = ox.random_draw(b.prob, 240, noise_r)
    p_full = ox.edge_prob_gcn(params, x, ei, ei[:, ridx], 0.3, None, True).squeeze(-1)
    noise_l = ox.exponential_noise(1200)
    smp = ox.sample_topq(p_full.detach(), b.prob, 240, noise_l)
    lo = ox.gnn_forward(params, x, ei[:, smp.sel], p_full[smp.sel], 0.3, None, True)
    ro = ox.gnn_forward(params, x, ei[:, ridx], None, 0.3, None, True)
    lc, rc = ox.accuracy_counts(lo, b.y, b.train_mask), ox.accuracy_counts(ro, b.y, b.train_mask)
    if lc[0] > rc[0]:
        l2 = ox.hybrid_loss(lo, p_full[smp.sel], ei[:, smp.sel], b.y, b.train_mask)
    else:
        l2 = torch.nn.functional.cross_entropy(ro[b.train_mask], b.y[b.train_mask])
    assert abs(float(l2) - loss) < 1e-5


def test_sage_state_dict_and_forward_live():
    """EdgeProbSAGE: same state_dict keys / initial values as the reference under the same seed (the 'gcn' name filter
    of main.py:100 therefore selects the same tensors), and the oracle restatement equals the reference forward."""
    ref = ref_loader.load()
    from sgs_gnn_b200.model import GNNModel
    torch.manual_seed(4)
    a = ref.model.GNNModel(10, 16, 3, 0.3, "GSAGE")
    torch.manual_seed(4)
    b = GNNModel(10, 16, 3, 0.3, "GSAGE")
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert [n for n, _ in a.named_parameters() if "gcn" in n] == [n for n, _ in b.named_parameters() if "gcn" in n]
    g = synth.make_graph(None, seed=3, n=120, e=900, f=10, c=3)
    a.eval()
    with torch.no_grad():
        want = a.edge_prob_mlp(g.x, g.edge_index, g.edge_index[:, ::5])
        got = ox.edge_prob_sage(dict(sa), g.x, g.edge_index, g.edge_index[:, ::5], training=False)
    assert torch.allclose(got, want, atol=1e-6)
