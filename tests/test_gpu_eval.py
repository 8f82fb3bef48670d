"""GPU parity of the inference path: the drop-in evaluate / ensemble_evaluate against the
reference's own evaluate.py run on the CPU (fixture eval_small.npz: F1 triples, mean ensemble logits, and the
Exp(1) noise torch.multinomial drew for every ensemble member, injected here)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import FixtureBatch, load_golden, t

pytestmark = pytest.mark.gpu


def _setup(dev):
    from sgs_gnn_b200.model import GNNModel
    z = load_golden("eval_small.npz")
    b = FixtureBatch(z, dev)
    b.val_mask, b.test_mask = t(z["val_mask"], dev), t(z["test_mask"], dev)
    f, c, h = b.x.size(1), int(b.y.max()) + 1, int(z["hidden"])
    model = GNNModel(f, h, c, 0.3, "GCN")
    model.load_state_dict({k[3:]: t(v) for k, v in z.items() if k.startswith("sd.")})
    args = SimpleNamespace(degree_bias_coef=0.3, num_samples_eval=int(z["members"]))
    return z, b, model.to(dev), args


def test_ensemble_evaluate_matches_reference(dev):
    from sgs_gnn_b200 import evaluate, sampling
    z, b, model, args = _setup(dev)
    n_masks = [int(b.train_mask.sum()), int(b.val_mask.sum()), int(b.test_mask.sum())]
    sampling.clear_injected()
    sampling.inject_noise([row.contiguous() for row in t(z["noises"], dev)])
    f1 = evaluate.ensemble_evaluate(args, model, [b], dev, q=int(z["q"]), mode="learned")
    # F1 = correct / count: allow one borderline argmax flip per mask (fp32 summation order differs)
    for got, want, cnt in zip(f1, z["f1_ensemble"], n_masks):
        assert abs(got - float(want)) <= 1.0 / cnt + 1e-9, (f1, z["f1_ensemble"])
    sampling.clear_injected()
    sampling.inject_noise([t(z["noise_one"], dev)])
    f1 = evaluate.evaluate(args, model, [b], dev, q=int(z["q"]), mode="learned")
    for got, want, cnt in zip(f1, z["f1_single"], n_masks):
        assert abs(got - float(want)) <= 1.0 / cnt + 1e-9, (f1, z["f1_single"])
    assert not model.training


def test_ensemble_member_logits_match_reference(dev):
    """The mean ensemble logits themselves (1e-4 relative): same sampled edge sets, same weights."""
    from sgs_gnn_b200 import evaluate, sampling
    z, b, model, args = _setup(dev)
    model.eval()
    sampling.clear_injected()
    sampling.inject_noise([row.contiguous() for row in t(z["noises"], dev)])
    cache, acc = {}, None
    with torch.no_grad():
        for _ in range(args.num_samples_eval):
            out = evaluate._member_logits(args, model, b, int(z["q"]), "learned", cache)
            acc = out.clone() if acc is None else acc + out
    mean = (acc / args.num_samples_eval).cpu()
    want = t(z["mean_logits"])
    assert float((mean - want).abs().max() / want.abs().max()) < 1e-4


@pytest.mark.parametrize("mode", ["random", "edge", "full"])
def test_baseline_eval_modes_run(dev, mode):
    from sgs_gnn_b200 import evaluate
    z, b, model, args = _setup(dev)
    f1 = evaluate.ensemble_evaluate(args, model, [b], dev, q=int(z["q"]), mode=mode)
    assert len(f1) == 3 and all(0.0 <= v <= 1.0 for v in f1)
    with pytest.raises(ValueError, match="Invalid mode"):
        evaluate.evaluate(args, model, [b], dev, q=int(z["q"]), mode="bogus")
