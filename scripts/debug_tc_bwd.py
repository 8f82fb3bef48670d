import sys, torch
sys.path.insert(0, '.')
from sgs_gnn_b200 import ops
dev = torch.device('cuda:0')
def setup(n_nodes, e, h, seed):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n_nodes, (2, e), generator=g)
    out = torch.relu(torch.randn(n_nodes, h, generator=g))
    W1 = (torch.rand(h, 2 * h, generator=g) - 0.5) * (2 / (2 * h) ** 0.5)
    b1 = (torch.rand(h, generator=g) - 0.5) * 0.1
    w2 = (torch.rand(h, generator=g) - 0.5) * (2 / h ** 0.5)
    b2 = torch.tensor([0.05])
    return ei, out, W1, b1, w2, b2
for (h, e, p_drop) in [(256, 128, 0.0), (256, 5000, 0.0), (128, 4000, 0.3)]:
    n_nodes = 600
    ei, out, W1, b1, w2, b2 = setup(n_nodes, e, h, h + e + 1)
    graph = ops.graph_of(ei.to(dev), n_nodes)
    g = torch.Generator().manual_seed(3)
    gup = (torch.randn(e, generator=g) * 1e-6).to(dev)
    grads = {}
    for mode in ("fp32", "bf16", "fp16"):
        leaves = [t.to(dev).clone().requires_grad_(True) for t in (out, W1, b1, w2.reshape(1, -1), b2)]
        p = ops.edge_score(*leaves, graph, None, p_drop, 1234, None, ops._PRECISION[mode])
        grads[mode] = torch.autograd.grad((p * gup).sum(), leaves)
    for mode in ("bf16", "fp16"):
        for name, a, r in zip(("d_out", "dW1", "db1", "dw2", "db2"), grads[mode], grads["fp32"]):
            a = a.flatten().double(); r = r.flatten().double()
            err = float((a - r).abs().max() / (r.abs().max() + 1e-30))
            cos = float((a * r).sum() / (a.norm() * r.norm() + 1e-300))
            ratio = float(a.norm() / (r.norm() + 1e-300))
            print(f"h={h} e={e} drop={p_drop} {mode} {name}: maxrel={err:.3e} cos={cos:.6f} norm_ratio={ratio:.4f}")
        if mode == "bf16":
            a, r = grads[mode][1], grads["fp32"][1]
            # dW1 halves: product part vs difference part
            print("   dW1 prod-part rel", float((a[:, :h]-r[:, :h]).abs().max()/r.abs().max()), "diff-part rel", float((a[:, h:]-r[:, h:]).abs().max()/r.abs().max()))
