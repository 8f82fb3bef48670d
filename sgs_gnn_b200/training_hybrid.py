"""Drop-in for the reference's training_hybrid.py: same `train` signature and return tuple
(training_hybrid.py:7-8,189).  Gradient flow: edge_probs_full is computed for all E edges, a
detached copy drives sampling, and edge_probs_full[mask] carries the gradient -- so only the q
sampled edges are back-propagated through the scorer."""
from ._train_core import train_epoch


def train(args, epoch, max_epoch, model, optimizer_gnn, optimizer_edge_prob, optimizer, criterion, cluster_loader,
          q=500, alternate_frequency=1):
    if epoch == 0:
        use_checkpoint = bool(getattr(args, "hybrid_checkpoint", False))
        print(f"[hybrid] checkpoint={'on' if use_checkpoint else 'off'} (fused scorer: always recomputes)")
    return train_epoch("hybrid", args, epoch, max_epoch, model, optimizer_gnn, optimizer_edge_prob, optimizer,
                       criterion, cluster_loader, q=q, alternate_frequency=alternate_frequency)
