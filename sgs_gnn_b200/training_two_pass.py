"""Drop-in for the reference's training_two_pass.py (same `train` signature and return tuple):
pass 1 scores all E edges without gradient, pass 2 samples q of them, pass 3 re-runs the scorer
(message passing AND scoring) on the sampled subgraph with gradients enabled
(training_two_pass.py:48-81)."""
from ._train_core import train_epoch


def train(args, epoch, max_epoch, model, optimizer_gnn, optimizer_edge_prob, optimizer, criterion, cluster_loader,
          q=500, alternate_frequency=1):
    return train_epoch("two_pass", args, epoch, max_epoch, model, optimizer_gnn, optimizer_edge_prob, optimizer,
                       criterion, cluster_loader, q=q, alternate_frequency=alternate_frequency)
