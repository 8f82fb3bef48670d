"""Drop-in for the reference's training.py dispatcher (training.py:6-49).  The two_pass pipeline
is outside this build's hot path and raises."""
from .training_hybrid import train as train_hybrid
from .training_straight_through import train as train_straight_through


def train(args, epoch, max_epoch, model, optimizer_gnn, optimizer_edge_prob, optimizer, criterion, cluster_loader,
          q=500, alternate_frequency=1):
    pipeline = getattr(args, "pipeline", "two_pass")
    fn = {"straight_through": train_straight_through, "hybrid": train_hybrid}.get(pipeline)
    if fn is None:
        raise NotImplementedError(f"pipeline {pipeline!r} is not part of the B200 hot path (hybrid | straight_through)")
    return fn(args, epoch, max_epoch, model, optimizer_gnn, optimizer_edge_prob, optimizer, criterion,
              cluster_loader, q=q, alternate_frequency=alternate_frequency)
