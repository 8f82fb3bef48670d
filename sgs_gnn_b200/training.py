"""Drop-in for the reference's training.py dispatcher (training.py:6-49): two_pass (the default, as in the
reference), straight_through, hybrid."""
from .training_hybrid import train as train_hybrid
from .training_straight_through import train as train_straight_through
from .training_two_pass import train as train_two_pass


def train(args, epoch, max_epoch, model, optimizer_gnn, optimizer_edge_prob, optimizer, criterion, cluster_loader,
          q=500, alternate_frequency=1):
    pipeline = getattr(args, "pipeline", "two_pass")
    fn = {"straight_through": train_straight_through, "hybrid": train_hybrid}.get(pipeline, train_two_pass)
    return fn(args, epoch, max_epoch, model, optimizer_gnn, optimizer_edge_prob, optimizer, criterion,
              cluster_loader, q=q, alternate_frequency=alternate_frequency)
