"""Drop-in for the reference's training_straight_through.py (train signature :7-8, return :176):
the GNN consumes (p * st)[mask].clamp(0,1) and the gradient reaches every edge through the
sum(p) normaliser."""
from ._train_core import train_epoch


def train(args, epoch, max_epoch, model, optimizer_gnn, optimizer_edge_prob, optimizer, criterion, cluster_loader,
          q=500, alternate_frequency=1):
    return train_epoch("straight_through", args, epoch, max_epoch, model, optimizer_gnn, optimizer_edge_prob,
                       optimizer, criterion, cluster_loader, q=q, alternate_frequency=alternate_frequency)
