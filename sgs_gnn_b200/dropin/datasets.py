"""Drop-in shim for the reference's datasets.add_degree."""
from sgs_gnn_b200.datasets import *  # noqa: F401,F403
