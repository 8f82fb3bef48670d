"""Drop-in shim: put sgs_gnn_b200/dropin first on sys.path and the reference's main.py /
evaluate.py import this module instead of their own sampling.py."""
from sgs_gnn_b200.sampling import *  # noqa: F401,F403
from sgs_gnn_b200.sampling import gumbel_softmax_sampling, random_edge_sampling  # noqa: F401
