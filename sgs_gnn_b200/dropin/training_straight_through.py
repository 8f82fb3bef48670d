"""Drop-in shim: put sgs_gnn_b200/dropin first on sys.path and the reference's main.py /
evaluate.py import this module instead of their own training_straight_through.py."""
from sgs_gnn_b200.training_straight_through import *  # noqa: F401,F403
