"""Drop-in shim: put sgs_gnn_b200/dropin first on sys.path and the reference's main.py /
evaluate.py import this module instead of their own model.py."""
from sgs_gnn_b200.model import *  # noqa: F401,F403
