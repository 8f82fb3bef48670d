"""Drop-in shim: put sgs_gnn_b200/dropin first on sys.path and the reference's main.py imports this module
instead of its own training_two_pass.py."""
from sgs_gnn_b200.training_two_pass import *  # noqa: F401,F403
