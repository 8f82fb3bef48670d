"""sgs_gnn_b200 -- B200-native (sm_100a) implementation of the SGS-GNN learned-sparsifier
training step behind the reference's own Python API.  See DESIGN.md.

Importing the package never touches CUDA; every compute entry point goes through the
C-ABI library `libsgs_b200.so` (include/sgs_b200.h) and raises if it is missing --
there is no CPU fallback.
"""
__version__ = "0.1.0"
