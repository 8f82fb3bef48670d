"""TEST INFRASTRUCTURE ONLY (oracle shim) -- not a product path.

Minimal stand-in for PyTorch-Geometric 2.3.1 so that the reference's own
model.py / utils.py import unmodified.  Only GCNConv
carries arithmetic; everything else is a name stub.
"""
