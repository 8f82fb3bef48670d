"""TEST INFRASTRUCTURE ONLY (oracle shim)."""


def to_networkx(*a, **k):
    raise NotImplementedError
