"""TEST INFRASTRUCTURE ONLY (oracle shim)."""


class Data:
    pass
