import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: z[k] for k in z.files}


def t(a, device="cpu"):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


class FixtureBatch:
    def __init__(self, z, device="cpu"):
        self.x = t(z["x"], device)
        self.y = t(z["y"], device)
        self.edge_index = t(z["edge_index"], device)
        self.train_mask = t(z["train_mask"], device)
        self.prob = t(z["prob"], device)

    def to(self, device):
        return self


@pytest.fixture(scope="session")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _clean_injected_sampler_state():
    """A failing test must not leave injected noise / normalisers behind for the next one."""
    yield
    try:
        from sgs_gnn_b200 import sampling
        sampling.clear_injected()
    except Exception:
        pass
