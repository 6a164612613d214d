import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def rel_l2(a, b):
    "relative L2 error of a against reference b (numpy or torch, real or complex)"
    import torch
    a = torch.as_tensor(np.asarray(a)) if not hasattr(a, "detach") else a.detach().cpu()
    b = torch.as_tensor(np.asarray(b)) if not hasattr(b, "detach") else b.detach().cpu()
    if a.is_complex() or b.is_complex():
        a, b = a.to(torch.complex128), b.to(torch.complex128)
    else:
        a, b = a.double(), b.double()
    den = torch.linalg.vector_norm(b).item()
    num = torch.linalg.vector_norm(a - b).item()
    return num / den if den > 0 else num


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]

    return get
