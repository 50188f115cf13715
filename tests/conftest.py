import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def weights1234():
    from clair_b200 import weights as W
    return W.random_weights(seed=1234)


@pytest.fixture(scope="session")
def golden_forward():
    with np.load(os.path.join(GOLDEN, "forward_b8.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def gpu_model(weights1234):
    """One engine for the whole GPU session (seed-1234 weights)."""
    from clair_b200.model import Clair
    m = Clair(max_sites=8192, batch_sites=1000)
    m.set_weights(weights1234)
    yield m
    m.close()
