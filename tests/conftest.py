import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def weights_default():
    from erweitern_55_b200 import weights as W
    return W.init_weights(seed=55, perturbed=False)


@pytest.fixture(scope="session")
def weights_perturbed():
    from erweitern_55_b200 import weights as W
    return W.init_weights(seed=55, perturbed=True)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "forward_golden.npz"))
