import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "wmf_py_golden.npz"))


def golden_names(prefix):
    g = np.load(os.path.join(ROOT, "tests", "golden", "wmf_py_golden.npz"))
    return sorted({k.split("/")[1] for k in g.files if k.startswith(prefix + "/")})
