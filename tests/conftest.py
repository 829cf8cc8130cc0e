import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def H():
    import healpixmpi_jl_b200 as H_

    return H_


@pytest.fixture(scope="session")
def O():
    from oracle import oracle as O_

    O_.lib()
    return O_


@pytest.fixture(scope="session")
def P():
    from oracle import cpu_port as P_

    P_.lib()
    return P_


def rel_l2(a, b):
    import numpy as np

    a = np.asarray(a)
    b = np.asarray(b)
    den = np.sqrt(np.sum(np.abs(b) ** 2))
    return float(np.sqrt(np.sum(np.abs(a - b) ** 2)) / (den if den > 0 else 1.0))
