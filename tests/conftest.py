import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def space_m1():
    from bayesianinferencedl_b200 import get_space
    return get_space(40, m=1)


@pytest.fixture(scope="session")
def space_m2():
    from bayesianinferencedl_b200 import get_space
    return get_space(40, m=2)


@pytest.fixture(scope="session")
def space_m3():
    from bayesianinferencedl_b200 import get_space
    return get_space(40)


@pytest.fixture(scope="session")
def oracle_m1(space_m1):
    from oracle.thermal_fin_oracle import FinOracle
    return FinOracle(space_m1.mesh().coordinates(), space_m1.mesh().cells())


@pytest.fixture(scope="session")
def oracle_m2(space_m2):
    from oracle.thermal_fin_oracle import FinOracle
    return FinOracle(space_m2.mesh().coordinates(), space_m2.mesh().cells())


@pytest.fixture(scope="session")
def oracle_m3(space_m3):
    from oracle.thermal_fin_oracle import FinOracle
    return FinOracle(space_m3.mesh().coordinates(), space_m3.mesh().cells())


@pytest.fixture(scope="session")
def pod_m3(oracle_m3):
    from oracle.thermal_fin_oracle import pod_basis
    return pod_basis(oracle_m3, n_snapshots=200, basis_size=81, seed=0)


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
