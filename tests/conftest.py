"""pytest configuration: the `gpu` marker and shared fixtures (inputs, oracle)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def meridian_raw():
    """The reference's own 32-column regression input (test/ifs/ecrad_meridian.nc), as committed fixture."""
    raw = dict(np.load(os.path.join(GOLDEN, "ecrad_meridian_inputs.npz")))
    return {k: np.array(v, dtype=np.float64) for k, v in raw.items()}


@pytest.fixture(scope="session")
def golden_noaer():
    return dict(np.load(os.path.join(GOLDEN, "ecrad_meridian_noaer_ref.npz")))


@pytest.fixture(scope="session")
def golden_cloudless():
    return dict(np.load(os.path.join(GOLDEN, "ecrad_meridian_cloudless_ref.npz")))


@pytest.fixture(scope="session")
def golden_default():
    return dict(np.load(os.path.join(GOLDEN, "ecrad_meridian_default_ref.npz")))


@pytest.fixture(scope="session")
def golden_expexp():
    return dict(np.load(os.path.join(GOLDEN, "ecrad_meridian_expexp_ref.npz")))


@pytest.fixture(scope="session")
def golden_tripleclouds():
    return dict(np.load(os.path.join(GOLDEN, "ecrad_meridian_tripleclouds_ref.npz")))


@pytest.fixture(scope="session")
def golden_ecckd_mcica():
    return dict(np.load(os.path.join(GOLDEN, "ecrad_meridian_ecckd_mcica_ref.npz")))


@pytest.fixture(scope="session")
def golden_ecckd_tc():
    return dict(np.load(os.path.join(GOLDEN, "ecrad_meridian_ecckd_tc_ref.npz")))
