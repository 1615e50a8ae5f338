import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import oracle_lib  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracles():
    # building the checkers is not using them; the reference build only happens where /root/reference is mounted
    oracle_lib.build_oracles()


@pytest.fixture(scope="session")
def port():
    return oracle_lib.Oracle("port")


@pytest.fixture(scope="session")
def ref():
    try:
        return oracle_lib.Oracle("ref")
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built (no /root/reference on this box and no prebuilt library)")


@pytest.fixture(scope="session")
def best_oracle():
    """The strongest checker present: the reference build when it exists, else the port."""
    try:
        return oracle_lib.Oracle("ref")
    except FileNotFoundError:
        return oracle_lib.Oracle("port")


@pytest.fixture
def rng():
    return np.random.Generator(np.random.PCG64(0x5E1E217E))
