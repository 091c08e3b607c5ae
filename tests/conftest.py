import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


@pytest.fixture(scope="session")
def golden_process():
    return np.load(os.path.join(GOLDEN, "process.npz"))


@pytest.fixture(scope="session")
def golden_decode():
    return np.load(os.path.join(GOLDEN, "decode.npz"))


@pytest.fixture(scope="session")
def golden_bursts():
    return np.load(os.path.join(GOLDEN, "bursts.npz"))


@pytest.fixture(scope="session")
def golden_psd():
    return np.load(os.path.join(GOLDEN, "psd.npz"))


@pytest.fixture(scope="session")
def process_inputs():
    from tests.golden.cases import process_cases
    return {name: (fs, f, iq) for name, fs, f, iq in process_cases()}
