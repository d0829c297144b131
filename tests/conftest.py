import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "calypso-gap_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle
    return Oracle("parity")


@pytest.fixture(scope="session")
def shipped_pot(oracle):
    return oracle.read(os.path.join(GOLDEN, "gap_parameters"))


@pytest.fixture(scope="session")
def golden_frames():
    return np.load(os.path.join(GOLDEN, "ase_traj_frames.npz"))


@pytest.fixture(scope="session")
def bc_structure():
    return np.load(os.path.join(GOLDEN, "bc_structure.npz"))
