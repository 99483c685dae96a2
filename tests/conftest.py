import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: CPU test that takes more than a few seconds")


@pytest.fixture(scope="session")
def golden_ks2d():
    return np.load(GOLDEN / "ks2d_small.npz")


@pytest.fixture(scope="session")
def golden_basic():
    return np.load(GOLDEN / "basic.npz")


@pytest.fixture(scope="session")
def golden_patch():
    return np.load(GOLDEN / "patch.npz")


@pytest.fixture(scope="session")
def golden_configs():
    return json.loads((GOLDEN / "ks2d_configs.json").read_text())


@pytest.fixture(scope="session")
def ks_default_stack():
    """The reference's default simulated stack (ks2d:751-782), made by the oracle's
    restatement of simulate(); pinned to the reference by the U_checksum goldens."""
    from oracle import ks2d

    return ks2d.simulate(ks2d.SimConfig())
