import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cost_golden():
    return np.load(os.path.join(GOLDEN, "cost_golden.npz"))


@pytest.fixture(scope="session")
def lap_golden():
    return np.load(os.path.join(GOLDEN, "lap_golden.npz"))


@pytest.fixture(scope="session")
def engine():
    """The device engine; GPU tests fail loudly (no skip) if CUDA or the .so is missing."""
    import torch
    assert torch.cuda.is_available(), "gpu-marked test run without a CUDA device"
    from cytospace_b200.engine import AssignmentEngine
    return AssignmentEngine()


LAP_NAMES = ["uniform16", "uniform64", "negative33", "ties24", "constant9", "one", "two",
             "duprows30", "dupcols28", "diag40", "pearson48"]
