import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _load_npz(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def dtu_weights():
    """DTU checkpoint weights (tests/golden/dtu_weights.npz) as {state_dict key: fp32 CPU tensor}."""
    return {k: torch.from_numpy(v) for k, v in _load_npz("dtu_weights.npz").items()}


@pytest.fixture(scope="session")
def stage_kats():
    return _load_npz("stage_kats.npz")


@pytest.fixture(scope="session")
def e2e_d8():
    return _load_npz("e2e_d8.npz")


@pytest.fixture(scope="session")
def e2e_d32():
    return _load_npz("e2e_d32.npz")


@pytest.fixture(scope="session")
def e2e_allpred():
    """Reference Pipeline(test=False).eval() outputs + full_loss (tests/golden/make_golden_train.py)."""
    return _load_npz("e2e_allpred_d32.npz")


@pytest.fixture(scope="session")
def loss_kat():
    return _load_npz("loss_kat.npz")


@pytest.fixture(scope="session")
def fusion_kat():
    """Outputs of the reference's reproject_with_depth / check_geometric_consistency (tests/golden/make_golden_fusion.py)."""
    return _load_npz("fusion_kat.npz")
