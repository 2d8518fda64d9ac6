import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
MODELS = os.path.join(ROOT, "baseline", "_ref", "models")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def golden_rules():
    return load_golden("rules")


@pytest.fixture(scope="session")
def golden_simulate():
    return load_golden("simulate")


@pytest.fixture(scope="session")
def golden_nets():
    return load_golden("nets")


@pytest.fixture(scope="session")
def rollout_weights():
    """RolloutPolicy parameters. 82 floats; a copy lives in tests/golden so CPU tests never need baseline/_ref."""
    z = np.load(os.path.join(GOLDEN, "rollout_model.npz"))
    return z["conv1/W"], z["bias2/b"]


def model_file(name):
    p = os.path.join(MODELS, name)
    if not os.path.isfile(p):
        pytest.fail(f"{p} missing: run `python oracle/fetch_ref.py` in the build container "
                    "(baseline/_ref is git-ignored but ships with gpurun)")
    return p


@pytest.fixture(scope="session")
def cref():
    from oracle import cref as c
    c.build()
    return c


@pytest.fixture(scope="session")
def engine(rollout_weights):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("a -m gpu test is running without a GPU")
    import iago_b200
    eng = iago_b200.default_engine(0)
    eng.load_rollout(*rollout_weights)
    return eng
