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


@pytest.fixture(scope="session")
def params():
    from dc_tts_b200.params import init_params
    return init_params(0, "perturbed")


@pytest.fixture(scope="session")
def engine(params):
    """One engine per test session with the 'perturbed' seed-0 weights committed."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dc_tts_b200.engine import Engine, set_engine
    e = Engine(0)
    e.load_params(params)
    set_engine(e)
    return e


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(params=[0, 1, 2], ids=["fp32path", "tensorpath", "fp32chain"])
def path(engine, request):
    """Runs a GPU test once per kernel set (include/dctts.h: dctts_set_tensor_path)."""
    engine.set_tensor_path(request.param)
    yield request.param
    engine.set_tensor_path(1)
