import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) device and the built libafb200.so")


@pytest.fixture(scope="session")
def golden_model():
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "model_golden.npz")))


@pytest.fixture(scope="session")
def golden_crop():
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "crop_golden.npz")))


@pytest.fixture(scope="session")
def state_dict():
    import afb200
    return afb200.synthetic.synthetic_state_dict(0)
