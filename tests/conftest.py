import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _has_gpu():
    try:
        from sdr_iq_visualizer_b200 import _native
        return _native.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # -m gpu on a box without a GPU must fail loudly, not skip silently: leave the tests alone.
    pass


@pytest.fixture(scope="session")
def golden_stream():
    z = np.load(os.path.join(GOLDEN, "stream_frames.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


@pytest.fixture(scope="session")
def golden_classifier():
    z = np.load(os.path.join(GOLDEN, "classifier_cases.npz"))
    with open(os.path.join(GOLDEN, "classifier_cases.json")) as fh:
        res = json.load(fh)
    return z, res
