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


def pytest_sessionfinish(session, exitstatus):
    """Observed parity margins of this session -> gpurun_out/parity_margins.json (the GPU box only brings gpurun_out/ back;
    the committed copy is profiles/r02_parity_margins.json)."""
    try:
        from tests import parity
        if not parity.MARGINS:
            return
        out_dir = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "parity_margins.json"), "w") as fh:
            json.dump({"tolerances": {"db_above_floor": parity.DB_TOL, "rel": parity.REL_TOL, "u8_tie": parity.TIE_TOL},
                       "checks": parity.MARGINS}, fh, indent=1)
    except Exception:
        pass
