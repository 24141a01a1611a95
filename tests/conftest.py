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


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def sub(d, prefix):
    """Collect `prefix*` scalar entries of a golden file into a float dict."""
    return {k[len(prefix):]: float(v) for k, v in d.items() if k.startswith(prefix) and np.ndim(v) == 0}


@pytest.fixture(scope="session")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vo_single_camera_sos_b200 import ops
    return ops.Context(0)
