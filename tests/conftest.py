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


def unpack_images(z):
    """inverse of tests/golden/make_golden.py:pack_images"""
    out, o = [], 0
    for s in z["shapes"]:
        n = int(np.prod(s))
        out.append(z["flat"][o:o + n].reshape(s))
        o += n
    return out


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def have_cv2():
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False
