import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "reference_golden.npz"))


@pytest.fixture(scope="session")
def golden_fullsize():
    """Reference outputs at BASELINE's own sizes (tests/golden/make_golden_fullsize.py)."""
    return np.load(os.path.join(GOLDEN_DIR, "reference_golden_fullsize.npz"))


@pytest.fixture(scope="session")
def img01():
    from PIL import Image
    return (np.array(Image.open(os.path.join(GOLDEN_DIR, "img0.pgm"))),
            np.array(Image.open(os.path.join(GOLDEN_DIR, "img1.pgm"))))


@pytest.fixture(scope="session")
def oracle():
    from oracle import klt_oracle
    klt_oracle.lib()
    return klt_oracle


def load_reference():
    """The unmodified reference from oracle/_ref (sourceless bytecode + Cython .so, see oracle/build_ref.py), or None."""
    from oracle import ref_loader
    if not ref_loader.available():
        return None
    return ref_loader.load()


@pytest.fixture(scope="session")
def reference():
    mods = load_reference()
    if mods is None:
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py needs /root/reference)")
    return mods


@pytest.fixture(scope="session")
def gpu_ctx():
    from pyfeaturetrack_b200 import _capi
    return _capi.default_ctx()      # raises loudly if the CUDA library or the GPU is missing
