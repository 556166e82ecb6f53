import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (directory name has a hyphen)."""
    return importlib.import_module("xai-audio-deepfakes_b200")


@pytest.fixture(scope="session")
def built_lib(pkg):
    pkg._lib.build()
    return pkg._lib.lib()


def golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name))
