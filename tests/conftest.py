import glob
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A machine without any NVIDIA device node (the build container) skips the gpu-marked tests instead of failing
    them in jpgb_encoder_create. A machine that HAS a device runs them, and a broken driver or a missing library
    fails loudly there: there is no CPU fallback to fall back to."""
    if glob.glob("/dev/nvidia[0-9]*"):
        return
    skip = pytest.mark.skip(reason="no NVIDIA device on this machine")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
