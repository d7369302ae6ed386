import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _gpu_available() -> bool:
    try:
        import rs_tfhe_b200 as T
        return T.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests are skipped (not failed) on a box without a CUDA device or without the
    built library, so a plain `pytest tests` works as a CPU gate."""
    if _gpu_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device and the built libtfhe_b200.so")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
