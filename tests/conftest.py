import os
import sys

import pytest

# Ranks emulated by several contexts of ONE process (tests/test_gpu_sharded.py: peer-exchange kernels that wait for each
# other) need their streams on distinct hardware work queues: with the default of 8 connections two streams can share a
# queue, and a kernel queued behind another rank's waiting kernel would never start.  Must be set before CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def gpu_ctx():
    from qurious_b200 import _lib
    return _lib.default_context()
