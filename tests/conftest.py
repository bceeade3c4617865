import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box via gpurun)")


def _have_gpu():
    try:
        import ctypes
        cudart = ctypes.CDLL("libcudart.so.12")
        n = ctypes.c_int(0)
        return cudart.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


HAVE_GPU = None


def pytest_collection_modifyitems(config, items):
    global HAVE_GPU
    if HAVE_GPU is None:
        HAVE_GPU = _have_gpu()
    if HAVE_GPU:
        # torch bundles its own NCCL: import it BEFORE libfourierflows_b200 dlopens the system libnccl.so.2 (the library then reuses
        # the copy torch loaded; the other order makes a later `import torch` fail with an ImportError)
        try:
            import torch  # noqa: F401
        except Exception:  # noqa: BLE001
            pass
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
