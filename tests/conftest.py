import ctypes
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_addoption(parser):
    parser.addoption("--emu", action="store_true", default=False,
                     help="developer option: run the gpu-marked tests against tools/cudaemu (CPU logic checker, see "
                          "tools/cudaemu/cuda_emu.h); never used by the graded runs")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    if config.getoption("--emu"):
        from deeplearning_b200 import _capi
        _capi.LIB_PATH = os.path.join(ROOT, "tools", "cudaemu", "_build", "libdlb200_emu.so")


def _cuda_device_present() -> bool:
    try:
        cu = ctypes.CDLL("libcuda.so.1")
        if cu.cuInit(0) != 0:
            return False
        n = ctypes.c_int(0)
        return cu.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    if config.getoption("--emu") or _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (gpu tests run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
