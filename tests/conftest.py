"""pytest configuration: the `gpu` marker, import paths, shared fixtures.

-m "not gpu"  : CPU oracle vs the committed golden vectors, host logic, C-ABI symbol checks, gloo window
                test. No compute call is made into libgsr_b200.so without a GPU.
-m gpu        : the parity tests proper; they call through the C ABI (ctypes) on cuda:0.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_library():
    """Build (incrementally: only what is older than its sources) and load libgsr_b200.so, so that a stale prebuilt
    binary shipped with the snapshot is never what gets tested. Loading does not need a GPU."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gsr_build", os.path.join(ROOT, "gs-livm_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    import gs_livm_b200  # noqa: F401
    from gs_livm_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu but no CUDA device is visible (no CPU fallback exists)")
    return torch.device("cuda:0")
