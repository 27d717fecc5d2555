import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "reinmav-gym_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _ensure_library():
    """librmv.so is a build artefact (git-ignored).  Build it when a fresh checkout has none, so that the
    C-ABI tests never silently run against nothing."""
    lib = os.path.join(ROOT, "reinmav-gym_b200", "gym_reinmav_b200", "lib", "librmv.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "reinmav-gym_b200", "csrc")],
                              stdout=subprocess.DEVNULL)


def pytest_configure(config):
    _ensure_library()
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir("/root/reference/gym_reinmav")
    have_gpu = None
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference not present on this machine"))
        if "gpu" in item.keywords:
            if have_gpu is None:
                have_gpu = _has_gpu()
            if not have_gpu:
                item.add_marker(pytest.mark.skip(reason="no CUDA device"))


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def host_twin():
    """Host build of csrc/rmv_math.cuh (tests/host_twin) — checks the kernel arithmetic without a GPU."""
    import ctypes
    from tests.host_twin.build import build
    lib = ctypes.CDLL(build())
    lib.twin_step.restype = ctypes.c_int
    return lib
