import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle module (oracle/lapper_oracle.py), built on demand with gcc."""
    from oracle import lapper_oracle

    lapper_oracle.load()
    return lapper_oracle


def pytest_sessionfinish(session, exitstatus):
    """With a -DLAPPER_DEBUG_CHECKS build of the library (LAPPER_B200_LIB=...liblapper_b200_dbg.so) the query kernels
    count violated bounds checks on the device; report them and fail the run if there were any."""
    if "rust_lapper_b200._ffi" not in sys.modules:
        return
    try:
        lib = sys.modules["rust_lapper_b200._ffi"].lib()
        n = int(lib.lapper_debug_violations())
    except Exception:
        return
    if n >= 0:
        print(f"\n[lapper] debug-checks build: {n} violated bounds checks in the query kernels")
        if n > 0:
            session.exitstatus = 1
