import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="session")
def orc():
    import __graft_entry__ as ge
    return ge.load_oracle()


@pytest.fixture(scope="session")
def emu(pkg):
    """CPU-thread emulation build of the library: same host code, same kernel sources compiled by g++.
    Validates planner / executor / kernel index arithmetic without a GPU; it is a test tool, not a product path."""
    if not os.path.exists(pkg.EMU_LIB_PATH):
        import subprocess
        subprocess.check_call(["make", "-j8", "emu"], cwd=ROOT)
    return pkg.load(emulated=True).setup()


@pytest.fixture(scope="session")
def gpu(pkg):
    """the real library on a CUDA device; fails loudly when the device or the library is missing"""
    lib = pkg.load(emulated=False).setup()
    assert lib.have_device(), "libp3dfft.3.so found no usable CUDA device"
    return lib
