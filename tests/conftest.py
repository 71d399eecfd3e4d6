import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "coordinatedescent.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


ORACLE_SO = os.path.join(ROOT, "oracle", "libcdref.so")


@pytest.fixture(scope="session")
def ref():
    """Backend bound to the CPU oracle (tests only)."""
    import cdgpu
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return cdgpu.Backend(cdgpu.Lib(ORACLE_SO, "cdref"))


@pytest.fixture(scope="session")
def gpu():
    """Backend bound to libcdgpu.so; fails loudly when it is missing."""
    import cdgpu
    return cdgpu.default()
