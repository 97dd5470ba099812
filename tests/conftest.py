"""pytest configuration: `-m "not gpu"` runs here (no GPU): oracle KATs, host logic, ABI export checks, gloo tests;
`-m gpu` runs on a B200: the parity tests proper, all of them through the C-ABI of include/rtx_b200.h."""
import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    import make_assets
    make_assets.ensure_assets()
    csrc = os.path.join(ROOT, "go-raytracing_b200", "csrc")
    if not (os.path.exists(os.path.join(csrc, "librtx_b200.so")) and os.path.exists(os.path.join(csrc, "librt_host.so"))):
        subprocess.check_call(["make", "-C", csrc, "-s"])
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])


@pytest.fixture(scope="session")
def grt():
    _ensure_built()
    return importlib.import_module("go-raytracing_b200")


@pytest.fixture(scope="session")
def orc():
    _ensure_built()
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def ctx(grt):
    """One device context shared by the GPU tests (fails loudly if the CUDA library cannot create one)."""
    c = grt.Context(0)
    yield c
    c.close()


REAL_HDR = os.path.join(ROOT, "assets", "hdri", "abandoned_hall_01_1k.hdr")
SYN_HDR = os.path.join(ROOT, "assets", "hdri", "synthetic_hall_1k.hdr")
