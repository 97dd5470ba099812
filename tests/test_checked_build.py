"""compute-sanitizer is closed on the GPU pool this repo is measured on (profiles/r02_sanitizer_closed.log), so the race / bounds evidence
for the trace kernels' shared-memory hand-over protocol comes from a CHECKED BUILD of the library (make -C go-raytracing_b200/csrc checked):
every slot claim / publish and every index the traversal dereferences is asserted at run time (csrc/rtx_trace.cuh, RTX_CHECKED) and
violations are counted. This test runs the checked library in a subprocess over flat, hierarchy, flat-top-level and all-features variants of six
scenes, demands zero violations, a working negative control, and the same hit records / ray counts as the production library."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "go-raytracing_b200", "csrc")


def run(lib):
    env = dict(os.environ)
    if lib:
        env["RTX_B200_LIB"] = lib
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_checked.py")], env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.gpu
def test_checked_build_reports_no_violation():
    lib = os.path.join(CSRC, "librtx_b200_checked.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", CSRC, "-s", "checked"])
    chk = run(lib)
    assert chk["checked_build"] == 1
    assert chk["violations"] == 0, f"checked build: {chk['by_kind']}"
    assert chk["selftest_violations"] == 5, "the negative control must be counted: the assertion macro is live"
    ref = run(None)
    assert ref["checked_build"] == 0
    for name, runs in chk["scenes"].items():
        for a, b in zip(runs, ref["scenes"][name]):
            assert (a["entry_sum"], a["prim_sum"], a["t_sum"], a["ext"], a["shadow"]) == (b["entry_sum"], b["prim_sum"], b["t_sum"], b["ext"], b["shadow"]), (name, a["opts"])


def test_checked_build_compiles():
    """CPU side: the checked library builds and exports the ABI."""
    import ctypes
    lib = os.path.join(CSRC, "librtx_b200_checked.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", CSRC, "-s", "checked"])
    L = ctypes.CDLL(lib)
    assert L.rtx_abi_version() == 2 and hasattr(L, "rtx_create_multi")
