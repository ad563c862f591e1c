"""CPU tests of the drop-in boundary: libfrangi_gpu.so loads without a GPU, exports every
function include/frangi_gpu.h declares, and every compute entry point fails loudly
(FRANGI_GPU_ECUDA + message) instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import pnr_b200
from pnr_b200 import frangi as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "frangi_gpu.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(frangi_gpu_[a-z0-9_]+)\s*\(", src)))


def test_header_and_bindings_agree():
    names = _declared()
    assert "frangi_gpu_create" in names and "frangi_gpu_run" in names and "frangi_gpu_destroy" in names
    assert sorted(F.SYMBOLS) == names


def test_library_exports_every_declared_symbol():
    lib = pnr_b200.load_library()          # binds every symbol; raises when one is missing
    out = subprocess.run(["nm", "-D", "--defined-only", F.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (frangi_gpu_[a-z0-9_]+)", out))
    assert exported == set(_declared())
    assert b"sm_100a" in lib.frangi_gpu_version()
    # nothing but the C-ABI leaks out of the shared object
    others = [ln for ln in out.splitlines() if " T " in ln and "frangi_gpu_" not in ln]
    assert not others, others


def test_built_for_sm_100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", F.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def _no_gpu():
    return pnr_b200.load_library().frangi_gpu_device_count() == 0


def test_no_cpu_fallback_without_a_gpu():
    if not _no_gpu():
        pytest.skip("a GPU is visible; the failure path is for GPU-less hosts")
    lib = pnr_b200.load_library()
    with pytest.raises(pnr_b200.FrangiGpuError) as e:
        pnr_b200.FrangiPlan([2.0], 2.0, .5, .5, 500., False, 16, 16, 16)
    assert "error 2" in str(e.value) and "no CPU fallback" in str(e.value)
    I = np.zeros((8, 8, 8), np.uint8)
    with pytest.raises(pnr_b200.FrangiGpuError):
        pnr_b200.Frangi([2.0], 2.0, .5, .5, 500.).frangi3d(I)
    with pytest.raises(pnr_b200.FrangiGpuError):
        pnr_b200.Frangi.imgaussian(I, 2.0, 2.0)
    assert lib.frangi_gpu_launch_count() == 0


def test_argument_errors_reported_before_touching_the_device():
    lib = pnr_b200.load_library()
    h = C.c_void_p()
    s = (C.c_float * 1)(2.0)
    rc = lib.frangi_gpu_create(C.byref(h), s, 0, 2.0, .5, .5, 500., 0, 16, 16, 16, None, 1, 0)
    assert rc == 1 and b"nsig" in lib.frangi_gpu_last_error()
    rc = lib.frangi_gpu_create(C.byref(h), s, 1, 2.0, .5, .5, 500., 0, 16, 16, 1, None, 1, 0)
    assert rc == 1 and b"2-D" in lib.frangi_gpu_last_error()
    assert lib.frangi_gpu_run(None, None, None, None, None, None, None, None, None, None, None) == 1
    lib.frangi_gpu_destroy(None)           # tolerated, like delete on a null pointer
