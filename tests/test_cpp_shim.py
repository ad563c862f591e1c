"""The C++ drop-in: the reference's call site (Advantra_plugin.cpp:2488-2512) compiled against
pnr_b200/csrc/frangi.h + libfrangi_shim.so.  CPU: it builds, links and fails loudly without a
GPU.  GPU: its outputs equal the ctypes path bit for bit and match the oracle within tolerance."""
import os
import subprocess

import numpy as np
import pytest

import pnr_b200
from pnr_b200.synth import make_volume
from tests import parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "pnr_b200", "_lib")


@pytest.fixture(scope="module")
def callsite(tmp_path_factory):
    if not os.path.exists(os.path.join(LIBDIR, "libfrangi_shim.so")):
        pytest.skip("libfrangi_shim.so not built")
    exe = str(tmp_path_factory.mktemp("cpp") / "callsite")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "pnr_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "callsite.cpp"), "-o", exe, "-L", LIBDIR,
                    "-lfrangi_shim", "-lfrangi_gpu", "-Wl,-rpath," + LIBDIR], check=True)
    return exe


def _run(exe, I, tmp_path, sigs="2,4,6"):
    l, h, w = I.shape
    inp = tmp_path / "in.u8"
    I.tofile(inp)
    return subprocess.run([exe, str(inp), str(w), str(h), str(l), str(tmp_path / "out"), sigs],
                          capture_output=True, text=True)


def test_callsite_builds_and_fails_loudly_without_gpu(callsite, tmp_path):
    if pnr_b200.load_library().frangi_gpu_device_count() > 0:
        pytest.skip("a GPU is visible")
    r = _run(callsite, make_volume(24, 20, 12, seed=3, n_neurites=2), tmp_path)
    assert r.returncode == 3
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_callsite_matches_oracle_and_ctypes_path(callsite, tmp_path, oracle):
    I = make_volume(96, 80, 24, seed=1, n_neurites=4)
    r = _run(callsite, I, tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    J = np.fromfile(tmp_path / "out.J", np.float32).reshape(I.shape)
    V = [np.fromfile(tmp_path / f"out.{c}", np.uint8).reshape(I.shape) for c in ("Vx", "Vy", "Vz")]
    J8 = np.fromfile(tmp_path / "out.J8", np.uint8).reshape(I.shape)
    ref = oracle.frangi3d(I, [2.0, 4.0, 6.0], 2.0)
    rep = parity.vesselness_report(J, ref["J"])
    assert rep["n_bad"] == 0, rep
    crep = parity.code_report(V, (ref["Vx"], ref["Vy"], ref["Vz"]), ref["J"])
    assert crep["n_bad"] == 0, crep
    g = pnr_b200.Frangi([2.0, 4.0, 6.0], 2.0, .5, .5, 500.).frangi3d_full(I)
    assert np.array_equal(g["J"], J) and np.array_equal(g["J8"], J8)
    for a, b in zip(V, (g["Vx"], g["Vy"], g["Vz"])):
        assert np.array_equal(a, b)
