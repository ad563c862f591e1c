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


# ---- link-completeness: every Frangi:: use of Advantra_plugin.cpp (:1727, :2346, :2432, :2438, :2488-2497) and every
# other public member of the reference's frangi.h, compiled and linked against the shim; the host helpers are run
# against fixtures generated from the compiled reference (tools/make_golden.py, cases C and G)
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def members(tmp_path_factory):
    if not os.path.exists(os.path.join(LIBDIR, "libfrangi_shim.so")):
        pytest.skip("libfrangi_shim.so not built")
    exe = str(tmp_path_factory.mktemp("cpp") / "frangi_members")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "pnr_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "frangi_members.cpp"), "-o", exe, "-L", LIBDIR,
                    "-lfrangi_shim", "-lfrangi_gpu", "-Wl,-rpath," + LIBDIR], check=True)
    return exe


def _pipe(exe, args, data):
    r = subprocess.run([exe] + [str(a) for a in args], input=data, capture_output=True)
    assert r.returncode == 0, (r.returncode, r.stderr)
    return r.stdout


def test_every_frangi_member_of_the_plugin_links_and_eigen_is_bit_identical(members):
    g = np.load(os.path.join(GOLD, "case_c_eigen.npz"))
    A = np.ascontiguousarray(g["A"], np.float64)
    out = np.frombuffer(_pipe(members, ["eigen", len(A)], A.tobytes()), np.float64).reshape(len(A), 12)
    assert np.array_equal(out[:, :9].reshape(-1, 3, 3), g["V"])      # columns = eigenvectors, reference's signs
    assert np.array_equal(out[:, 9:], g["d"])                        # |d0| <= |d1| <= |d2| with its tie rules
    # SURVEY 8c known answers
    assert np.array_equal(out[0, 9:], [0, 0, 0]) and np.array_equal(out[0, :9].reshape(3, 3)[:, 0], [1, 0, 0])
    assert np.array_equal(out[4, 9:], [1, -2, 3])


def test_device_instantiation_of_the_reference_solver_on_the_host_is_bit_identical(tmp_path):
    """pnr_b200/csrc/ref_eigen.h with T = rdouble -- the very instantiation the device pass of
    FRANGI_GPU_FLAG_REFERENCE_DIRECTION runs -- compiled by nvcc for the host (rdouble's operators = plain IEEE double
    operations; on the device the round-to-nearest intrinsics, the same roundings): vectors, signs and values equal the
    compiled reference's on the eigen fixture."""
    nvcc = "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not present")
    exe = str(tmp_path / "ref_eigen_rdouble")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", os.path.join(ROOT, "pnr_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "ref_eigen_rdouble.cu"), "-o", exe], check=True, capture_output=True)
    g = np.load(os.path.join(GOLD, "case_c_eigen.npz"))
    A = np.ascontiguousarray(g["A"], np.float64)
    out = np.frombuffer(_pipe(exe, [len(A)], A.tobytes()), np.float64).reshape(len(A), 12)
    assert np.array_equal(out[:, :9].reshape(-1, 3, 3), g["V"])
    assert np.array_equal(out[:, 9:], g["d"])


def test_host_helpers_match_the_reference(members):
    g = np.load(os.path.join(GOLD, "case_g_cold.npz"))
    t3 = np.frombuffer(_pipe(members, ["dirs3", 90], b""), np.float32).reshape(-1, 3)
    t2 = np.frombuffer(_pipe(members, ["dirs2", 30], b""), np.float32).reshape(-1, 3)
    assert np.array_equal(t3, g["dirs3d"]) and np.array_equal(t2, g["dirs2d"])
    for mode, tab, q, want in (("idx3", g["dirs3d"], g["q3"], g["idx3"]),
                               ("idx2", g["dirs2d"], np.concatenate([g["q2"], np.zeros((len(g["q2"]), 1), np.float32)], 1),
                                g["idx2"])):
        data = np.int32(len(tab)).tobytes() + np.ascontiguousarray(tab, np.float32).tobytes() + \
            np.ascontiguousarray(q, np.float32).tobytes()
        assert np.array_equal(np.frombuffer(_pipe(members, [mode], data), np.uint8), want)
    F = g["F"]
    l, h, w = F.shape
    q = np.concatenate([g["xy"].astype(np.float32), g["zq"][:, None]], 1)
    got = np.frombuffer(_pipe(members, ["interp", w, h, l], F.tobytes() + np.ascontiguousarray(q, np.float32).tobytes()),
                        np.float32)
    assert np.array_equal(got, g["interp"])
    one = np.frombuffer(_pipe(members, ["interp", w, h, 1], F[:1].tobytes() + np.float32([3, 5, 0.7]).tobytes()), np.float32)
    assert one[0] == g["interp_plane"]
