"""FRANGI_GPU_FLAG_REFERENCE_DIRECTION: the direction bytes with the reference's eigenvector SIGN.

The reference writes column 0 of what its double-precision Householder / QL solver returns (frangi.cpp:198, 239-250,
1269-1495); the sign of that vector is an accident of the iteration.  The fast path's closed form returns the same axis
with the other sign in half of the voxels, and the plugin downstream is not indifferent to it (test_plugin_e2e.py).
With the flag the device re-solves every voxel a scale wins with the reference's own algorithm (pnr_b200/csrc/ref_eigen.h,
the code the host members of the drop-in class run, bit-identical to the compiled reference on the CPU:
test_cpp_shim.py), so that Vx / Vy / Vz are the reference's bytes."""
import numpy as np
import pytest

import pnr_b200
from pnr_b200.synth import make_volume

SIGS = [2.0, 4.0, 6.0]
KEYS = ("Vx", "Vy", "Vz")


def _gpu(I, sigs, flags, devices=(0,)):
    f = pnr_b200.Frangi(sigs, 2.0, 0.5, 0.5, 500.0, devices=devices, flags=flags)
    g = f.frangi3d_full(I, want_J8=True)
    f.close()
    return g


def test_flag_is_declared_and_distinct():
    flags = [pnr_b200.FLAG_FMA_SMOOTHING, pnr_b200.FLAG_DIR_F32, pnr_b200.FLAG_SCALE_IDX, pnr_b200.FLAG_LOCAL_HALO,
             pnr_b200.FLAG_OVERLAP_Z, pnr_b200.FLAG_REFERENCE_DIRECTION]
    assert len(set(flags)) == len(flags) and pnr_b200.FLAG_REFERENCE_DIRECTION == 32
    with open(pnr_b200.frangi.__file__.replace("pnr_b200/frangi.py", "include/frangi_gpu.h")) as f:
        assert "FRANGI_GPU_FLAG_REFERENCE_DIRECTION = 32" in f.read()


@pytest.mark.gpu
@pytest.mark.parametrize("shape,sigs,seed", [((192, 160, 48), SIGS, 11), ((37, 29, 11), [1.0, 2.0, 3.0, 4.0], 5)])
def test_direction_bytes_are_the_references(reference, shape, sigs, seed):
    """Bit-exact smoothing + the flag: Vx, Vy, Vz equal the reference's bytes in every voxel (faces and the J = 0 background
    included), up to a handful where the two arithmetics pick a different arg-max scale; without the flag about half of
    the voxels carry the other sign.  J, Jmin, Jmax, J8 do not depend on the flag."""
    I = make_volume(*shape, seed=seed, n_neurites=6)
    r = reference.frangi3d(I, sigs)
    plain = _gpu(I, sigs, 0)
    g = _gpu(I, sigs, pnr_b200.FLAG_REFERENCE_DIRECTION)
    n = I.size
    differ = np.zeros(I.shape, bool)
    differ_plain = np.zeros(I.shape, bool)
    for k in KEYS:
        differ |= g[k] != r[k]
        differ_plain |= plain[k] != r[k]
    print("voxels", n, "direction bytes differing from the reference: with the flag", int(differ.sum()),
          "without", int(differ_plain.sum()))
    assert differ_plain.sum() > 0.2 * n                      # the closed form's sign is a coin toss against the QL iteration's
    assert differ.sum() <= 2 + 1e-4 * n, int(differ.sum())
    assert np.array_equal(g["J"], plain["J"]) and np.array_equal(g["J8"], plain["J8"])
    assert g["Jmin"] == plain["Jmin"] and g["Jmax"] == plain["Jmax"]


@pytest.mark.gpu
def test_flag_with_slabs_float_directions_and_fma_smoothing(reference):
    """Three z-slabs == one slab, bit for bit; the float direction is the vector the bytes quantise; with fused
    multiply-add smoothing the second differences change in their last bits, the signs still are the reference's
    wherever the direction is well determined (strong voxels: the axis agrees within 0.5 degrees and the dot product
    with the reference's direction is positive)."""
    I = make_volume(150, 70, 48, seed=33, n_neurites=6)
    fl = pnr_b200.FLAG_REFERENCE_DIRECTION | pnr_b200.FLAG_DIR_F32
    one = _gpu(I, SIGS, fl)
    many = _gpu(I, SIGS, fl, devices=(0, 0, 0))
    for k in KEYS + ("J", "J8", "dir"):
        assert np.array_equal(one[k], many[k]), k
    dec = np.stack([one[k].astype(np.float32) / 255 * 2 - 1 for k in KEYS])
    assert np.max(np.abs(dec - one["dir"])) <= 1.0 / 255 + 1e-6
    r = reference.frangi3d(I, SIGS)
    fma = _gpu(I, SIGS, pnr_b200.FLAG_REFERENCE_DIRECTION | pnr_b200.FLAG_FMA_SMOOTHING)
    strong = r["J"] > 0.01 * r["Jmax"]
    vr = np.stack([r[k].astype(np.float32) / 255 * 2 - 1 for k in KEYS])
    vf = np.stack([fma[k].astype(np.float32) / 255 * 2 - 1 for k in KEYS])
    dot = np.sum(vr * vf, axis=0)[strong]
    print("fma smoothing + flag: strong voxels", int(strong.sum()), "with the other sign", int(np.sum(dot < 0)))
    assert np.sum(dot < 0) <= 2 + 1e-3 * strong.sum()


@pytest.mark.gpu
def test_whole_plugin_with_reference_signs_writes_the_reference_files():
    """The reference's unmodified plugin with the drop-in class and PNR_FRANGI_FLAGS=32 (the only way an unchanged call
    site can pass a flag: pnr_b200/csrc/frangi_shim.cpp) against the all-reference build: every file the plugin
    writes -- seeds, every node list, the final SWC -- is identical."""
    from oracle import Plugin
    from tests.plugin_arms import compare_files, run_arm
    from tests.test_plugin_e2e import DUMP, FINAL, README_PARAMS
    for arm in ("ref", "gpu"):
        if not Plugin.available(arm):
            pytest.skip(f"oracle/_ref/libpnr_plugin_{arm}.so not built")
    I = make_volume(192, 160, 48, seed=11)
    a = run_arm("gpu", I, README_PARAMS, 10, env=dict(PNR_FRANGI_FLAGS="32"))
    b = run_arm("ref", I, README_PARAMS, 10)
    rep = compare_files(a["files"], b["files"])
    print("whole plugin, reference signs:", {k: v["identical"] for k, v in rep.items()})
    assert FINAL in a["files"]
    assert all(v["identical"] for k, v in rep.items() if k != DUMP), {k: v for k, v in rep.items() if not v["identical"]}
    d = rep[DUMP]         # every 10th voxel with J8 > 0: the one J8 voxel in 1.5 million that differs may add or drop a row
    assert d["identical"] or (d["loci_match"] >= 0.9999 and d["sign_flipped"] <= 1 and d["direction_over_1deg"] <= 1), d
