"""SURVEY 8f row f4 (second part): the helpers of the plugin's soma branch (Advantra_plugin.cpp:2426-2440) --
Frangi::imerode, Frangi::imdilate and the in-place xy Frangi::imgaussian -- byte for byte.  The oracle ports are pinned
against a golden fixture generated from the compiled reference and against the reference itself where built."""
import os

import numpy as np
import pytest

from pnr_b200.synth import make_volume

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "case_f_soma.npz")
CASES = [((64, 48, 6), 3.0), ((37, 29, 5), 2.5), ((9, 7, 3), 5.0), ((130, 20, 2), 4.0), ((1, 1, 1), 1.0)]


def _vol(w, h, l, seed=3):
    return make_volume(max(w, 8), max(h, 8), max(l, 8), seed=seed)[:l, :h, :w].copy()


def test_ports_match_the_golden_fixture(oracle):
    g = np.load(GOLD)
    I, rad = g["I"], float(g["rad"])
    assert np.array_equal(oracle.imerode(I, rad), g["eroded"])
    assert np.array_equal(oracle.imdilate(I, rad), g["dilated"])
    assert np.array_equal(oracle.imgaussian_xy(I, rad), g["blurred"])
    assert np.array_equal(oracle.imgaussian_xy(oracle.imerode(I, 2.0), 2.0), g["chain"])
    assert g["eroded"].max() < I.max() < 256 and g["dilated"].min() >= I.min()


@pytest.mark.parametrize("shape,rad", CASES)
def test_ports_equal_reference_where_built(oracle, reference, shape, rad):
    if not reference.has_soma:
        pytest.skip("oracle/_ref built without the soma wrappers")
    I = _vol(*shape)
    assert np.array_equal(oracle.imerode(I, rad), reference.imerode(I, rad))
    assert np.array_equal(oracle.imdilate(I, rad), reference.imdilate(I, rad))
    assert np.array_equal(oracle.imgaussian_xy(I, rad), reference.imgaussian_xy(I, rad))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,rad", CASES + [((515, 300, 3), 7.5)])
def test_gpu_helpers_byte_for_byte(oracle, shape, rad):
    import pnr_b200
    I = _vol(*shape)
    assert np.array_equal(pnr_b200.imerode(I, rad), oracle.imerode(I, rad))
    assert np.array_equal(pnr_b200.imdilate(I, rad), oracle.imdilate(I, rad))
    assert np.array_equal(pnr_b200.imgaussian_xy(I, rad), oracle.imgaussian_xy(I, rad))
    # the plugin's chain: erode, then blur the result in place
    assert np.array_equal(pnr_b200.imgaussian_xy(pnr_b200.imerode(I, rad), rad), oracle.imgaussian_xy(oracle.imerode(I, rad), rad))


@pytest.mark.gpu
def test_gpu_helpers_reject_bad_arguments():
    import pnr_b200
    I = _vol(16, 16, 2)
    with pytest.raises(pnr_b200.FrangiGpuError):
        pnr_b200.imgaussian_xy(I, 0.0)
    with pytest.raises(pnr_b200.FrangiGpuError):
        pnr_b200.imgaussian_xy(I, 11.0)          # tap radius 33 > the supported 30


# ---- the two overloads no live code calls: z-scaled erosion (frangi.h:46), 2-D smoothing (frangi.h:44) ----
GOLD_G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "case_g_cold.npz")


def test_cold_overload_ports_match_the_golden_fixture(oracle):
    g = np.load(GOLD_G)
    I, rad, zdist = g["I"], float(g["rad"]), float(g["zdist"])
    assert np.array_equal(oracle.imerode_z(I, rad, zdist), g["eroded_z"])
    assert np.array_equal(oracle.imerode_z(I[:1], rad, zdist), g["eroded_z_plane"])     # l == 1: no z pass
    assert np.array_equal(oracle.imgaussian2d(I[4], 2.0), g["smooth2d"])
    assert (g["eroded_z"] <= oracle.imerode(I, rad)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("shape,rad,zdist", [((64, 48, 9), 3.0, 2.0), ((37, 29, 5), 2.5, 1.0), ((9, 7, 1), 5.0, 2.0),
                                             ((130, 20, 3), 4.0, 0.5)])
def test_gpu_imerode_z_is_byte_exact(oracle, shape, rad, zdist):
    import pnr_b200
    from pnr_b200.frangi import imerode_z
    I = _vol(*shape)
    assert np.array_equal(imerode_z(I, rad, zdist), oracle.imerode_z(I, rad, zdist))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,sigma", [((96, 80), 2.0), ((37, 29), 4.0), ((515, 33), 1.5), ((2, 2), 1.0)])
def test_gpu_imgaussian2d_is_bit_exact(oracle, shape, sigma):
    from pnr_b200.frangi import imgaussian2d
    w, h = shape
    I = _vol(w, h, 1)[0]
    assert np.array_equal(imgaussian2d(I, sigma), oracle.imgaussian2d(I, sigma))
