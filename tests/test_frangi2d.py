"""SURVEY 8f row f4 (first part): the 2-D path, Frangi::frangi2d / hessian2d (frangi.cpp:392-560).
The oracle port is pinned bit for bit against a golden fixture generated from the compiled reference
(tools/make_golden.py 2d) and against the reference itself where oracle/_ref is built; the GPU path is held to
the BASELINE tolerances against the port (second differences bit-exact in the default smoothing mode)."""
import os

import numpy as np
import pytest

from pnr_b200.synth import make_volume
from tests import parity

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "case_e_frangi2d.npz")


def _image(w, h, seed=4):
    return make_volume(w, h, 8, seed=seed, n_neurites=4)[3]


def test_port_matches_the_golden_fixture(oracle):
    g = np.load(GOLD)
    I, sig = g["I"], [float(s) for s in g["sigmas"]]
    o = oracle.frangi2d(I, sig)
    assert np.array_equal(o["J"], g["J"]) and o["Jmin"] == float(g["Jmin"]) and o["Jmax"] == float(g["Jmax"])
    for k in ("Vx", "Vy", "Vz"):
        assert np.array_equal(o[k], g[k])
    assert np.array_equal(oracle.frangi2d(255 - I, sig, blackwhite=True)["J"], g["J_blackwhite"])
    H = oracle.hessian2d(I, 2.0)
    for k in ("Dyy", "Dxy", "Dxx"):
        assert np.array_equal(H[k], g["H_" + k])
    assert o["Jmax"] > 0.5 and (o["J"] > 0).sum() > 100 and not o["Vz"].any()


@pytest.mark.parametrize("shape,sigs,bw", [((96, 128), [2., 4., 6.], False), ((37, 29), [1., 2.], False),
                                            ((64, 80), [2., 3.], True), ((2, 2), [2.], False), ((5, 300), [3.], False)])
def test_port_equals_reference_where_built(oracle, reference, shape, sigs, bw):
    if not reference.has_2d:
        pytest.skip("oracle/_ref built without the 2-D wrappers")
    h, w = shape
    I = _image(w, h) if min(h, w) > 8 else (np.arange(h * w) % 251).astype(np.uint8).reshape(h, w)
    if bw:
        I = 255 - I
    o, r = oracle.frangi2d(I, sigs, blackwhite=bw), reference.frangi2d(I, sigs, blackwhite=bw)
    assert np.array_equal(o["J"], r["J"]) and o["Jmin"] == r["Jmin"] and o["Jmax"] == r["Jmax"]
    for k in ("Vx", "Vy", "Vz"):
        assert np.array_equal(o[k], r[k])
    ho, hr = oracle.hessian2d(I, sigs[0]), reference.hessian2d(I, sigs[0])
    for k in ("Dyy", "Dxy", "Dxx"):
        assert np.array_equal(ho[k], hr[k])


@pytest.mark.gpu
@pytest.mark.parametrize("shape,sigs,bw,fma", [
    ((96, 128), [2., 4., 6.], False, False), ((37, 29), [1., 2.], False, False), ((64, 80), [2., 3.], True, False),
    ((2, 2), [2.], False, False), ((5, 300), [3.], False, False), ((513, 1027), [2., 4., 6.], False, False),
    ((96, 128), [2., 4., 6.], False, True),
])
def test_gpu_frangi2d_against_the_port(oracle, shape, sigs, bw, fma):
    import pnr_b200
    from pnr_b200.frangi import FLAG_FMA_SMOOTHING
    h, w = shape
    I = _image(w, h) if min(h, w) > 8 else (np.arange(h * w) % 251).astype(np.uint8).reshape(h, w)
    if bw:
        I = 255 - I
    f = pnr_b200.Frangi(sigs, 1.0, .5, .5, 500., 0.5, 15.0, flags=FLAG_FMA_SMOOTHING if fma else 0)
    f.blackwhite = bw
    g = f.frangi2d(I)
    o = oracle.frangi2d(I, sigs, blackwhite=bw)
    if True:      # the second differences are the reference's operations one by one (the FMA flag is ignored in 2-D)
        H, Ho = f.hessian2d(I, sigs[0]), oracle.hessian2d(I, sigs[0])
        for k in ("Dyy", "Dxy", "Dxx"):
            assert np.array_equal(H[k], Ho[k]), k
    rep = parity.vesselness_report(g["J"], o["J"])
    assert rep["n_bad"] == 0, rep
    assert abs(g["Jmax"] - o["Jmax"]) <= max(1e-4 * o["Jmax"], 1e-6) and abs(g["Jmin"] - o["Jmin"]) <= 1e-6
    assert not g["Vz"].any()
    if o["Jmax"] > 0:
        strong = o["J"] > 0.01 * o["Jmax"]
        dv = np.maximum(np.abs(g["Vx"].astype(int) - o["Vx"].astype(int)), np.abs(g["Vy"].astype(int) - o["Vy"].astype(int)))
        assert (dv[strong] <= 1).all(), int((dv[strong] > 1).sum())
        if True:
            assert (g["J"] == o["J"]).mean() > 0.99      # exp through the double routine: nearly always the same float
