"""SURVEY 8f row f3, second half: the per-seed correlation score of the plugin's seed filter, Tracker::znccBBB
(tracker.cpp:1891-1964) as called for every extracted seed at Advantra_plugin.cpp:2561-2573.  The fixture
(tests/golden/case_h_zncc.npz, tools/make_golden.py) holds the UNMODIFIED reference's scores for the seeds extractSeeds
finds on a synthetic volume plus crafted ones (off-grid positions, directions along z, seeds at the corners and outside
the volume).  The GPU kernel walks the samples in the reference's order with the reference's float / double mix, so
scores are compared bit for bit: the znccth filter and the sort by score downstream then decide exactly as the plugin."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "case_h_zncc.npz")


def test_fixture_is_the_reference(reference):
    if not reference.has_zncc:
        pytest.skip("oracle/_ref built without ref_seed_zncc")
    g = np.load(GOLD)
    corr, sig = reference.seed_zncc(g["I"], g["sigmas"], g["seeds"])
    assert np.array_equal(corr, g["corr"]) and np.array_equal(sig, g["sig"])
    assert (g["corr"] >= 0.3).sum() > 100 and (g["corr"] < 0.3).sum() > 100      # both sides of the plugin's threshold


@pytest.mark.gpu
def test_gpu_scores_are_bit_identical():
    import pnr_b200
    g = np.load(GOLD)
    corr, sig = pnr_b200.seed_zncc(g["I"], g["sigmas"], g["seeds"])
    assert np.array_equal(corr, g["corr"]), float(np.abs(corr - g["corr"]).max())
    assert np.array_equal(sig, g["sig"])
    # the handle form: the image the handle holds on its device after a run
    l, h, w = g["I"].shape
    p = pnr_b200.FrangiPlan(list(g["sigmas"]), 2.0, .5, .5, 500., False, w, h, l)
    p.run(g["I"], want_J8=True)
    c2, s2 = p.seed_zncc(g["seeds"])
    p.close()
    assert np.array_equal(c2, g["corr"]) and np.array_equal(s2, g["sig"])
    # the plugin's filter and order (Advantra_plugin.cpp:2570-2586): identical decisions
    keep = corr >= 0.3
    assert np.array_equal(keep, g["corr"] >= 0.3)
    assert np.array_equal(np.argsort(-corr[keep], kind="stable"), np.argsort(-g["corr"][keep], kind="stable"))


@pytest.mark.gpu
def test_gpu_zncc_rejects_bad_arguments():
    import pnr_b200
    g = np.load(GOLD)
    with pytest.raises(pnr_b200.FrangiGpuError):
        pnr_b200.seed_zncc(g["I"][:1], g["sigmas"], g["seeds"])          # single plane: the 2-D template is not provided
    c, s = pnr_b200.seed_zncc(g["I"], g["sigmas"], np.zeros((0, 6), np.float32))
    assert len(c) == 0
