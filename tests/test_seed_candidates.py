"""the per-layer pre-pass of SeedExtractor::extractSeeds (seed.cpp:574-632) on the GPU,
bit for bit against the oracle's restatement (which is pinned through oracle_extract_seeds == the compiled
reference, tests/test_oracle_golden.py)."""
import numpy as np
import pytest

from pnr_b200.synth import make_volume

SIGS = [2.0, 4.0, 6.0]


def _same(g, o):
    assert np.array_equal(g["layer_min"], o["layer_min"]) and np.array_equal(g["layer_max"], o["layer_max"])
    assert np.array_equal(g["n_max"], o["n_max"])
    assert np.array_equal(g["keys"], o["keys"])


def test_oracle_candidates_are_what_extract_seeds_ranks(oracle):
    """CPU: keys decode to interior pixels that differ from the layer minimum and have no greater neighbour."""
    rng = np.random.default_rng(1)
    J8 = rng.integers(0, 6, (3, 17, 23), dtype=np.uint8)
    c = oracle.seed_candidates(J8)
    assert c["n_max"].sum() == len(c["keys"]) and len(c["keys"]) > 0
    off = 0
    for z in range(J8.shape[0]):
        k = c["keys"][off:off + c["n_max"][z]]; off += c["n_max"][z]
        assert np.all(np.diff(k) > 0)
        p = (k & 0xffffffff).astype(np.int64)
        y, x = p // 23, p % 23
        assert np.all((x > 0) & (x < 22) & (y > 0) & (y < 16))
        v = J8[z, y, x].astype(int)
        assert np.all(v != c["layer_min"][z])
        nb = np.stack([J8[z, y + dy, x + dx] for dy in (-1, 0, 1) for dx in (-1, 0, 1)]).astype(int)
        assert np.all(nb.max(0) <= v)
        fac = np.float32(2e9 / float(np.float32(c["layer_max"][z]) - np.float32(c["layer_min"][z])))
        iv = ((v - int(c["layer_min"][z])).astype(np.float32) * fac).astype(np.int64)
        assert np.array_equal(k >> 32, iv)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,kind", [
    ((5, 64, 96), "random"), ((3, 33, 37), "plateaus"), ((4, 2, 2), "random"), ((2, 1, 9), "random"),
    ((3, 40, 50), "flat"), ((7, 129, 255), "smooth"), ((2, 300, 1031), "random"),
])
def test_candidates_on_crafted_volumes(oracle, shape, kind):
    import pnr_b200
    rng = np.random.default_rng(7)
    if kind == "random":
        J8 = rng.integers(0, 256, shape, dtype=np.uint8)
    elif kind == "plateaus":            # few levels: large equal-height plateaus, many ties
        J8 = rng.integers(0, 3, shape, dtype=np.uint8) * 100
    elif kind == "flat":                # max == min: 2e9/0 = inf, no candidate at all
        J8 = np.full(shape, 17, np.uint8); J8[1] = 0
    else:
        z, y, x = np.meshgrid(*[np.arange(n) for n in shape], indexing="ij")
        J8 = (127 + 120 * np.sin(x / 7.0) * np.cos(y / 5.0 + z)).astype(np.uint8)
    _same(pnr_b200.seed_candidates(J8), oracle.seed_candidates(J8))


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [(0,), (0, 0, 0)])
def test_candidates_of_a_resident_run(oracle, devices):
    """After a run the pre-pass works on the J8 left on the device (every slab its own layers)."""
    import pnr_b200
    w, h, l = 160, 128, 48
    I = make_volume(w, h, l, seed=2)
    p = pnr_b200.FrangiPlan(SIGS, 2.0, .5, .5, 500., False, w, h, l, devices=devices)
    out = p.run(I, want_J8=True)
    g = p.seed_candidates()
    p.close()
    o = oracle.seed_candidates(out["J8"])
    _same(g, o)
    assert len(g["keys"]) > 100
