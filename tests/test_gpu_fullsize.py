"""Parity at BASELINE.json's full size (configs[3]: sigma = 2,4,6 on 2048x2048x512, 2^31 voxels),
where the reference itself cannot run in one piece (its int voxel index overflows, SURVEY.md
section 0 item 4).  The workload is the benchmark's: a seeded 512x512x128 block tiled 4x4x4.

  * crops: the oracle run on a cropped neighbourhood (crop + halo of ceil(3 sigma) + 2 in x / y
    and ceil(3 sigma / zdist) + 2 in z) is exact on the crop, whatever the size of the volume
    around it -- the same argument as the slab-assembled oracle.  Crops sit at the volume's
    first voxel, at its last voxel (linear index 2^31 - 1: 64-bit indexing), on faces, edges
    and in the interior.
  * periodicity: two steps of halo away from the faces the output must repeat with the tile
    period, bit for bit (same arithmetic on the same neighbourhood).
  * Jmax equals the maximum over the crops' blocks; the J8 map obeys the reference's rule.
"""
import numpy as np
import pytest

import pnr_b200
from pnr_b200 import FLAG_FMA_SMOOTHING, FrangiPlan
from pnr_b200.synth import make_volume
from tests import parity

pytestmark = pytest.mark.gpu

W, H, L = 2048, 2048, 512
BW, BH, BL = 512, 512, 128
SIGS = [2.0, 4.0, 6.0]
HXY, HZ = 20, 11          # ceil(3*6)+2, ceil(3*6/2)+2


@pytest.fixture(scope="module")
def fullsize():
    lib = pnr_b200.load_library()
    base = make_volume(BW, BH, BL)
    I = np.tile(base, (L // BL, H // BH, W // BW))
    out = {}
    for name, flags in (("exact", 0), ("fma", FLAG_FMA_SMOOTHING)):
        p = FrangiPlan(SIGS, 2.0, .5, .5, 500., False, W, H, L, flags=flags)
        r = p.run(I, want_J8=True)
        p.close()
        out[name] = r
    return I, out


def _halo(sigs, zdist=2.0):
    import math
    return math.ceil(3 * max(sigs)) + 2, math.ceil(3 * max(sigs) / zdist) + 2


def crop_compare(oracle, I, g, sigs, crop, exact_dirs, want_scale=False, want_dir=False, zdist=2.0):
    """The oracle on `crop` + halo of the volume I against the same region of the GPU outputs g.  Vesselness within
    BASELINE's tolerance everywhere; direction codes (and float directions) within one code / 0.5 degrees modulo sign
    where the response is strong -- everywhere when the smoothing is bit-identical (`exact_dirs`), and otherwise
    everywhere the reference's own direction is well conditioned (parity.direction_gap); arg-max scale exact away
    from ties."""
    l, h, w = I.shape
    x0, x1, y0, y1, z0, z1 = crop
    hxy, hz = _halo(sigs, zdist)
    xa, xb = max(x0 - hxy, 0), min(x1 + hxy, w)
    ya, yb = max(y0 - hxy, 0), min(y1 + hxy, h)
    za, zb = max(z0 - hz, 0), min(z1 + hz, l)
    # a crop edge that is not a volume face would be treated as one by the oracle: only valid
    # because the halo keeps it hxy / hz away from the region that is compared
    sub = np.ascontiguousarray(I[za:zb, ya:yb, xa:xb])
    ref = oracle.frangi3d(sub, sigs, zdist, want_scale=False, want_dir=want_dir)
    sl = (slice(z0 - za, z1 - za), slice(y0 - ya, y1 - ya), slice(x0 - xa, x1 - xa))
    gs = (slice(z0, z1), slice(y0, y1), slice(x0, x1))
    Jr, Jg = ref["J"][sl], g["J"][gs]
    err = np.abs(Jg.astype(np.float64) - Jr)
    tol = np.maximum(parity.J_RTOL * np.abs(Jr), parity.J_ATOL)
    assert (err <= tol).all(), (crop, float(err.max()))
    strong = Jr > parity.STRONG_FRAC * g["Jmax"]
    singles = None
    if want_scale or (strong.any() and not exact_dirs):
        singles = [oracle.frangi3d(sub, [sg], zdist, want_scale=False, want_dir=False)["J"][sl] for sg in sigs]
    if strong.any():
        a = np.stack([g[k][gs] for k in ("Vx", "Vy", "Vz")]).astype(np.int32)
        b = np.stack([ref[k][sl] for k in ("Vx", "Vy", "Vz")]).astype(np.int32)
        bad = strong & ~((np.abs(a - b) <= 1).all(0) | (np.abs(a - (255 - b)) <= 1).all(0))
        if want_dir:
            d, r = g["dir"][(slice(None),) + gs].astype(np.float64), ref["dir"][(slice(None),) + sl].astype(np.float64)
            dot = np.abs((d * r).sum(0)) / np.maximum(np.sqrt((d * d).sum(0) * (r * r).sum(0)), 1e-300)
            bad |= strong & (np.degrees(np.arccos(np.clip(dot, 0, 1))) > parity.DIR_DEG)
        if exact_dirs:
            assert not bad.any(), (crop, int(bad.sum()), int(strong.sum()))
        else:
            # FMA smoothing moves the Hessian in its last bits; that may only turn a direction the reference itself
            # does not determine.  Every offending voxel must be one of: (a) a tie between scales -- the two best
            # single-scale responses agree within the vesselness tolerance, so which scale's direction is written is
            # decided by the last bit (north_star: arg-max scale exact AWAY FROM TIES); (b) |l1| ~ |l2| at the winning
            # scale in the REFERENCE's arithmetic (parity.direction_gap).
            assert int(bad.sum()) <= 8, (crop, int(bad.sum()))
            S = np.sort(np.stack([np.asarray(q, np.float64) for q in singles]), 0)
            won = np.stack(singles).argmax(0)
            for z, y, x in zip(*np.nonzero(bad)):
                best, second = S[-1, z, y, x], (S[-2, z, y, x] if len(singles) > 1 else -1.0)
                if best - second <= max(parity.J_RTOL * best, parity.J_ATOL):
                    continue
                gap = parity.direction_gap(oracle, sub, sigs, zdist, (z + z0 - za, y + y0 - ya, x + x0 - xa), int(won[z, y, x]))
                assert gap < parity.ILL_CONDITIONED_GAP, (crop, (z, y, x), gap, best, second)
    if want_scale:
        srep = parity.scale_report(g["scale"][gs], singles)
        assert srep["n_bad"] == 0, (crop, srep)
    return int(strong.sum())


CROPS = [
    (0, 48, 0, 40, 0, 20),                          # first voxel: three faces meet
    (W - 48, W, H - 40, H, L - 20, L),              # last voxel: linear index 2^31 - 1
    (1000, 1048, 0, 40, 250, 270),                  # y face
    (0, 48, 1500, 1540, 100, 120),                  # x face
    (700, 760, 900, 950, L - 20, L),                # z face
    (1210, 1270, 1330, 1380, 300, 324),             # interior, straddling tile seams (x 1024+.., y 1024+.., z 256+..)
    (W - 60, W, 3, 50, 127, 150),                   # x face at a z seam
]


@pytest.mark.parametrize("mode", ["exact", "fma"])
def test_crops_against_the_oracle(oracle, fullsize, mode):
    I, out = fullsize
    g = out[mode]
    assert g["Jmax"] > 0 and g["Jmin"] == 0.0
    n_strong = sum(crop_compare(oracle, I, g, SIGS, c, exact_dirs=(mode == "exact")) for c in CROPS)
    assert n_strong > 1000


def _tiled(w, h, l, bw, bh, bl, seed=None):
    base = make_volume(bw, bh, bl) if seed is None else make_volume(bw, bh, bl, seed=seed)
    return np.ascontiguousarray(np.tile(base, (l // bl, h // bh, w // bw)))


def test_config2_size_512x512x128(oracle):
    """BASELINE.json configs[1]: sigma = 2,4,6 on 512x512x128 (the benchmark's base block itself), both smoothing modes,
    arg-max scale and float direction kept; crops on faces, at the last voxel and inside; seeds >= 99.9 % on a crop."""
    w, h, l = 512, 512, 128
    I = make_volume(w, h, l)
    crops = [(0, 56, 0, 48, 0, 24), (w - 56, w, h - 48, h, l - 24, l), (200, 264, 300, 356, 50, 80), (0, 48, 230, 280, 100, 128)]
    for flags, exact in ((pnr_b200.FLAG_DIR_F32 | pnr_b200.FLAG_SCALE_IDX, True),
                         (pnr_b200.FLAG_DIR_F32 | pnr_b200.FLAG_SCALE_IDX | FLAG_FMA_SMOOTHING, False)):
        p = FrangiPlan(SIGS, 2.0, .5, .5, 500., False, w, h, l, flags=flags)
        g = p.run(I, want_J8=True)
        p.close()
        n = sum(crop_compare(oracle, I, g, SIGS, c, exact_dirs=exact, want_scale=True, want_dir=True) for c in crops)
        assert n > 500
    # downstream seed set (the reference's extractSeeds on the GPU's J8 / V vs on the oracle's), on a z range:
    # extractSeeds works layer by layer, so the layers [40, 72) of the full maps are a complete sub-problem
    za, zb = 40 - HZ, 72 + HZ
    ref = oracle.frangi3d(np.ascontiguousarray(I[za:zb]), SIGS, 2.0, want_scale=False, want_dir=False)
    sl = slice(40 - za, 72 - za)
    j8_ref = oracle.j_to_j8(ref["J"][sl], g["Jmin"], g["Jmax"])        # same global scalars as the GPU run
    s_r = oracle.extract_seeds(5.0, j8_ref, *[np.ascontiguousarray(ref[k][sl]) for k in ("Vx", "Vy", "Vz")])
    s_g = oracle.extract_seeds(5.0, np.ascontiguousarray(g["J8"][40:72]), *[np.ascontiguousarray(g[k][40:72]) for k in ("Vx", "Vy", "Vz")])
    rep = parity.seed_report(s_g, s_r)
    assert rep["n_ref"] > 50 and rep["match"] >= parity.SEED_MATCH, rep


def test_config3_size_1024x1024x256_six_scales_with_directions(oracle):
    """BASELINE.json configs[2]: sigma = 1..6 on 1024x1024x256 with the direction-vector (and arg-max scale) outputs."""
    w, h, l = 1024, 1024, 256
    sigs = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0]
    I = _tiled(w, h, l, 256, 256, 64)
    crops = [(0, 40, 0, 36, 0, 16), (w - 40, w, h - 36, h, l - 16, l), (500, 548, 250, 290, 120, 140),
             (240, 280, 0, 36, 60, 76), (700, 740, 500, 540, l - 16, l)]
    for flags, exact in ((pnr_b200.FLAG_DIR_F32 | pnr_b200.FLAG_SCALE_IDX, True),
                         (pnr_b200.FLAG_DIR_F32 | pnr_b200.FLAG_SCALE_IDX | FLAG_FMA_SMOOTHING, False)):
        p = FrangiPlan(sigs, 2.0, .5, .5, 500., False, w, h, l, flags=flags)
        g = p.run(I, want_J8=True)
        p.close()
        assert g["Jmin"] == 0.0 and float(g["J"].max()) == g["Jmax"]
        n = sum(crop_compare(oracle, I, g, sigs, c, exact_dirs=exact, want_scale=True, want_dir=True) for c in crops)
        assert n > 300
        # every scale wins somewhere (the synthetic radii span 1..6 voxels)
        assert len(np.unique(g["scale"][g["J"] > parity.STRONG_FRAC * g["Jmax"]])) >= 4


def test_periodic_interior_and_global_scalars(oracle, fullsize):
    I, out = fullsize
    g = out["exact"]
    J = g["J"]
    # interior of tile (1,1,1) against tiles (2,1,1), (1,2,2), (2,2,2): identical neighbourhoods
    a = J[BL + HZ:2 * BL - HZ, BH + HXY:2 * BH - HXY, BW + HXY:2 * BW - HXY]
    for dz, dy, dx in ((1, 0, 0), (0, 1, 1), (1, 1, 1)):
        b = J[BL * (1 + dz) + HZ:BL * (2 + dz) - HZ, BH * (1 + dy) + HXY:BH * (2 + dy) - HXY,
              BW * (1 + dx) + HXY:BW * (2 + dx) - HXY]
        assert np.array_equal(a, b), (dz, dy, dx)
    for k in ("Vx", "Vy", "Vz"):
        assert np.array_equal(g[k][BL + HZ:2 * BL - HZ, BH + HXY:2 * BH - HXY, BW + HXY:2 * BW - HXY],
                              g[k][2 * BL + HZ:3 * BL - HZ, 2 * BH + HXY:3 * BH - HXY, 2 * BW + HXY:3 * BW - HXY])
    # global scalars: the maximum is attained, J8 obeys the reference's rule on a large sample
    assert float(J.max()) == g["Jmax"] and float(J.min()) == g["Jmin"] == 0.0
    sl = (slice(200, 264), slice(0, 2048), slice(0, 2048))
    assert np.array_equal(g["J8"][sl], oracle.j_to_j8(np.ascontiguousarray(J[sl]), g["Jmin"], g["Jmax"]))
    assert int(g["J8"].max()) == 255


def test_runs_are_deterministic_and_the_chunked_call_matches():
    """Repeatability at a size where every kernel runs many CTAs per SM: resident runs must reproduce bit for bit,
    and the chunk-pipelined frangi_gpu_run (copies beside kernels, per-chunk launches) must equal them.  This is the
    test that catches a shared-memory slot refilled while a load of it is still in flight (tools/debug_streamed.py):
    such a race shows as a handful of 16-byte quads that differ from run to run."""
    import zlib
    from pnr_b200.synth import make_volume
    w, h, l = 1024, 1024, 96
    base = make_volume(512, 512, 96, seed=17)
    I = np.ascontiguousarray(np.tile(base, (1, 2, 2)))
    p = pnr_b200.FrangiPlan([2.0, 4.0, 6.0], 2.0, .5, .5, 500., False, w, h, l, flags=pnr_b200.FLAG_FMA_SMOOTHING)
    p.upload(I)
    crcs = set()
    for _ in range(6):
        p.run_resident()
        out = p.download(want_J8=True)
        crcs.add(tuple(zlib.crc32(out[k].tobytes()) for k in ("J", "Vx", "Vy", "Vz", "J8")))
    assert len(crcs) == 1, crcs
    for _ in range(3):
        many = p.run(I, want_J8=True)
        assert tuple(zlib.crc32(many[k].tobytes()) for k in ("J", "Vx", "Vy", "Vz", "J8")) in crcs
    p.close()
