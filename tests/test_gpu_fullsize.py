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


def _crop_check(oracle, I, g, x0, x1, y0, y1, z0, z1):
    xa, xb = max(x0 - HXY, 0), min(x1 + HXY, W)
    ya, yb = max(y0 - HXY, 0), min(y1 + HXY, H)
    za, zb = max(z0 - HZ, 0), min(z1 + HZ, L)
    # a crop edge that is not a volume face would be treated as one by the oracle: only valid
    # because the halo keeps it HXY / HZ away from the region that is compared
    ref = oracle.frangi3d(np.ascontiguousarray(I[za:zb, ya:yb, xa:xb]), SIGS, 2.0, want_scale=False, want_dir=False)
    sl = (slice(z0 - za, z1 - za), slice(y0 - ya, y1 - ya), slice(x0 - xa, x1 - xa))
    gs = (slice(z0, z1), slice(y0, y1), slice(x0, x1))
    return ref["J"][sl], [ref[k][sl] for k in ("Vx", "Vy", "Vz")], g["J"][gs], [g[k][gs] for k in ("Vx", "Vy", "Vz")]


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
    jmax = g["Jmax"]
    assert jmax > 0 and g["Jmin"] == 0.0
    for c in CROPS:
        Jr, Vr, Jg, Vg = _crop_check(oracle, I, g, *c)
        err = np.abs(Jg.astype(np.float64) - Jr)
        tol = np.maximum(parity.J_RTOL * np.abs(Jr), parity.J_ATOL)
        assert (err <= tol).all(), (c, float(err.max()))
        strong = Jr > parity.STRONG_FRAC * jmax
        if strong.any():
            a = np.stack(Vg).astype(np.int32)[:, strong]
            b = np.stack(Vr).astype(np.int32)[:, strong]
            ok = (np.abs(a - b) <= 1).all(0) | (np.abs(a - (255 - b)) <= 1).all(0)
            # With bit-identical smoothing every strong voxel must agree.  FMA smoothing moves the
            # Hessian in its last bits, which can turn the direction where |l1| ~ |l2| makes it
            # ill-conditioned (1 voxel in 21 000 here, its response agreeing to 2e-6 relative):
            # at most one such voxel per 5 000 strong ones is tolerated in that mode.
            allowed = 0 if mode == "exact" else max(1, int(strong.sum()) // 5000)
            assert int((~ok).sum()) <= allowed, (c, int((~ok).sum()), int(strong.sum()))


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
