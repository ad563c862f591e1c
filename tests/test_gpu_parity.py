"""GPU parity tests: the CUDA path (through the C-ABI, via ctypes) against the
CPU oracle on the same seeded inputs.  Tolerances are those of
BASELINE.json:north_star and are written in tests/parity.py."""
import numpy as np
import pytest

import pnr_b200
from pnr_b200 import FLAG_DIR_F32, FLAG_FMA_SMOOTHING, FLAG_SCALE_IDX, Frangi
from pnr_b200.synth import make_volume, straight_tube
from tests import parity

pytestmark = pytest.mark.gpu

SIGS = [2.0, 4.0, 6.0]


def _frangi(sigs=SIGS, zdist=2.0, flags=FLAG_DIR_F32 | FLAG_SCALE_IDX, **kw):
    return Frangi(sigs, zdist, 0.5, 0.5, 500.0, 0.5, 15.0, flags=flags, **kw)


# ---------------------------------------------------------------- stage: smoothing
@pytest.mark.parametrize("shape,sigma,zdist", [
    ((24, 80, 96), 2.0, 2.0),
    ((24, 80, 96), 6.0, 2.0),
    ((11, 29, 37), 1.0, 2.0),     # ragged, narrower than one strip
    ((11, 29, 37), 4.0, 1.0),     # radius 12 > every dimension's half
    ((5, 300, 515), 3.0, 2.0),    # w not a multiple of 4: scalar load path, 3 strips
    ((2, 2, 2), 2.0, 2.0),        # smallest legal volume
    ((40, 33, 260), 2.5, 2.0),    # non-integer sigma: radius 8 padded to template 9
])
def test_imgaussian_bit_exact(oracle, shape, sigma, zdist):
    l, h, w = shape
    I = make_volume(w, h, l, seed=7, n_neurites=4)
    F = Frangi.imgaussian(I, sigma, zdist)
    R = oracle.imgaussian(I, sigma, zdist)
    assert np.array_equal(F, R), f"max diff {np.abs(F - R).max()}"


def test_imgaussian_fma_mode_close(oracle):
    I = make_volume(96, 80, 24, seed=7, n_neurites=4)
    F = Frangi.imgaussian(I, 4.0, 2.0, flags=FLAG_FMA_SMOOTHING)
    R = oracle.imgaussian(I, 4.0, 2.0)
    assert np.abs(F - R).max() <= 2e-5 * max(1.0, R.max())


# ---------------------------------------------------------------- stage: second differences
@pytest.mark.parametrize("shape,sigma", [((24, 80, 96), 2.0), ((11, 29, 37), 4.0), ((2, 2, 2), 1.0),
                                         ((3, 5, 4), 2.0)])
def test_hessian3d_bit_exact(oracle, shape, sigma):
    l, h, w = shape
    I = make_volume(w, h, l, seed=11, n_neurites=4)
    D = _frangi().hessian3d(I, sigma, 2.0)
    R = oracle.hessian3d(I, sigma, 2.0)
    for k in R:
        assert np.array_equal(D[k], R[k]), f"{k}: max diff {np.abs(D[k] - R[k]).max()}"


# ---------------------------------------------------------------- stage: eigen + vesselness
@pytest.mark.parametrize("scalar", [False, True])
def test_vesselness_stage_random_and_degenerate(oracle, scalar):
    rng = np.random.default_rng(3)
    n = 200000
    M = rng.normal(size=(n, 3, 3)).astype(np.float32) * rng.choice([1e-2, 1.0, 50.0], size=(n, 1, 1)).astype(np.float32)
    A = (M + M.transpose(0, 2, 1)) / 2
    # tube-like: two close negative eigenvalues, one near zero, random orientation
    Q, _ = np.linalg.qr(rng.normal(size=(n // 2, 3, 3)))
    lam = np.stack([rng.normal(scale=0.05, size=n // 2), -1 - rng.random(n // 2) * 1e-3 * rng.choice([0, 1, 100], n // 2),
                    -np.ones(n // 2)], 1) * rng.choice([1.0, 20.0], size=(n // 2, 1))
    T = np.einsum("nij,nj,nkj->nik", Q, lam, Q).astype(np.float32)
    A[: n // 2] = T
    A[-1000:] = 0                                   # zero Hessians (flat background)
    A[-2000:-1000, 0, 1] = A[-2000:-1000, 1, 0] = 0  # partially diagonal
    A[-3000:-2000] *= np.eye(3, dtype=np.float32)    # diagonal
    D = dict(Dxx=A[:, 0, 0], Dxy=A[:, 0, 1], Dxz=A[:, 0, 2], Dyy=A[:, 1, 1], Dyz=A[:, 1, 2], Dzz=A[:, 2, 2])
    D = {k: np.ascontiguousarray(v) for k, v in D.items()}
    f = _frangi()
    v, d, lam_g = f.vesselness_stage(D, scalar=scalar)
    vr, dr, lam_r = oracle.vesselness_stage(D, want_lambda=True)
    # eigenvalues: absolute error relative to the spectral norm
    scale = np.abs(lam_r).max(1)
    err = np.abs(lam_g.astype(np.float64) - lam_r).max(1)
    ok = err <= 4e-6 * np.maximum(scale, 1e-30)
    assert ok.mean() > 0.9999, f"eigenvalue errors: worst {np.max(err / np.maximum(scale, 1e-30))}"
    rep = parity.vesselness_report(v, vr)
    # random matrices sit on the discontinuous sign gate more often than image data
    assert rep["n_bad"] <= 5, rep
    # direction where well conditioned (gap between |l1| and |l2| at least 1 % of |l3|)
    gap = (np.abs(lam_r[:, 1]) - np.abs(lam_r[:, 0])) / np.maximum(np.abs(lam_r[:, 2]), 1e-30)
    well = (gap > 1e-2) & (scale > 0)
    dot = np.abs((d.astype(np.float64) * dr).sum(0))
    ang = np.degrees(np.arccos(np.clip(dot, 0, 1)))
    assert ang[well].max() < 0.05, ang[well].max()
    # zero matrix conventions: v = 0, direction (1,0,0)
    assert np.all(v[-1000:] == 0) and np.all(d[0, -1000:] == 1) and np.all(d[1:, -1000:] == 0)


# ---------------------------------------------------------------- whole path
def _full_check(oracle, I, sigs=SIGS, zdist=2.0, blackwhite=False, flags=FLAG_DIR_F32 | FLAG_SCALE_IDX,
                seeds=True):
    f = _frangi(sigs, zdist, flags=flags)
    f.blackwhite = blackwhite
    g = f.frangi3d_full(I)
    r = oracle.frangi3d(I, sigs, zdist, blackwhite=blackwhite)
    rep = parity.vesselness_report(g["J"], r["J"])
    assert rep["n_bad"] == 0, rep
    assert abs(g["Jmax"] - r["Jmax"]) <= max(1e-4 * r["Jmax"], 1e-6)
    assert abs(g["Jmin"] - r["Jmin"]) <= 1e-6
    if r["Jmax"] > 0:
        drep = parity.direction_report(g["dir"], r["dir"], r["J"])
        assert drep["n_bad"] == 0, drep
        crep = parity.code_report((g["Vx"], g["Vy"], g["Vz"]), (r["Vx"], r["Vy"], r["Vz"]), r["J"])
        assert crep["n_bad"] == 0, crep
        singles = [oracle.frangi3d(I, [s], zdist, blackwhite=blackwhite, want_scale=False, want_dir=False)["J"]
                   for s in sigs]
        srep = parity.scale_report(g["scale"], singles)
        assert srep["n_bad"] == 0, srep
    # J8 of the GPU (computed on device from its own J, Jmin, Jmax) vs the oracle's rule
    j8_rule = oracle.j_to_j8(g["J"], g["Jmin"], g["Jmax"])
    assert np.array_equal(g["J8"], j8_rule)
    if seeds and r["Jmax"] > 0:
        j8_ref = oracle.j_to_j8(r["J"], r["Jmin"], r["Jmax"])
        s_g = oracle.extract_seeds(5.0, g["J8"], g["Vx"], g["Vy"], g["Vz"])
        s_r = oracle.extract_seeds(5.0, j8_ref, r["Vx"], r["Vy"], r["Vz"])
        seedrep = parity.seed_report(s_g, s_r)
        assert seedrep["match"] >= parity.SEED_MATCH, seedrep
        assert seedrep["bad_dir"] <= max(1, seedrep["n_common"] // 1000), seedrep
    f.close()
    return g, r


def test_frangi3d_known_answer_tube(oracle):
    """x-aligned Gaussian tube, values recorded from the reference."""
    T = straight_tube()
    g, r = _full_check(oracle, T, seeds=False)
    assert g["Jmin"] == 0.0
    assert abs(g["Jmax"] - 0.0076699215) < 1e-7
    assert abs(float(g["J"][16, 32, 32]) - 0.0076699215) < 1e-7
    assert abs(float(g["J"][16, 34, 32]) - 0.00469563901) < 1e-7
    assert abs(float(g["J"][17, 32, 32]) - 0.00658458145) < 1e-7
    assert (g["Vx"][16, 32, 32], g["Vy"][16, 32, 32], g["Vz"][16, 32, 32]) in ((255, 128, 128), (0, 127, 127), (0, 128, 128))
    assert int((g["J"] > 0).sum()) == 4800
    hist = np.bincount(g["scale"][g["J"] > 0], minlength=3)
    assert list(hist) == [0, 1472, 3328]


def test_frangi3d_config1_256x256x64(oracle):
    """BASELINE.json configs[0]: sigma = 2,4,6 on a synthetic 256x256x64 neuron volume."""
    I = make_volume(256, 256, 64)
    _full_check(oracle, I)


@pytest.mark.parametrize("shape,sigs,zdist,bw", [
    ((11, 29, 37), [1.0, 2.0], 2.0, False),            # ragged
    ((9, 20, 20), [2.0, 4.0, 6.0], 2.0, False),        # every radius larger than the volume
    ((24, 80, 96), [1.0, 2.0, 3.0, 4.0, 5.0, 6.0], 2.0, False),  # config-3 scale set
    ((24, 80, 96), [2.0, 4.0], 1.0, False),            # isotropic z
    ((24, 80, 96), [2.0, 4.0, 6.0], 2.0, True),        # dark ridges
    ((16, 64, 64), [3.0], 2.0, False),                 # single scale
    ((2, 2, 2), [2.0], 2.0, False),                    # smallest legal volume
    ((10, 20, 257), [2.0, 3.0], 2.0, False),           # (w - 4) mod 124 = 5: the streaming kernel leaves its last five columns to the shell
    ((8, 12, 132), [1.0, 2.0], 2.0, False),            # one tile column + a four-column remainder
    ((5, 24, 140), [1.0, 2.0], 2.0, False),            # every plane but one lies next to a z face (the z-face form of the tile kernels)
])
def test_frangi3d_edge_cases(oracle, shape, sigs, zdist, bw):
    l, h, w = shape
    I = make_volume(w, h, l, seed=5, n_neurites=5)
    if bw:
        I = 255 - I
    _full_check(oracle, I, sigs, zdist, bw, seeds=False)


def test_constant_and_empty_volumes(oracle):
    for val in (0, 255, 17):
        I = np.full((8, 16, 16), val, np.uint8)
        g, r = _full_check(oracle, I, seeds=False)
        assert g["Jmax"] == 0.0 and g["Jmin"] == 0.0
        assert np.all(g["J8"] == 0)
        assert np.all(g["Vx"] == 255) and np.all(g["Vy"] == 128) and np.all(g["Vz"] == 128)


def test_fma_mode_within_tolerance(oracle):
    I = make_volume(128, 96, 32, seed=9, n_neurites=8)
    _full_check(oracle, I, flags=FLAG_DIR_F32 | FLAG_SCALE_IDX | FLAG_FMA_SMOOTHING)


def test_bad_arguments_fail_loudly():
    from pnr_b200 import FrangiGpuError, FrangiPlan
    with pytest.raises(FrangiGpuError):
        FrangiPlan([2.0], 2.0, .5, .5, 500., False, 16, 16, 1)       # 2-D image: out of scope
    with pytest.raises(FrangiGpuError):
        FrangiPlan([], 2.0, .5, .5, 500., False, 16, 16, 16)
    with pytest.raises(FrangiGpuError):
        FrangiPlan([50.0], 2.0, .5, .5, 500., False, 16, 16, 16)    # radius beyond the largest instantiation
    with pytest.raises(FrangiGpuError):
        FrangiPlan([2.0], 0.0, .5, .5, 500., False, 16, 16, 16)


def test_repeatable_and_handle_reuse(oracle):
    I1 = make_volume(96, 80, 24, seed=1, n_neurites=4)
    I2 = make_volume(96, 80, 24, seed=2, n_neurites=4)
    f = _frangi()
    a = f.frangi3d_full(I1); b = f.frangi3d_full(I2); c = f.frangi3d_full(I1)
    for k in ("J", "Vx", "Vy", "Vz", "J8", "scale"):
        assert np.array_equal(a[k], c[k]), k
    assert not np.array_equal(a["J"], b["J"])


def test_overlapped_schedule_is_bit_identical():
    """FRANGI_GPU_FLAG_OVERLAP_Z: z pass on a second stream beside its neighbours, two buffer pairs -- same bits."""
    from pnr_b200.frangi import FLAG_OVERLAP_Z
    I = make_volume(160, 128, 72, seed=9)
    outs = []
    for flags in (0, FLAG_OVERLAP_Z):
        p = pnr_b200.FrangiPlan(SIGS, 2.0, .5, .5, 500., False, 160, 128, 72, flags=flags)
        p.set_stream_chunk(0)            # one-piece run: the schedule under test
        for _ in range(2):               # twice: buffer reuse across runs
            o = p.run(I, want_J8=True)
        outs.append(o)
        p.close()
    a, b = outs
    assert a["Jmax"] == b["Jmax"] and a["Jmin"] == b["Jmin"]
    for k in ("J", "Vx", "Vy", "Vz", "J8"):
        assert np.array_equal(a[k], b[k]), k


def test_back_to_back_asynchronous_runs_do_not_share_jmin_jmax():
    """frangi_gpu_run_device with Jmin = Jmax = NULL does not synchronise.  Two different volumes launched back to
    back must each start from Jmin = FLT_MAX, Jmax = -FLT_MAX (frangi.cpp:176-177): the pair is initialised on the
    device, so the first run's result copy can never feed the second run's start value.  The second volume is a
    dimmed copy of the first (smaller Jmax): a leaked Jmax would change every J8 code."""
    import torch
    w, h, l = 160, 128, 40
    A = make_volume(w, h, l, seed=11, n_neurites=6)
    B = (make_volume(w, h, l, seed=12, n_neurites=6) // 3).astype(np.uint8)
    p = pnr_b200.FrangiPlan(SIGS, 2.0, .5, .5, 500., False, w, h, l)
    want = {}
    for name, I in (("A", A), ("B", B)):
        p.upload(I)
        lo, hi = p.run_resident()
        want[name] = dict(p.download(want_J8=True), Jmin=lo, Jmax=hi)
    assert want["B"]["Jmax"] < 0.5 * want["A"]["Jmax"]
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    torch.cuda.synchronize()
    for _ in range(5):                       # no host synchronisation between the two runs
        p.run_device(dA.data_ptr(), sync=False)
        p.run_device(dB.data_ptr(), sync=False)
    p.sync()
    got = p.download(want_J8=True)
    for k in ("J", "J8", "Vx", "Vy", "Vz"):
        assert np.array_equal(got[k], want["B"][k]), k
    lo, hi = p.run_device(dA.data_ptr())     # and the synchronous form after the asynchronous ones
    assert (lo, hi) == (want["A"]["Jmin"], want["A"]["Jmax"])
    p.close()
