#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libpnr_ref.so,
compiled in place from /root/reference by `make -C oracle ref`).

The reference repository ships no tests or golden vectors, so the
fixtures that pin oracle/frangi_oracle.c are outputs of the reference itself on
seeded inputs.  They travel with the repository; the reference does not.

    python tools/make_golden.py        # needs oracle/_ref (only buildable where /root/reference exists)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import Oracle, Reference  # noqa: E402
from pnr_b200.synth import make_volume, straight_tube, volume_hash  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
PARAMS = dict(zdist=2.0, alpha=0.5, beta=0.5, Cc=500.0)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = Reference()
    port = Oracle()   # only for the J -> J8 rule (Advantra_plugin.cpp:2499-2512 cannot be compiled without Qt)

    # ---- case A: small ragged volume, every stage stored in full ---------------------
    I = make_volume(37, 29, 13, seed=101, n_neurites=3)
    sig = [1.0, 2.0, 3.0]
    F = ref.imgaussian(I, 2.0, 2.0)
    D = ref.hessian3d(I, 2.0, 2.0)
    R = ref.frangi3d(I, sig, **PARAMS)
    singles = [ref.frangi3d(I, [s], **PARAMS)["J"] for s in sig]
    np.savez_compressed(os.path.join(OUT, "case_a_stages.npz"), I=I, sigmas=np.float32(sig),
                        F_sigma2=F, **{"H_" + k: v for k, v in D.items()},
                        J=R["J"], Jmin=np.float32(R["Jmin"]), Jmax=np.float32(R["Jmax"]),
                        Vx=R["Vx"], Vy=R["Vy"], Vz=R["Vz"], J_single=np.stack(singles))

    # ---- case B: config-1-like volume with seeds (outputs stored in full, compressed) ---
    I = make_volume(96, 80, 24, seed=1, n_neurites=4)
    sig = [2.0, 4.0, 6.0]
    R = ref.frangi3d(I, sig, **PARAMS)
    J8 = port.j_to_j8(R["J"], R["Jmin"], R["Jmax"])
    seeds = ref.extract_seeds(5.0, J8, R["Vx"], R["Vy"], R["Vz"])
    Rb = ref.frangi3d(255 - I, sig, blackwhite=True, **PARAMS)
    np.savez_compressed(os.path.join(OUT, "case_b_seeds.npz"), I=I, sigmas=np.float32(sig),
                        J=R["J"], Jmin=np.float32(R["Jmin"]), Jmax=np.float32(R["Jmax"]),
                        Vx=R["Vx"], Vy=R["Vy"], Vz=R["Vz"], J8=J8, seeds=seeds,
                        J_blackwhite=Rb["J"])

    # ---- case C: eigen conventions (eigen_decomposition_static) --------------------------
    rng = np.random.default_rng(5)
    mats = [np.zeros((3, 3)), np.diag([-1e-3, -5, -5]), np.diag([-5, -1e-3, -5]), np.diag([-5, -5, -1e-3]),
            np.diag([3.0, -2, 1]), np.diag([2.0, -2, 1]),
            np.array([[-2, .5, .25], [.5, -3, .75], [.25, .75, -.1]]),
            np.array([[-2, -.5, -.25], [-.5, -3, -.75], [-.25, -.75, -.1]])]
    for _ in range(56):
        M = rng.normal(size=(3, 3)) * rng.choice([1e-3, 1.0, 40.0])
        mats.append((M + M.T) / 2)
    A = np.stack(mats)
    V = np.empty_like(A)
    d = np.empty((len(A), 3))
    for i, M in enumerate(A):
        V[i], d[i] = ref.eigen3(M)
    np.savez_compressed(os.path.join(OUT, "case_c_eigen.npz"), A=A, V=V, d=d)

    # ---- case D: SURVEY.md 8c known-answer tube, summary numbers only ---------------------
    T = straight_tube()
    R = ref.frangi3d(T, [2.0, 4.0, 6.0], **PARAMS)
    singles = [ref.frangi3d(T, [s], **PARAMS) for s in (2.0, 4.0, 6.0)]
    S = np.stack([s["J"] for s in singles])
    J8 = port.j_to_j8(R["J"], R["Jmin"], R["Jmax"])
    pos = R["J"] > 0
    summary = dict(
        input_hash=volume_hash(T), input_sum=int(T.sum()),
        single_jmax=[float(s["Jmax"]) for s in singles],
        jmin=float(R["Jmin"]), jmax=float(R["Jmax"]), j_sum=float(R["J"].astype(np.float64).sum()),
        n_positive=int(pos.sum()),
        j_32_34_16=float(R["J"][16, 34, 32]), j_32_32_17=float(R["J"][17, 32, 32]), j_centre=float(R["J"][16, 32, 32]),
        v_centre=[int(R["Vx"][16, 32, 32]), int(R["Vy"][16, 32, 32]), int(R["Vz"][16, 32, 32])],
        v_voxel0=[int(R["Vx"][0, 0, 0]), int(R["Vy"][0, 0, 0]), int(R["Vz"][0, 0, 0])],
        scale_hist=[int(x) for x in np.bincount(S.argmax(0)[pos], minlength=3)],
        j8_sum=int(J8.astype(np.int64).sum()),
        n_seeds=int(len(ref.extract_seeds(5.0, J8, R["Vx"], R["Vy"], R["Vz"]))),
    )
    with open(os.path.join(OUT, "case_d_tube.json"), "w") as fh:
        json.dump(summary, fh, indent=1)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def case_e_2d():
    """case E: the 2-D path (Frangi::frangi2d / hessian2d) on one plane of a synthetic volume and its inverse."""
    ref = Reference()
    I = make_volume(96, 80, 8, seed=21, n_neurites=4)[3]
    sig = [2.0, 3.0, 4.0]
    R = ref.frangi2d(I, sig, 0.5, 15.0)
    Rb = ref.frangi2d(255 - I, sig, 0.5, 15.0, blackwhite=True)
    H = ref.hessian2d(I, 2.0)
    np.savez_compressed(os.path.join(OUT, "case_e_frangi2d.npz"), I=I, sigmas=np.float32(sig), J=R["J"],
                        Jmin=np.float32(R["Jmin"]), Jmax=np.float32(R["Jmax"]), Vx=R["Vx"], Vy=R["Vy"], Vz=R["Vz"],
                        J_blackwhite=Rb["J"], **{"H_" + k: v for k, v in H.items()})


def case_f_soma():
    """case F: the soma helpers (imerode, imdilate, in-place xy imgaussian) on a small volume."""
    ref = Reference()
    I = make_volume(72, 56, 8, seed=31, n_neurites=3)[2:6]
    np.savez_compressed(os.path.join(OUT, "case_f_soma.npz"), I=I, rad=np.float32(3.0), eroded=ref.imerode(I, 3.0),
                        dilated=ref.imdilate(I, 3.0), blurred=ref.imgaussian_xy(I, 3.0),
                        chain=ref.imgaussian_xy(ref.imerode(I, 2.0), 2.0))      # Advantra_plugin.cpp:2432,2438


def case_g_cold():
    """case G: the Frangi members no live code calls (frangi.h:28-31,44,46,51): z-scaled erosion, 2-D smoothing,
    direction tables and their lookup, z interpolation."""
    ref = Reference()
    I = make_volume(56, 40, 10, seed=41, n_neurites=3)
    rng = np.random.default_rng(9)
    t3, t2 = ref.unit_directions(True, 90), ref.unit_directions(False, 30)
    q3 = rng.normal(size=(200, 3)).astype(np.float32)
    q3 /= np.linalg.norm(q3, axis=1, keepdims=True)
    q2 = q3[:, :2] / np.linalg.norm(q3[:, :2], axis=1, keepdims=True)
    F = ref.imgaussian(I, 2.0, 2.0)
    zq = np.float32(rng.uniform(-1.5, I.shape[0] + 0.5, size=64))
    xy = np.stack([rng.integers(0, I.shape[2], 64), rng.integers(0, I.shape[1], 64)], 1).astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "case_g_cold.npz"), I=I, rad=np.float32(3.0), zdist=np.float32(2.0),
                        eroded_z=ref.imerode_z(I, 3.0, 2.0), eroded_z_plane=ref.imerode_z(I[:1], 3.0, 2.0),
                        smooth2d=ref.imgaussian2d(I[4], 2.0), dirs3d=t3, dirs2d=t2, q3=q3, q2=q2,
                        idx3=np.array([ref.direction_idx(v, t3) for v in q3], np.uint8),
                        idx2=np.array([ref.direction_idx(v, t2) for v in q2], np.uint8),
                        F=F, zq=zq, xy=xy,
                        interp=np.array([ref.interpz(int(x), int(y), float(z), F) for (x, y), z in zip(xy, zq)], np.float32),
                        interp_plane=np.float32(ref.interpz(3, 5, 0.7, F[:1])))


def zncc_case(ref, port):
    """The seeds of a small synthetic volume as extractSeeds produces them, plus crafted ones: off-grid positions, random
    unit directions, directions along z (the nrm <= 1e-4 branch of tracker.cpp:1895), seeds at the volume's corners and
    outside it (the clamps of Tracker::interp)."""
    I = make_volume(96, 80, 24, seed=1, n_neurites=4)
    sig = [2.0, 4.0, 6.0]
    R = ref.frangi3d(I, sig, **PARAMS)
    J8 = port.j_to_j8(R["J"], R["Jmin"], R["Jmax"])
    seeds = ref.extract_seeds(5.0, J8, R["Vx"], R["Vy"], R["Vz"])[:, :6]
    rng = np.random.default_rng(12)
    n = 160
    pos = rng.uniform([-3, -3, -2], [99, 83, 26], size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:12] = [0, 0, 1]; d[12:20] = [0, 0, -1]; d[20:26] = [1e-5, -2e-5, 1]; d[26:30] = [0, -1, 0]; d[30:34] = [1, 0, 0]
    pos[:4] = [[0, 0, 0], [95, 79, 23], [0, 79, 0], [95, 0, 23]]
    crafted = np.concatenate([pos, d], 1).astype(np.float32)
    allseeds = np.ascontiguousarray(np.concatenate([seeds, crafted], 0), np.float32)
    return I, sig, allseeds


def case_h_zncc():
    """case H: Tracker::znccBBB, the per-seed score of the plugin's seed filter (Advantra_plugin.cpp:2561-2573)."""
    ref = Reference()
    I, sig, seeds = zncc_case(ref, Oracle())
    corr, best = ref.seed_zncc(I, sig, seeds)
    np.savez_compressed(os.path.join(OUT, "case_h_zncc.npz"), I=I, sigmas=np.float32(sig), seeds=seeds, corr=corr, sig=best)
    print("case H:", len(seeds), "seeds, corr in", float(corr.min()), float(corr.max()), "kept at 0.3:", int((corr >= 0.3).sum()))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"     # "2d" / "soma": only that case (others untouched)
    if which in ("all",):
        main()
    if which in ("all", "2d"):
        case_e_2d()
    if which in ("all", "soma"):
        case_f_soma()
    if which in ("all", "cold"):
        case_g_cold()
    if which in ("all", "zncc"):
        case_h_zncc()
