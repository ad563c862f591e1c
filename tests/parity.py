"""Comparison helpers for the parity tests (tolerances are BASELINE.json:north_star's).

  vesselness : |dJ| <= max(1e-4 * |J_ref|, 1e-6)
  direction  : <= 0.5 degrees, sign-invariant, where J_ref > 1 % of Jmax
  argmax scale : exact away from ties between the two best single-scale responses
  seeds      : >= 99.9 % identical positions (directions compared modulo sign)
"""
from __future__ import annotations

import numpy as np

J_RTOL = 1e-4
J_ATOL = 1e-6
DIR_DEG = 0.5
STRONG_FRAC = 0.01
SEED_MATCH = 0.999


def vesselness_report(J, J_ref):
    J = np.asarray(J, np.float64)
    R = np.asarray(J_ref, np.float64)
    err = np.abs(J - R)
    tol = np.maximum(J_RTOL * np.abs(R), J_ATOL)
    bad = err > tol
    strong = R > STRONG_FRAC * R.max() if R.max() > 0 else np.zeros(R.shape, bool)
    rel = err[strong] / R[strong] if strong.any() else np.zeros(1)
    return dict(n=int(J.size), n_bad=int(bad.sum()), max_abs=float(err.max()),
                max_rel_strong=float(rel.max()), median_rel_strong=float(np.median(rel)),
                worst_ratio=float((err / tol).max()), bitexact=int((J == R).sum()))


def decode_dir(Vx, Vy, Vz):
    """u8 codes -> unit vectors, as the consumer does (seed.cpp:767-771)."""
    v = np.stack([Vx, Vy, Vz]).astype(np.float64) / 255.0 * 2.0 - 1.0
    n = np.sqrt((v * v).sum(0))
    n[n == 0] = 1.0
    return v / n


def direction_report(d, d_ref, J_ref, max_deg=DIR_DEG):
    """d, d_ref: [3, ...] float vectors (not necessarily unit)."""
    d = np.asarray(d, np.float64)
    r = np.asarray(d_ref, np.float64)
    strong = J_ref > STRONG_FRAC * J_ref.max()
    dn = np.sqrt((d * d).sum(0)); rn = np.sqrt((r * r).sum(0))
    dn[dn == 0] = 1; rn[rn == 0] = 1
    dot = np.abs((d * r).sum(0) / (dn * rn))
    ang = np.degrees(np.arccos(np.clip(dot, 0, 1)))
    a = ang[strong]
    return dict(n_strong=int(strong.sum()), n_bad=int((a > max_deg).sum()),
                max_deg=float(a.max()) if a.size else 0.0,
                p999_deg=float(np.quantile(a, 0.999)) if a.size else 0.0)


def code_report(V, V_ref, J_ref):
    """u8 direction codes modulo sign: |c - c_ref| <= 1 on all three channels, or
    |c - (255 - c_ref)| <= 1 on all three, where J_ref is strong."""
    strong = J_ref > STRONG_FRAC * J_ref.max()
    a = np.stack(V).astype(np.int32)[:, strong]
    b = np.stack(V_ref).astype(np.int32)[:, strong]
    same = (np.abs(a - b) <= 1).all(0)
    flip = (np.abs(a - (255 - b)) <= 1).all(0)
    ok = same | flip
    return dict(n_strong=int(strong.sum()), n_bad=int((~ok).sum()), n_flipped=int((flip & ~same).sum()))


def scale_report(scale, J_single, tol_r=J_RTOL, tol_a=J_ATOL):
    """scale: GPU argmax index; J_single: list of single-sigma reference responses.
    The oracle arg-max is first-wins; a GPU index may differ only where the best
    two responses are within tolerance of each other."""
    S = np.stack([np.asarray(j, np.float64) for j in J_single])
    ref = S.argmax(0)                      # first maximum wins, like the strict '>' update
    diff = scale != ref
    best = S.max(0)
    chosen = np.take_along_axis(S, scale[None].astype(np.int64), 0)[0]
    tie = np.abs(best - chosen) <= np.maximum(tol_r * best, tol_a)
    return dict(n=int(scale.size), n_diff=int(diff.sum()), n_bad=int((diff & ~tie).sum()))


def seed_report(seeds, seeds_ref, dir_deg=1.0):
    """seeds: rows (x,y,z,vx,vy,vz).  Position sets must agree; directions modulo sign."""
    key = lambda s: {(int(r[0]), int(r[1]), int(r[2])): r[3:6] for r in s}
    a, b = key(seeds), key(seeds_ref)
    common = set(a) & set(b)
    union = set(a) | set(b)
    frac = len(common) / max(1, len(union))
    bad_dir = 0
    for k in common:
        dot = abs(float(np.dot(a[k], b[k])))
        if np.degrees(np.arccos(min(1.0, dot))) > dir_deg:
            bad_dir += 1
    return dict(n=len(a), n_ref=len(b), n_common=len(common), match=frac, bad_dir=bad_dir)


def direction_gap(oracle, crop, sigmas, zdist, voxel, scale_index):
    """How well the reference itself determines the direction at `voxel` (z, y, x of `crop`) at scale `scale_index`:
    (|l2| - |l1|) / |l3| of the reference's eigenvalues of the reference's Hessian there.  The direction written out
    is the eigenvector of the smallest-magnitude eigenvalue (frangi.cpp:240-250); when |l1| ~ |l2| a last-bit change
    of the Hessian turns it freely inside the plane of the two (a float-vs-double rebuild
    of the reference flips such voxels too)."""
    D = oracle.hessian3d(np.ascontiguousarray(crop), float(sigmas[scale_index]), zdist)
    z, y, x = voxel
    A = np.array([[D["Dxx"][z, y, x], D["Dxy"][z, y, x], D["Dxz"][z, y, x]],
                  [D["Dxy"][z, y, x], D["Dyy"][z, y, x], D["Dyz"][z, y, x]],
                  [D["Dxz"][z, y, x], D["Dyz"][z, y, x], D["Dzz"][z, y, x]]], np.float64)
    _, d = oracle.eigen3(A)
    m = np.abs(d)
    return float((m[1] - m[0]) / max(m[2], 1e-300))


ILL_CONDITIONED_GAP = 2e-3     # (|l2| - |l1|) / |l3| below this: the direction is not determined to 0.5 degrees
