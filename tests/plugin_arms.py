"""TEST INFRASTRUCTURE.  Runs one arm of the whole-plugin comparison (oracle.Plugin: the reference's unmodified
Advantra_plugin.cpp, either with the reference's Frangi or with the drop-in class) in a CHILD process, so that the two
libraries -- which define the same C++ symbols -- never share a process and a crash in the tracker cannot take pytest
down.  Used by tests/test_plugin_e2e.py and tools/plugin_e2e.py.

    python -m tests.plugin_arms <arm> <vol.npy> <workdir> <max_traces> <p1> ... <p11>
"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_arm(arm, I, params, max_traces, workdir=None, timeout=3600, replay=None, env=None):
    """-> dict(files={suffix: text}, seconds=wall time of dofunc, log=the plugin's stdout).
    replay = (J8, Vx, Vy, Vz) for arm "replay": the filter outputs the plugin's call site is handed.
    env: extra environment of the child, e.g. PNR_FRANGI_FLAGS for the drop-in class (pnr_b200/csrc/frangi_shim.cpp)."""
    workdir = workdir or tempfile.mkdtemp(prefix=f"pnr_plugin_{arm}_")
    vol = os.path.join(workdir, "_vol.npy")
    np.save(vol, np.ascontiguousarray(I, np.uint8))
    if replay is not None:
        np.save(os.path.join(workdir, "_replay.npy"), np.stack([np.ascontiguousarray(v, np.uint8) for v in replay]))
    cmd = [sys.executable, "-m", "tests.plugin_arms", arm, vol, workdir, str(int(max_traces))] + [str(p) for p in params]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, errors="replace", timeout=timeout,
                         env=dict(os.environ, **env) if env else None)
    os.remove(vol)
    if replay is not None:
        os.remove(os.path.join(workdir, "_replay.npy"))
    if out.returncode != 0:
        raise RuntimeError(f"plugin arm {arm!r} failed (rc {out.returncode}): {out.stderr[-2000:]}")
    with open(os.path.join(workdir, "_result.json")) as f:
        res = json.load(f)
    os.remove(os.path.join(workdir, "_result.json"))
    res["log"] = out.stdout
    return res


def plugin_j8(J, Jmin, Jmax):
    """The plugin's own 8-bit form of the vesselness (Advantra_plugin.cpp:2499-2512), float arithmetic."""
    J = np.asarray(J, np.float32)
    if abs(float(Jmax) - float(Jmin)) <= np.finfo(np.float32).tiny:
        return np.zeros(J.shape, np.uint8)
    v = (J - np.float32(Jmin)) / (np.float32(Jmax) - np.float32(Jmin)) * np.float32(255)
    return np.clip(np.floor(np.abs(v) + np.float32(0.5)) * np.sign(v), 0, 255).astype(np.uint8)


def load_capture(path, r):
    """The GPU path's outputs (J8, Vx, Vy, Vz) for the volume of tests/test_plugin_e2e.py, rebuilt from the reference's
    outputs `r` on that volume and the differences tools/capture_gpu_outputs.py stored: which eigenvector signs the GPU
    chose differently (one bit per voxel), the few voxels whose direction code differs beyond the sign, the J8 voxels
    that differ.  -> dict(J8, Vx, Vy, Vz, flip, signs_only=(J8, Vx, Vy, Vz) of the reference with the GPU's signs and
    nothing else changed, beyond_sign={array: count})."""
    z = np.load(path)
    shape = r["Vx"].shape
    n = int(np.prod(shape))
    flip = np.unpackbits(z["flip"])[:n].astype(bool).reshape(shape)
    j8 = plugin_j8(r["J"], r["Jmin"], r["Jmax"])
    signs_only = [j8] + [np.where(flip, 255 - r[k], r[k]).astype(np.uint8) for k in ("Vx", "Vy", "Vz")]
    out = dict(flip=flip, signs_only=tuple(signs_only), beyond_sign={}, input_hash=str(z["input_hash"]))
    for k, base in zip(("J8", "Vx", "Vy", "Vz"), signs_only):
        a = base.copy().ravel()
        a[z[k + "_idx"]] = z[k + "_val"]
        out[k] = a.reshape(shape)
        out["beyond_sign"][k] = int(len(z[k + "_idx"]))
    return out


def swc_rows(text):
    """The numeric rows of an SWC file: [n, 7] float64 (n type x y z r parent)."""
    rows = [ln.split() for ln in text.splitlines() if ln and not ln.startswith("#")]
    return np.array(rows, np.float64).reshape(-1, 7) if rows else np.zeros((0, 7))


def _pairs(rows):
    """_Seeds.swc / _VxVyVz.swc: a locus row (parent -1) followed by a tip row = locus + length * direction."""
    loc, tip = rows[0::2], rows[1::2]
    assert len(loc) == len(tip) and np.all(loc[:, 6] == -1) and np.all(tip[:, 6] == loc[:, 0])
    return loc, tip[:, 2:5] - loc[:, 2:5]


def _pair_report(ra, rb):
    """Loci as a set and in order; directions MODULO SIGN (the sign of an eigenvector is arbitrary: the reference
    writes whatever its QL iteration returns, frangi.cpp:239-250)."""
    (la, da), (lb, db) = _pairs(ra), _pairs(rb)
    ka = {tuple(x[2:5]): d for x, d in zip(la, da)}
    kb = {tuple(x[2:5]): d for x, d in zip(lb, db)}
    common = set(ka) & set(kb)
    flipped = bad = 0
    for k in common:
        c = float(np.dot(ka[k], kb[k])) / max(1e-30, float(np.linalg.norm(ka[k]) * np.linalg.norm(kb[k])))
        flipped += c < 0
        bad += np.degrees(np.arccos(min(1.0, abs(c)))) > 1.0
    same_order = len(la) == len(lb) and bool(np.array_equal(la[:, 2:5], lb[:, 2:5]))
    first = None
    if not same_order:
        m = min(len(la), len(lb))
        d = np.where(np.any(la[:m, 2:5] != lb[:m, 2:5], axis=1))[0]
        first = int(d[0]) if len(d) else m
    return dict(loci=[len(ka), len(kb)], loci_match=len(common) / max(1, len(set(ka) | set(kb))), same_order=same_order,
                first_order_difference=first, sign_flipped=int(flipped), direction_over_1deg=int(bad),
                score_max_abs_diff=float(np.max(np.abs(la[:, 5] - lb[:, 5]))) if same_order else None)


def _near(pa, pb, radius):
    """Fraction of the points of pa that have a point of pb within `radius` voxels."""
    from scipy.spatial import cKDTree
    if len(pa) == 0:
        return 1.0
    if len(pb) == 0:
        return 0.0
    d, _ = cKDTree(pb).query(pa, k=1)
    return float(np.mean(d <= radius))


def compare_files(a, b):
    """Per file: identical text; for SWC node files the fraction of node positions present in both (3 decimals) and the
    fraction of nodes with a node of the other file within half a voxel (the smaller of the two directions); for the
    seed list and the direction dump the sign-aware report of _pair_report."""
    rep = {}
    for k in sorted(set(a) | set(b)):
        if k not in a or k not in b:
            rep[k] = dict(identical=False, missing_in="ref" if k not in b else "gpu")
            continue
        if not isinstance(a[k], str) or not isinstance(b[k], str):      # packed by tools/plugin_e2e.py: hash and row count
            rep[k] = dict(identical=a[k] == b[k])
            continue
        r = dict(identical=a[k] == b[k], bytes=len(a[k]))
        if k in ("_Seeds.swc", "_VxVyVz.swc"):
            if not r["identical"]:
                r.update(_pair_report(swc_rows(a[k]), swc_rows(b[k])))
        elif k.endswith(".swc"):
            ra, rb = swc_rows(a[k]), swc_rows(b[k])
            sa = {tuple(np.round(x[2:5], 3)) for x in ra}
            sb = {tuple(np.round(x[2:5], 3)) for x in rb}
            r.update(nodes=[len(ra), len(rb)], position_match=len(sa & sb) / max(1, len(sa | sb)),
                     within_half_voxel=min(_near(ra[:, 2:5], rb[:, 2:5], 0.5), _near(rb[:, 2:5], ra[:, 2:5], 0.5)),
                     within_two_voxels=min(_near(ra[:, 2:5], rb[:, 2:5], 2.0), _near(rb[:, 2:5], ra[:, 2:5], 2.0)))
        rep[k] = r
    return rep


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    from oracle import Plugin
    arm, vol, workdir, max_traces = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
    I = np.load(vol)
    p = Plugin(arm)
    t0 = time.time()
    rp = os.path.join(workdir, "_replay.npy")
    files = p.run(I, workdir, params=sys.argv[5:], max_traces=max_traces,
                  single_tree=os.environ.get("PNR_PLUGIN_SINGLE_TREE", "1") != "0",
                  replay=tuple(np.load(rp)) if arm == "replay" else None)
    dt = time.time() - t0
    with open(os.path.join(workdir, "_result.json"), "w") as f:
        json.dump(dict(files=files, seconds=dt), f)
