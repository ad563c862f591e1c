#!/usr/bin/env python
"""BASELINE.json configs[4] at its stated size: GPU Frangi -> the reference's seed.cpp -> the reference's tracker.cpp
on a synthetic 1024x1024x256 neuron, against the same chain fed by the UNMODIFIED reference Frangi.

The reference Frangi cannot be run in one piece at a useful speed (one thread, ~1.2e6 voxel/s: 4 minutes), so it is run
the way SURVEY.md 8c prescribes for big volumes: overlapping z-slabs [z0-11, z1+11), one process per host core, planes
[z0, z1) kept (exact: every stage is z-local with radius ceil(3*6/2)+2 = 11).  Everything downstream of the filter is
the unmodified reference compiled into oracle/_ref (ref_trace restates the plugin's one call site,
Advantra_plugin.cpp:2416-2719, README parameters `2,4,6 0 5 0.3 3 2 200 20 2 4 1`; srand(time(NULL)) pinned).
Both arms trace in parallel processes.  Prints one JSON report (also written to gpurun_out/).

    python tools/config5_at_size.py [--traces 200] [--size 1024x1024x256]

A tool, not a pytest test: ~6 minutes of 16 host cores.  oracle/ is used here as the checker only.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SIGS = [2.0, 4.0, 6.0]
HALO = 11
_G = {}


def _slab(args):
    z0, z1 = args
    from oracle import Reference
    I = _G["I"]
    za, zb = max(z0 - HALO, 0), min(z1 + HALO, I.shape[0])
    r = Reference().frangi3d(np.ascontiguousarray(I[za:zb]), SIGS)
    sl = slice(z0 - za, z1 - za)
    return z0, z1, r["J"][sl].copy(), r["Vx"][sl].copy(), r["Vy"][sl].copy(), r["Vz"][sl].copy()


def _trace(args):
    name, J8, Vx, Vy, Vz, traces = args
    from oracle import Reference
    t0 = time.time()
    tr = Reference().trace(_G["I"], J8, Vx, Vy, Vz, SIGS, max_traces=traces)
    tr["seconds"] = time.time() - t0
    return name, tr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--traces", type=int, default=200)
    ap.add_argument("--size", default="1024x1024x256")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_config5_at_size.json"))
    a = ap.parse_args()
    w, h, l = (int(x) for x in a.size.split("x"))
    import pnr_b200
    from oracle import Oracle, Reference
    from pnr_b200.synth import make_volume, volume_hash
    from tests import parity
    if not Reference.available() or not Reference().has_trace:
        raise SystemExit("needs oracle/_ref/libpnr_ref.so with ref_trace (make -C oracle ref where /root/reference exists)")
    rep = {"size": [w, h, l], "sigmas": SIGS, "traces": a.traces}
    t0 = time.time()
    I = make_volume(w, h, l, seed=20181009 + 4)
    _G["I"] = I
    rep["input"] = {"hash": volume_hash(I), "seconds": time.time() - t0}

    # ---- arm 1: the GPU filter (both smoothing modes), through the reference-facing class
    gpu = {}
    for name, flags in (("exact", 0), ("fma", pnr_b200.FLAG_FMA_SMOOTHING)):
        f = pnr_b200.Frangi(SIGS, 2.0, 0.5, 0.5, 500.0, flags=flags)
        t0 = time.time()
        gpu[name] = f.frangi3d_full(I, want_J8=True)
        gpu[name]["seconds"] = time.time() - t0
        f.close()

    # ---- arm 2: the unmodified reference filter on overlapping z-slabs, one process per core
    cores = len(os.sched_getaffinity(0))
    nsl = max(cores, 1)
    bounds = [(l * k // nsl, l * (k + 1) // nsl) for k in range(nsl)]
    t0 = time.time()
    with mp.get_context("fork").Pool(cores) as pool:
        parts = pool.map(_slab, bounds, chunksize=1)
    Jr = np.empty(I.shape, np.float32)
    Vr = [np.empty(I.shape, np.uint8) for _ in range(3)]
    for z0, z1, J, vx, vy, vz in parts:
        Jr[z0:z1] = J
        Vr[0][z0:z1], Vr[1][z0:z1], Vr[2][z0:z1] = vx, vy, vz
    jmin_r, jmax_r = 0.0, float(Jr.max())        # Jmin: every background voxel is 0 at scale 0
    assert float(Jr.min()) == 0.0
    rep["reference_frangi"] = {"seconds": time.time() - t0, "cores": cores, "slabs": nsl, "halo": HALO, "jmax": jmax_r}
    port = Oracle()
    J8r = port.j_to_j8(Jr, jmin_r, jmax_r)

    # ---- filter parity over the whole volume
    for name, g in gpu.items():
        v = parity.vesselness_report(g["J"], Jr)
        c = parity.code_report((g["Vx"], g["Vy"], g["Vz"]), tuple(Vr), Jr)
        rep["filter_" + name] = {"seconds": g["seconds"], "jmax": g["Jmax"], "jmin": g["Jmin"], "vesselness": v, "codes": c,
                                 "j8_equal_fraction": float((g["J8"] == J8r).mean())}

    # ---- downstream: seeds, correlation filter, SMC traces -- both arms in parallel processes
    jobs = [("reference", J8r, Vr[0], Vr[1], Vr[2], a.traces),
            ("gpu_exact", gpu["exact"]["J8"], gpu["exact"]["Vx"], gpu["exact"]["Vy"], gpu["exact"]["Vz"], a.traces),
            ("gpu_fma", gpu["fma"]["J8"], gpu["fma"]["Vx"], gpu["fma"]["Vy"], gpu["fma"]["Vz"], a.traces)]
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        out = dict(pool.map(_trace, jobs, chunksize=1))
    ref = out["reference"]
    for name in ("gpu_exact", "gpu_fma"):
        tr = out[name]
        srep = parity.seed_report(tr["seeds"][:, :6], ref["seeds"][:, :6])
        same_seeds = tr["seeds"].shape == ref["seeds"].shape and bool(np.array_equal(tr["seeds"], ref["seeds"]))
        key = lambda n: {tuple(np.round(r[:3], 3)) for r in n}
        na, nb = key(tr["nodes"]), key(ref["nodes"])
        rep["trace_" + name] = {
            "seconds": tr["seconds"], "seeds_extracted": tr["n_extracted"], "seeds_after_filter": len(tr["seeds"]),
            "traces": tr["n_traces"], "nodes": len(tr["nodes"]), "seed_report": srep, "seed_lists_identical": same_seeds,
            "node_lists_identical": bool(tr["nodes"].shape == ref["nodes"].shape and np.array_equal(tr["nodes"], ref["nodes"])
                                         and np.array_equal(tr["nbr"], ref["nbr"])),
            "node_position_match": len(na & nb) / max(1, len(na | nb)),
        }
    rep["trace_reference"] = {"seconds": ref["seconds"], "seeds_extracted": ref["n_extracted"],
                              "seeds_after_filter": len(ref["seeds"]), "traces": ref["n_traces"], "nodes": len(ref["nodes"])}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as fh:
        json.dump(rep, fh, indent=1)
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
