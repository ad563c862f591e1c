#!/usr/bin/env python
"""BASELINE.json configs[4] through the reference's WHOLE, UNMODIFIED plugin translation unit at a stated size: one arm
per invocation (the all-reference arm needs no GPU and is run in the build container; the drop-in arm runs on the GPU
box), then a comparison of what the two wrote.  See tests/test_plugin_e2e.py for what the arms are.

    python tools/plugin_e2e.py --arm ref --size 512x512x128 --traces 40 --out tests/golden/plugin_512_ref.json.gz
    python tools/plugin_e2e.py --arm gpu --size 512x512x128 --traces 40 --out gpurun_out/plugin_512_gpu.json.gz \\
           --compare tests/golden/plugin_512_ref.json.gz --report gpurun_out/plugin_512_report.json
    python tools/plugin_e2e.py --saved gpurun_out/plugin_512_gpu.json.gz --compare tests/golden/plugin_512_ref.json.gz \\
           --report profiles/r5_plugin_512_report.json                      (compare again, run nothing)

The saved file holds the text of every SWC / log file the plugin wrote, except the direction dump (_VxVyVz.swc, one row
per 10th voxel: its SHA-256 and row count only).  TEST INFRASTRUCTURE: imports oracle/."""
import argparse
import gzip
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BIG = ("_VxVyVz.swc",)


def pack(files):
    out = {}
    for k, v in files.items():
        if k in BIG:
            out[k] = dict(sha256=hashlib.sha256(v.encode()).hexdigest(), rows=v.count("\n"))
        else:
            out[k] = v
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", choices=["ref", "gpu"], default=None)
    ap.add_argument("--saved", default=None, help="with --compare and --report: compare two saved files, run nothing")
    ap.add_argument("--size", default="512x512x128")
    ap.add_argument("--traces", type=int, default=40)
    ap.add_argument("--seed", type=int, default=20181009 + 4)
    ap.add_argument("--params", default="2,4,6 0 5 0.3 3 2 200 20 2 4 1", help="the plugin's eleven parameters (README usage)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--compare", default=None, help="a file saved by the other arm")
    ap.add_argument("--report", default=None)
    a = ap.parse_args()
    from pnr_b200.synth import make_volume, volume_hash
    from tests.plugin_arms import compare_files, run_arm
    if a.saved:
        with gzip.open(a.saved, "rt") as f:
            rec = json.load(f)
        w, h, l = rec["size"]
    else:
        w, h, l = (int(v) for v in a.size.split("x"))
        t0 = time.time()
        I = make_volume(w, h, l, seed=a.seed)
        res = run_arm(a.arm, I, a.params.split(), a.traces, timeout=6 * 3600)
        rec = dict(arm=a.arm, size=[w, h, l], traces=a.traces, params=a.params, input_hash=volume_hash(I),
                   plugin_seconds=res["seconds"], total_seconds=time.time() - t0, files=pack(res["files"]))
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with gzip.open(a.out, "wt") as f:
            json.dump(rec, f)
        print(f"{a.arm}: {w}x{h}x{l}, {a.traces} traces, dofunc {res['seconds']:.1f} s, {len(res['files'])} files -> {a.out}")
    if a.compare:
        with gzip.open(a.compare, "rt") as f:
            other = json.load(f)
        assert other["input_hash"] == rec["input_hash"] and other["params"] == rec["params"] and other["traces"] == rec["traces"]
        mine, theirs = rec["files"], other["files"]
        rep = compare_files(mine, theirs)
        out = dict(size=[w, h, l], traces=rec["traces"], params=rec["params"], input_hash=rec["input_hash"],
                   seconds={rec["arm"]: rec["plugin_seconds"], other["arm"]: other["plugin_seconds"]},
                   all_identical=all(r.get("identical") for r in rep.values()),
                   identical_files=sorted(k for k, r in rep.items() if r.get("identical")), files=rep)
        print(json.dumps(out, indent=1))
        if a.report:
            with open(a.report, "w") as f:
                json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
