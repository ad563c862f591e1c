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


def run_arm(arm, I, params, max_traces, workdir=None, timeout=3600):
    """-> dict(files={suffix: text}, seconds=wall time of dofunc, log=the plugin's stdout)."""
    workdir = workdir or tempfile.mkdtemp(prefix=f"pnr_plugin_{arm}_")
    vol = os.path.join(workdir, "_vol.npy")
    np.save(vol, np.ascontiguousarray(I, np.uint8))
    cmd = [sys.executable, "-m", "tests.plugin_arms", arm, vol, workdir, str(int(max_traces))] + [str(p) for p in params]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, errors="replace", timeout=timeout)
    os.remove(vol)
    if out.returncode != 0:
        raise RuntimeError(f"plugin arm {arm!r} failed (rc {out.returncode}): {out.stderr[-2000:]}")
    with open(os.path.join(workdir, "_result.json")) as f:
        res = json.load(f)
    os.remove(os.path.join(workdir, "_result.json"))
    res["log"] = out.stdout
    return res


def swc_rows(text):
    """The numeric rows of an SWC file: [n, 7] float64 (n type x y z r parent)."""
    rows = [ln.split() for ln in text.splitlines() if ln and not ln.startswith("#")]
    return np.array(rows, np.float64).reshape(-1, 7) if rows else np.zeros((0, 7))


def compare_files(a, b):
    """Per file: identical text, or (for SWC files) the fraction of node positions present in both."""
    rep = {}
    for k in sorted(set(a) | set(b)):
        if k not in a or k not in b:
            rep[k] = dict(identical=False, missing_in="ref" if k not in b else "gpu")
            continue
        r = dict(identical=a[k] == b[k], bytes=len(a[k]))
        if k.endswith(".swc"):
            ra, rb = swc_rows(a[k]), swc_rows(b[k])
            sa = {tuple(np.round(x[2:5], 3)) for x in ra}
            sb = {tuple(np.round(x[2:5], 3)) for x in rb}
            r.update(nodes=[len(ra), len(rb)], position_match=len(sa & sb) / max(1, len(sa | sb)))
        rep[k] = r
    return rep


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    from oracle import Plugin
    arm, vol, workdir, max_traces = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
    I = np.load(vol)
    p = Plugin(arm)
    t0 = time.time()
    files = p.run(I, workdir, params=sys.argv[5:], max_traces=max_traces)
    dt = time.time() - t0
    with open(os.path.join(workdir, "_result.json"), "w") as f:
        json.dump(dict(files=files, seconds=dt), f)
