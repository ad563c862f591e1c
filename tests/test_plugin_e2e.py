"""BASELINE.json configs[4] through the reference's WHOLE plugin translation unit: Advantra_plugin.cpp, unmodified, built
against stand-ins for Qt and the Vaa3D API (oracle/plugin_wrap.cpp, oracle/stubs/) and driven through its batch entry
point Advantra::dofunc("advantra_func") -- image load, Frangi (Advantra_plugin.cpp:2488-2512), seeds, correlation filter,
SMC traces, reconstruct() (:2096-2181: interpolate_nodelist, non_blurring, group1, bfs2, tree extraction) and the SWC
export (save_nodelist :480-523).  Three builds of that one file: with the reference's frangi.h / frangi.cpp; with the
drop-in `class Frangi` of pnr_b200/csrc/frangi.h behind the same unchanged call site; and with a frangi3d that hands back
arrays the test supplies (filter outputs captured on the GPU box, or the reference's own with some eigenvector signs
turned), which brings the comparison to a host without a GPU.  The comparison is on every file the plugin writes (with its
intermediate-result switch on, and the single-tree export on -- with the defaults reconstruct() writes no final SWC at
all).  What it finds: the drop-in differs from the reference through the SIGN of the eigenvector and
through nothing else; tests/test_reference_direction.py runs the drop-in with the reference's signs
(FRANGI_GPU_FLAG_REFERENCE_DIRECTION) and gets the reference's files, all of them."""
import os
import subprocess

import numpy as np
import pytest

from oracle import PLUGIN_GPU_SO, PLUGIN_REF_SO, Plugin
from pnr_b200.synth import make_volume
from tests.plugin_arms import compare_files, run_arm, swc_rows

README_PARAMS = ("2,4,6", "0", "5", "0.3", "3", "2", "200", "20", "2", "4", "1")
FINAL = "_Advantra1.swc"


def _need(arm):
    if not Plugin.available(arm):
        pytest.skip(f"oracle/_ref/libpnr_plugin_{arm}.so not built (needs /root/reference: `make -C oracle ref`)")


def _dyn_symbols(path, flag):
    out = subprocess.run(["nm", "-D", "-C", flag, path], capture_output=True, text=True, check=True).stdout
    return [ln.split(None, 2 if flag == "--defined-only" else 1)[-1] for ln in out.splitlines() if ln.strip()]


def test_whole_plugin_runs_qt_free_and_is_reproducible(reference, oracle):
    """CPU, all-reference arm: dofunc runs to the final SWC, twice with identical files; and its raw node list and its
    seed list are the ones ref_trace (oracle/ref_wrap.cpp, the restated call site the other tests use) produces."""
    _need("ref")
    w, h, l = 96, 80, 24
    I = make_volume(w, h, l, seed=3, n_neurites=4)
    params = list(README_PARAMS)
    params[6] = "60"
    a = run_arm("ref", I, params, 3)
    b = run_arm("ref", I, params, 3)
    assert FINAL in a["files"] and len(swc_rows(a["files"][FINAL])) > 10
    for k in ("_Seeds.swc", "_n0_.swc", "_n0res_.swc", "_n1_.swc", "_n2_.swc", "_n2tree_.swc"):
        assert k in a["files"], sorted(a["files"])
    rep = compare_files(a["files"], b["files"])
    assert all(r["identical"] for r in rep.values()), rep
    # the header of the final SWC carries the parameters the plugin parsed (Advantra_plugin.cpp:2283-2307)
    head = a["files"][FINAL]
    assert "#neuritesigmas=2,4,6" in head and "#ni=60" in head and "#MAX_TRACE_COUNT=3" in head
    # with the plugin's own defaults (ENFORCE_SINGLE_TREE = false, Advantra_plugin.cpp:81) reconstruct() writes NO final SWC:
    # the export block sits inside `if (ENFORCE_SINGLE_TREE)` (:2142-2166) -- the unmodified file shows it
    c = run_arm("ref", I, params, 3, env=dict(PNR_PLUGIN_SINGLE_TREE="0"))
    assert FINAL not in c["files"] and "_n2tree_.swc" in c["files"]
    assert c["files"]["_n0_.swc"] == a["files"]["_n0_.swc"]
    # against the restated call site
    if reference.has_trace:
        r = reference.frangi3d(I, [2.0, 4.0, 6.0])
        j8 = oracle.j_to_j8(r["J"], r["Jmin"], r["Jmax"])
        tr = reference.trace(I, j8, r["Vx"], r["Vy"], r["Vz"], [2.0, 4.0, 6.0], max_traces=3, ni=60)
        n0 = swc_rows(a["files"]["_n0_.swc"])
        pos_plugin = {tuple(np.round(x[2:5], 3)) for x in n0}
        pos_trace = {tuple(np.round(x[:3].astype(np.float64), 3)) for x in tr["nodes"][1:]}
        assert pos_plugin == pos_trace, (len(pos_plugin), len(pos_trace))
        seeds = swc_rows(a["files"]["_Seeds.swc"])
        assert len(seeds) == 2 * len(tr["seeds"])          # export_seeds writes the locus and a direction tip per seed


def test_gpu_arm_differs_from_the_reference_arm_only_in_class_frangi():
    """CPU: the drop-in build of the plugin defines no Frangi member itself and takes every one it names from
    libfrangi_shim.so; the all-reference build defines them.  Everything else in the two libraries is the same source."""
    _need("ref")
    _need("gpu")
    undef = [s for s in _dyn_symbols(PLUGIN_GPU_SO, "--undefined-only") if s.startswith("Frangi::")]
    own = [s for s in _dyn_symbols(PLUGIN_GPU_SO, "--defined-only") if "Frangi::" in s]
    ref_own = [s for s in _dyn_symbols(PLUGIN_REF_SO, "--defined-only") if s.startswith("Frangi::")]
    assert not own, own
    names = {s.split("(")[0] for s in undef}
    assert {"Frangi::Frangi", "Frangi::~Frangi", "Frangi::frangi3d"} <= names, names
    assert any(s.startswith("Frangi::frangi3d") for s in ref_own)
    shim = os.path.join(os.path.dirname(os.path.dirname(PLUGIN_GPU_SO)), "..", "pnr_b200", "_lib", "libfrangi_shim.so")
    exported = set(_dyn_symbols(os.path.normpath(shim), "--defined-only"))
    missing = [s for s in undef if s not in exported]
    assert not missing, missing


SIGS = [2.0, 4.0, 6.0]
CAPTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gpu_capture_192.npz")
NODE_FILES = ("_n0_.swc", "_n0res_.swc", "_n0tree_.swc", "_n1_.swc", "_n2_.swc", "_n2tree_.swc", FINAL)
DUMP = "_VxVyVz.swc"       # every 10th voxel's direction: differs wherever a sign does


def _sign_aware_bars(rep):
    """GPU-filter arm against the all-reference arm.  The direction bytes agree up to the SIGN of the eigenvector (which
    is arbitrary: the reference writes whatever its QL iteration returns, frangi.cpp:239-250, and the closed form of
    the GPU path returns the other one in half of the voxels).  The plugin is not indifferent to that sign -- the
    correlation score of a seed is summed in a different order (a tie in the seed ORDER can turn), and a trace is run
    along +v first and -v second, so the traces that share a neurite meet in a different order -- hence not identical
    files but: the same seed loci with directions within 1 degree modulo sign, the same raw nodes to >= 99 %, every node
    of every stage within two voxels of a node of the other arm, and the nodes of the final tree within half a voxel."""
    s = rep["_Seeds.swc"]
    assert s["identical"] or (s["loci_match"] >= 0.999 and s["direction_over_1deg"] <= 0.001 * s["loci"][0]), s
    for k in NODE_FILES:
        r = rep[k]
        assert r["identical"] or r["within_two_voxels"] >= 0.99, (k, r)
    assert rep["_n0_.swc"]["identical"] or rep["_n0_.swc"]["position_match"] >= 0.99, rep["_n0_.swc"]
    assert rep[FINAL]["identical"] or rep[FINAL]["within_half_voxel"] >= 0.99, rep[FINAL]


def test_gpu_filter_outputs_through_the_unmodified_plugin_differ_from_the_reference_only_by_eigenvector_sign(reference):
    """CPU.  tests/golden/gpu_capture_192.npz holds what the GPU path returned for this volume (captured on the B200 box by
    tools/capture_gpu_outputs.py) as a difference from the reference's outputs.  The replay build of the plugin (the
    unmodified translation unit, frangi3d handing back supplied arrays) then shows, without a GPU:
      1. replaying the reference's own outputs reproduces the all-reference arm file for file (the replay is faithful);
      2. replaying the GPU's outputs and replaying the reference's outputs WITH ONLY THE GPU's EIGENVECTOR SIGNS give the
         same files: the few direction codes that differ by one and the J8 voxel that differs change nothing downstream;
      3. against the all-reference arm the sign-aware bars hold."""
    _need("ref")
    _need("replay")
    from tests.plugin_arms import load_capture, plugin_j8
    from pnr_b200.synth import volume_hash
    I = make_volume(192, 160, 48, seed=11)
    r = reference.frangi3d(I, SIGS)
    cap = load_capture(CAPTURE, r)
    assert cap["input_hash"] == volume_hash(I)
    n = I.size
    assert 0.3 < cap["flip"].mean() < 0.7                                  # the sign is a coin toss between the two solvers
    assert all(v <= 1e-4 * n for v in cap["beyond_sign"].values()), cap["beyond_sign"]
    for k in ("Vx", "Vy", "Vz"):                                           # ... and beyond the sign: one code unit at most
        d = np.abs(cap[k].astype(int) - cap["signs_only"][("J8", "Vx", "Vy", "Vz").index(k)].astype(int))
        assert d.max() <= 1
    ref = run_arm("ref", I, README_PARAMS, 10)["files"]
    own = run_arm("replay", I, README_PARAMS, 10, replay=(plugin_j8(r["J"], r["Jmin"], r["Jmax"]), r["Vx"], r["Vy"], r["Vz"]))["files"]
    rep = compare_files(own, ref)
    assert all(v["identical"] for v in rep.values()), rep                  # 1.
    gpu = run_arm("replay", I, README_PARAMS, 10, replay=(cap["J8"], cap["Vx"], cap["Vy"], cap["Vz"]))["files"]
    sig = run_arm("replay", I, README_PARAMS, 10, replay=cap["signs_only"])["files"]
    rep = compare_files(gpu, sig)
    assert all(v["identical"] for k, v in rep.items() if k != DUMP), rep   # 2.
    assert rep[DUMP]["identical"] or rep[DUMP]["direction_over_1deg"] == 0, rep[DUMP]
    rep = compare_files(gpu, ref)
    print("GPU capture vs all-reference arm:", {k: {a: b for a, b in v.items() if a != "bytes"} for k, v in rep.items()})
    assert len(swc_rows(ref[FINAL])) > 50
    assert not rep["_Seeds.swc"]["identical"] and rep["_Seeds.swc"]["sign_flipped"] > 100      # the finding itself
    _sign_aware_bars(rep)                                                  # 3.


@pytest.mark.gpu
def test_whole_plugin_gpu_frangi_vs_reference_frangi_swc(reference):
    """GPU: the same unchanged plugin with the drop-in Frangi (flags = 0: bit-exact smoothing) against the all-reference
    build, 192x160x48 and 10 traces (the reference tracer needs ~0.5 s per trace).
      1. the filter outputs on this box are the committed capture, bit for bit (so the CPU test above speaks for this path);
      2. the drop-in arm writes exactly the files the replay of those outputs writes (the class boundary adds nothing);
      3. against the all-reference arm the sign-aware bars hold."""
    import pnr_b200
    from tests.plugin_arms import load_capture, plugin_j8
    _need("ref")
    _need("gpu")
    _need("replay")
    I = make_volume(192, 160, 48, seed=11)
    f = pnr_b200.Frangi(SIGS, 2.0, 0.5, 0.5, 500.0, flags=0)
    g = f.frangi3d_full(I, want_J8=True)
    f.close()
    j8 = plugin_j8(g["J"], g["Jmin"], g["Jmax"])
    assert np.array_equal(j8, g["J8"])                                     # the plugin's host conversion == the device's (K4)
    cap = load_capture(CAPTURE, reference.frangi3d(I, SIGS))
    for k in ("J8", "Vx", "Vy", "Vz"):
        assert np.array_equal(cap[k], g[k]), (k, int(np.sum(cap[k] != g[k])))          # 1.
    a = run_arm("gpu", I, README_PARAMS, 10)
    c = run_arm("replay", I, README_PARAMS, 10, replay=(j8, g["Vx"], g["Vy"], g["Vz"]))
    rep = compare_files(a["files"], c["files"])
    assert all(v["identical"] for v in rep.values()), rep                  # 2.
    b = run_arm("ref", I, README_PARAMS, 10)
    rep = compare_files(a["files"], b["files"])
    print("whole plugin, gpu vs ref:", {k: {x: y for x, y in v.items() if x != "bytes"} for k, v in rep.items()},
          "seconds", a["seconds"], b["seconds"])
    assert set(a["files"]) == set(b["files"]) and FINAL in a["files"]
    assert len(swc_rows(b["files"][FINAL])) > 50
    _sign_aware_bars(rep)                                                  # 3.
