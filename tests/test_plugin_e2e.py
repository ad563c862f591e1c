"""BASELINE.json configs[4] through the reference's WHOLE plugin translation unit: Advantra_plugin.cpp, unmodified, built
against stand-ins for Qt and the Vaa3D API (oracle/plugin_wrap.cpp, oracle/stubs/) and driven through its batch entry
point Advantra::dofunc("advantra_func") -- image load, Frangi (Advantra_plugin.cpp:2488-2512), seeds, correlation filter,
SMC traces, reconstruct() (:2096-2181: interpolate_nodelist, non_blurring, group1, bfs2, tree extraction) and the SWC
export (save_nodelist :480-523).  Two builds of that one file: with the reference's frangi.h / frangi.cpp, and with the
drop-in `class Frangi` of pnr_b200/csrc/frangi.h behind the same unchanged call site.  The comparison is on every file
the plugin writes (with its intermediate-result switch on, and the single-tree export on -- with the defaults
reconstruct() writes no final SWC at all)."""
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


@pytest.mark.gpu
def test_whole_plugin_gpu_frangi_vs_reference_frangi_swc():
    """GPU: the same unchanged plugin with the drop-in Frangi (flags = 0: bit-exact smoothing) against the all-reference
    build, 192x160x48 and 10 traces (the reference tracer needs ~1 s per trace).  J8 / direction bytes may differ in a
    handful of voxels (fp32 closed-form eigen stage vs the reference's QL iteration), so the bar is the one of
    test_trace_e2e: identical files where the seed lists are identical, else >= 99 % common node positions."""
    _need("ref")
    _need("gpu")
    I = make_volume(192, 160, 48, seed=11)
    a = run_arm("gpu", I, README_PARAMS, 10)
    b = run_arm("ref", I, README_PARAMS, 10)
    rep = compare_files(a["files"], b["files"])
    print("whole plugin, gpu vs ref:", {k: (v["identical"], v.get("position_match")) for k, v in rep.items()},
          "seconds", a["seconds"], b["seconds"])
    assert set(a["files"]) == set(b["files"]) and FINAL in a["files"]
    assert len(swc_rows(b["files"][FINAL])) > 50
    if rep["_Seeds.swc"]["identical"]:
        for k in ("_n0_.swc", "_n0res_.swc", "_n1_.swc", "_n2_.swc", "_n2tree_.swc", FINAL):
            assert rep[k]["identical"], (k, rep[k])
    else:
        assert rep["_Seeds.swc"]["position_match"] >= 0.999, rep["_Seeds.swc"]
        for k in ("_n0_.swc", "_n2_.swc", FINAL):
            assert rep[k]["position_match"] >= 0.99, (k, rep[k])
