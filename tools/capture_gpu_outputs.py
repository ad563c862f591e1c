#!/usr/bin/env python
"""Captures the filter outputs of the GPU path for the volume of tests/test_plugin_e2e.py, so that they can be sent through
the reference's unmodified plugin on a host WITHOUT a GPU (oracle.Plugin("replay")).  Two steps:

    on the GPU box:       python tools/capture_gpu_outputs.py run   gpurun_out/gpu_192_outputs.npz
    in the build container (needs oracle/_ref, i.e. /root/reference):
                          python tools/capture_gpu_outputs.py pack  gpurun_out/gpu_192_outputs.npz tests/golden/gpu_capture_192.npz

`pack` stores the capture as a DIFFERENCE from the reference's own outputs on the same volume, which is also the finding:
the direction bytes are the reference's up to the sign of the eigenvector (one bit per voxel: which sign the GPU's closed
form chose), except a few dozen voxels that are off by one code; the 8-bit vesselness differs in a voxel or two.
tests/plugin_arms.load_capture() rebuilds the arrays.  TEST INFRASTRUCTURE: `pack` imports oracle/."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPE, SEED, SIGS = (192, 160, 48), 11, [2.0, 4.0, 6.0]       # w, h, l: the volume of tests/test_plugin_e2e.py


def run(out):
    import pnr_b200
    from pnr_b200.synth import make_volume
    I = make_volume(*SHAPE, seed=SEED)
    f = pnr_b200.Frangi(SIGS, 2.0, 0.5, 0.5, 500.0, flags=0)           # what the drop-in class runs: bit-exact smoothing
    g = f.frangi3d_full(I, want_J8=True)
    f.close()
    np.savez_compressed(out, J=g["J"], Jmin=g["Jmin"], Jmax=g["Jmax"], Vx=g["Vx"], Vy=g["Vy"], Vz=g["Vz"], J8=g["J8"])
    print("captured", out, float(g["Jmin"]), float(g["Jmax"]))


def plugin_j8(J, Jmin, Jmax):
    """Advantra_plugin.cpp:2499-2512, float arithmetic."""
    J = np.asarray(J, np.float32)
    v = (J - np.float32(Jmin)) / (np.float32(Jmax) - np.float32(Jmin)) * np.float32(255)
    return np.clip(np.floor(np.abs(v) + np.float32(0.5)) * np.sign(v), 0, 255).astype(np.uint8)


def pack(src, dst):
    from oracle import Reference
    from pnr_b200.synth import make_volume, volume_hash
    I = make_volume(*SHAPE, seed=SEED)
    G = np.load(src)
    r = Reference().frangi3d(I, SIGS)
    dec = lambda v: v.astype(np.float32) / 255 * 2 - 1
    dot = sum(dec(r[k]) * dec(G[k]) for k in ("Vx", "Vy", "Vz"))
    flip = dot < 0
    out = dict(input_hash=volume_hash(I), shape=np.array(SHAPE), flip=np.packbits(flip.ravel()))
    for k in ("Vx", "Vy", "Vz"):
        mine = np.where(flip, 255 - r[k], r[k]).astype(np.uint8)
        idx = np.flatnonzero(mine.ravel() != G[k].ravel())
        out[k + "_idx"], out[k + "_val"] = idx.astype(np.int64), G[k].ravel()[idx]
    j8_ref, j8_gpu = plugin_j8(r["J"], r["Jmin"], r["Jmax"]), plugin_j8(G["J"], G["Jmin"], G["Jmax"])
    assert np.array_equal(j8_gpu, G["J8"]), "the plugin's host conversion of the GPU's J and the device's J8 disagree"
    idx = np.flatnonzero(j8_ref.ravel() != j8_gpu.ravel())
    out["J8_idx"], out["J8_val"] = idx.astype(np.int64), j8_gpu.ravel()[idx]
    np.savez_compressed(dst, **out)
    print("packed", dst, "flipped", float(flip.mean()), "code differences beyond sign",
          {k: len(out[k + "_idx"]) for k in ("Vx", "Vy", "Vz")}, "J8 differences", len(idx), os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(sys.argv[2])
    else:
        pack(sys.argv[2], sys.argv[3])
