"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck): every kernel variant once,
on ragged shapes, with multi-slab and chunk-pipelined paths."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pnr_b200
from pnr_b200 import FLAG_DIR_F32, FLAG_FMA_SMOOTHING, FLAG_SCALE_IDX, FrangiPlan
from pnr_b200.synth import make_volume

def run(w, h, l, sigs, flags, devices=(0,), chunk=None, bw=False):
    I = make_volume(w, h, l, seed=3, n_neurites=4)
    p = FrangiPlan(sigs, 2.0, .5, .5, 500., bw, w, h, l, devices=devices, flags=flags)
    if chunk is not None:
        p.set_stream_chunk(chunk)
    r = p.run(I, want_J8=True)
    p.close()
    print(w, h, l, sigs, flags, devices, chunk, "Jmax", r["Jmax"], flush=True)

run(150, 70, 48, [2., 4., 6.], FLAG_DIR_F32 | FLAG_SCALE_IDX)
run(150, 70, 48, [2., 4., 6.], FLAG_FMA_SMOOTHING)
run(131, 37, 29, [1., 3.], FLAG_FMA_SMOOTHING, bw=True)
run(150, 70, 48, [2., 4., 6.], FLAG_FMA_SMOOTHING, devices=(0, 0, 0))
run(150, 70, 48, [2., 4., 6.], 0, chunk=11)
run(260, 40, 12, [5.], 0)
run(4, 4, 4, [2.], 0)
F = pnr_b200.Frangi.imgaussian(make_volume(37, 29, 11, seed=1, n_neurites=2), 4.0, 1.0)
D = pnr_b200.Frangi([2.], 2., .5, .5, 500.).hessian3d(make_volume(37, 29, 11, seed=1, n_neurites=2), 2.0, 2.0)
print("ok")
