#!/usr/bin/env python
"""Cost of FRANGI_GPU_FLAG_REFERENCE_DIRECTION: wall time of the second frangi3d call (host buffers, copies included)
with and without the flag on a 512x512x256 volume; the difference is the double-precision pass (three scales)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import pnr_b200
from pnr_b200.synth import make_volume

I = np.ascontiguousarray(np.tile(make_volume(256, 256, 64, seed=7), (4, 2, 2)))
out = dict(shape=list(I.shape[::-1]), voxels=int(I.size))
for name, fl in (("plain", 0), ("reference_direction", pnr_b200.FLAG_REFERENCE_DIRECTION)):
    f = pnr_b200.Frangi([2.0, 4.0, 6.0], 2.0, 0.5, 0.5, 500.0, flags=fl)
    f.frangi3d_full(I)
    ts = []
    for _ in range(3):
        t = time.perf_counter(); f.frangi3d_full(I); ts.append(time.perf_counter() - t)
    f.close()
    out[name + "_ms"] = 1e3 * min(ts)
out["pass_ms"] = out["reference_direction_ms"] - out["plain_ms"]
out["pass_ns_per_voxel"] = 1e6 * out["pass_ms"] / out["voxels"]
print(json.dumps(out))
