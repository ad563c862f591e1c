"""Compares the chunk-pipelined frangi_gpu_run with the one-piece resident run at a given size (ad hoc)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pnr_b200
from pnr_b200.synth import make_volume

w, h, l = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "2048x2048x128").split("x"))
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 1
base = make_volume(512, 512, 128)
I = np.tile(base, (max(1, l // 128), max(1, h // 512), max(1, w // 512)))[:l, :h, :w].copy()
p = pnr_b200.FrangiPlan([2., 4., 6.], 2.0, .5, .5, 500., False, w, h, l, flags=flags)
p.upload(I)
lo, hi = p.run_resident()
one = p.download(want_J8=True)
print("resident Jmin/Jmax", lo, hi, flush=True)
for rep in range(2):
    many = p.run(I, want_J8=True)
    print("streamed Jmin/Jmax", many["Jmin"], many["Jmax"], flush=True)
    for k in ("J", "Vx", "Vy", "Vz", "J8"):
        d = one[k] != many[k]
        n = int(d.sum())
        print(k, "mismatches", n)
        if n and k == "J":
            zz, yy, xx = np.nonzero(d)
            print("  z:", np.unique(zz)[:40], " count per z mod 32:", np.bincount(zz % 32, minlength=32))
            print("  y mod 8:", np.bincount(yy % 8, minlength=8), " x mod 120:", np.bincount(xx % 120, minlength=120)[:12], "...")
            for i in range(min(8, n)):
                print("  ", zz[i], yy[i], xx[i], one["J"][zz[i], yy[i], xx[i]], many["J"][zz[i], yy[i], xx[i]])
p.close()
