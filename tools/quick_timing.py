"""Quick per-kernel-class device timings (ad hoc; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pnr_b200
from pnr_b200.synth import make_volume

def run(w, h, l, sigs, flags=0, reps=3, tile=None):
    t = time.time()
    if tile is None:
        I = make_volume(w, h, l)
    else:
        base = make_volume(*tile)
        I = np.tile(base, (l // tile[2], h // tile[1], w // tile[0]))
    tg = time.time() - t
    p = pnr_b200.FrangiPlan(sigs, 2.0, .5, .5, 500., False, w, h, l, flags=flags)
    p.upload(I)
    for _ in range(2): p.run_resident()
    best = None
    for _ in range(reps):
        p.run_resident()
        tm = p.timings()
        if best is None or tm["total"] < best["total"]: best = tm
    vox = w * h * l
    print(f"{w}x{h}x{l} sig={sigs} flags={flags} gen={tg:.1f}s  " + " ".join(f"{k}={v:.3f}ms" for k, v in best.items()) +
          f"  -> {vox / best['total'] / 1e6:.1f} Gvox/s", flush=True)
    p.close()

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    print("lib:", os.environ.get("FRANGI_GPU_LIB", "default"))
    if which == "all":
        run(256, 256, 64, [2., 4., 6.])
        run(512, 512, 128, [2., 4., 6.], flags=1)
        run(1024, 1024, 256, [1., 2., 3., 4., 5., 6.], tile=(256, 256, 64))
    run(2048, 2048, 512, [2., 4., 6.], flags=1, tile=(512, 512, 128))
