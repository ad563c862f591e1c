"""A/B check of two builds of the library: run the pipeline on seeded volumes with each build (a subprocess per build,
FRANGI_GPU_LIB selects it) and report, per output array, whether the builds agree bit for bit and how many voxels differ.

    python tools/ab_outputs.py pnr_b200/_lib/libfrangi_gpu_base.so pnr_b200/_lib/libfrangi_gpu.so

Ad hoc development tool (bench.py and tests/ are the contract)."""
import os, subprocess, sys, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [  # w, h, l, sigmas, flags
    (300, 200, 40, [2., 4., 6.], 1),
    (300, 200, 40, [2., 4., 6.], 0),
    (2048, 72, 24, [2., 4.], 1),       # the width whose last tile column holds six voxel columns
    (250, 64, 9, [1., 2., 3.], 1),
    (257, 40, 12, [2., 3.], 1),        # (w - 4) mod 124 = 5: the last five interior columns go to the shell
    (128, 24, 10, [1., 2.], 1),        # a single, partly filled tile column
]


def worker(out):
    sys.path.insert(0, ROOT)
    import pnr_b200
    from pnr_b200.synth import make_volume
    res = {}
    for ci, (w, h, l, sigs, flags) in enumerate(CASES):
        I = make_volume(w, h, l)
        p = pnr_b200.FrangiPlan(sigs, 2.0, .5, .5, 500., False, w, h, l, flags=flags)
        p.upload(I)
        p.run_resident()
        d = p.download(want_J8=True)
        for k, v in d.items():
            if isinstance(v, np.ndarray):
                res[f"c{ci}_{k}"] = v
        p.close()
    np.savez(out, **res)


if __name__ == "__main__":
    if sys.argv[1] == "--worker":
        worker(sys.argv[2])
        sys.exit(0)
    libs = sys.argv[1:]
    outs = []
    tmp = tempfile.mkdtemp()
    for i, lib in enumerate(libs):
        o = os.path.join(tmp, f"o{i}.npz")
        env = dict(os.environ, FRANGI_GPU_LIB=os.path.abspath(lib))
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--worker", o], env=env)
        outs.append(np.load(o))
    a = outs[0]
    for i in range(1, len(outs)):
        b = outs[i]
        print(f"== {libs[0]} vs {libs[i]}")
        for k in a.files:
            x, y = a[k], b[k]
            if x.shape != y.shape:
                print(f"  {k}: shapes differ"); continue
            same = x.tobytes() == y.tobytes()
            nd = int((x != y).sum()) if not same else 0
            extra = ""
            if not same and x.dtype.kind == "f" and x.ndim > 0:
                den = np.maximum(np.abs(x.astype(np.float64)), 1e-30)
                extra = f" max rel {float((np.abs(x.astype(np.float64) - y) / den)[x != y].max()):.3g}"
            print(f"  {k}: {'identical' if same else f'{nd} of {x.size} differ' + extra}")
