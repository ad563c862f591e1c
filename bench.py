#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the Frangi hot path.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU frangi3d

One "step" = one pass of the multi-scale 3-D Frangi filter (frangi.cpp:152-289)
over the workload volume.  Default workload: BASELINE.json configs[3], the volume
the metric's target is quoted on -- sigma = 2,4,6 on 2048x2048x512 uint8 (2^31
voxels, 34 GB of device buffers on one B200) -- z-slab sharded over the N ranks
(one process per GPU, halo exchange of the xy-smoothed planes over NCCL
send/recv, Jmin/Jmax all-reduce), so scaling is "strong".

Our arm prints ONE JSON line with
  value      whole-job voxel/s, input resident in HBM, device-timed (CUDA events on
             the library's stream), max over ranks
  e2e        the same metric through the reference-facing C-ABI call
             frangi_gpu_run (host buffers in, host buffers out, copies inside)
  roofline   the dominant kernel's algorithmic bytes / its mean launch time
             (events around every launch of the timed region) vs the measured
             HBM peak, plus whole-pipeline HBM and FP32 fractions
  cpu_baseline  the unmodified reference (oracle/_ref) on the host cores, on a
             bounded sample (N = 1 only)

The reference arm times oracle/_ref/libpnr_ref.so (the reference's own frangi.cpp,
compiled unmodified; falls back to the C port in oracle/ when that file did not
travel) on all host cores, each step a bounded sample: one block of FULL-WIDTH rows
of the workload per core (w x 64 x 32 voxels: the x stride of the real planes, which
is what the reference's stride-w y pass sees; a z stride of 512 KB, which like the
real 16 MB maps every tap of the stride-w*h z pass to one cache set), every core
running the reference's single-threaded frangi3d on its own block concurrently.

Nothing here reads /root/reference.  oracle/ is executed only in the
cpu_baseline leg and in the reference arm, never on the measured product path.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "frangi_3scale_voxels_per_s"
UNIT = "voxel/s"
ZDIST, ALPHA, BETA, CC = 2.0, 0.5, 0.5, 500.0
BASE_BLOCK = (512, 512, 128)      # seeded synthetic block (w, h, l) that is tiled to the workload
CPU_BLOCK_HL = (64, 32)           # per-core sample of the CPU arms: full-width rows of the workload, 64 rows, 32 planes


def cpu_block(w):
    return (w, CPU_BLOCK_HL[0], CPU_BLOCK_HL[1])


def config_name(w, h, l, sigmas):
    """Which BASELINE.json config a workload is (the label printed in config.workload)."""
    sig = [float(x) for x in sigmas]
    table = {((256, 256, 64), (2.0, 4.0, 6.0)): "configs[0]", ((512, 512, 128), (2.0, 4.0, 6.0)): "configs[1]",
             ((1024, 1024, 256), (1.0, 2.0, 3.0, 4.0, 5.0, 6.0)): "configs[2]",
             ((2048, 2048, 512), (2.0, 4.0, 6.0)): "configs[3], the volume the metric's target is quoted on"}
    return table.get(((w, h, l), tuple(sig)), "not a BASELINE.json config")

# algorithmic work per voxel
def pipeline_bytes_per_voxel(S):          # 16 + 17 (S - 1)
    return 16 + 17 * (S - 1)


def pipeline_flops_per_voxel(sigmas, zdist):
    import math
    f = 0
    for s in sigmas:
        lxy = math.ceil(3 * s)
        lz = math.ceil(3 * (s / zdist))
        f += 2 * (2 * (2 * lxy + 1) + (2 * lz + 1)) + 24 + 120
    return f


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="2048x2048x512", help="WxHxL")
    ap.add_argument("--sigmas", default="2,4,6")
    ap.add_argument("--exact", action="store_true",
                    help="bit-exact smoothing (separate rounded multiply and add) instead of FMA")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    ap.add_argument("--pageable", action="store_true",
                    help="e2e with ordinary (pageable) host buffers, as an unmodified caller would pass them")
    ap.add_argument("--no-verify", action="store_true",
                    help="N > 1: skip the check of every rank's slab of J / V, bit for bit, against an independent one-GPU "
                         "run of the same planes (on by default)")
    ap.add_argument("--no-exact", action="store_true", help="skip the extra timing of the bit-exact smoothing mode")
    ap.add_argument("--outputs", default="", help="extra outputs kept by the handle: comma list of dir, scale")
    return ap.parse_args()


# ----------------------------------------------------------------------------- inputs
def workload_slab(w, h, l, z0, z1, out=None):
    """Planes [z0, z1) of the synthetic workload: the seeded 512x512x128 neuron block
    (pnr_b200.synth.make_volume) tiled periodically along x, y and z."""
    from pnr_b200.synth import make_volume
    bw, bh, bl = min(BASE_BLOCK[0], w), min(BASE_BLOCK[1], h), min(BASE_BLOCK[2], l)
    base = make_volume(bw, bh, bl)
    reps_y, reps_x = -(-h // bh), -(-w // bw)
    if out is None:
        out = np.empty((z1 - z0, h, w), np.uint8)
    for z in range(z0, z1):
        plane = np.tile(base[z % bl], (reps_y, reps_x))[:h, :w]
        out[z - z0] = plane
    return out


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [ln for (t, ln) in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [ln for (_, ln) in self.lines]
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arms
_WORKER_CACHE = {}


def _cpu_worker(args):
    """One core: the reference's single-threaded frangi3d on its own block of full-width rows."""
    kind, block_index, sigmas, wblk = args
    from oracle import Oracle, Reference
    from pnr_b200.synth import make_volume
    w, h, l = cpu_block(wblk)
    if "I" not in _WORKER_CACHE:          # each worker process keeps its block and its library
        base = make_volume(min(w, BASE_BLOCK[0]), h, l, seed=20181009 + os.getpid() % 1000)
        _WORKER_CACHE["I"] = np.ascontiguousarray(np.tile(base, (1, 1, -(-w // base.shape[2])))[:, :, :w])
        _WORKER_CACHE["impl"] = Reference() if kind == "reference" else Oracle()
    I, impl = _WORKER_CACHE["I"], _WORKER_CACHE["impl"]
    t0 = time.perf_counter()
    if kind == "reference":
        r = impl.frangi3d(I, sigmas, ZDIST, ALPHA, BETA, CC)
    else:
        r = impl.frangi3d(I, sigmas, ZDIST, ALPHA, BETA, CC, want_scale=False, want_dir=False)
    return time.perf_counter() - t0, float(r["Jmax"])


def cpu_kind():
    from oracle import Reference
    return "reference" if Reference.available() else "port"


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class CpuArm:
    """P worker processes, one per host core; one step = every worker runs one block."""

    def __init__(self, sigmas, w):
        import multiprocessing as mp
        from oracle import Oracle, Reference
        self.kind = cpu_kind()
        # the library is loaded in the parent BEFORE the workers fork, so that it shows in this process's map
        # (the driver's native_so_loaded) and every worker inherits it
        self.impl = Reference() if self.kind == "reference" else Oracle()
        self.cores = cpu_cores()
        self.sigmas = list(sigmas)
        self.w = w
        self.block = cpu_block(w)
        self.pool = mp.get_context("fork").Pool(self.cores)
        self.voxels_per_step = self.cores * self.block[0] * self.block[1] * self.block[2]

    def step(self):
        t0 = time.perf_counter()
        self.pool.map(_cpu_worker, [(self.kind, i, self.sigmas, self.w) for i in range(self.cores)], chunksize=1)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()

    def sample_text(self):
        b = self.block
        return (f"{self.cores} concurrent single-threaded frangi3d calls, one per host core, each on its own "
                f"{b[0]}x{b[1]}x{b[2]} block of full-width rows of the workload per step (the real x stride; "
                f"wall clock around the whole step)")


def run_reference_arm(a, sigmas, w, h, l):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    arm = CpuArm(sigmas, w)
    for _ in range(a.warmup):
        arm.step()
    t = [arm.step() for _ in range(a.steps)]
    arm.close()
    total = sum(t)
    value = arm.voxels_per_step * a.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"frangi3d sigma={a.sigmas} zdist=2 alpha=beta=0.5 C=500 on {w}x{h}x{l} uint8 "
                               f"(BASELINE.json {config_name(w, h, l, sigmas)})",
                   "sample": arm.sample_text()},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                         "sample": arm.sample_text()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------- our arm
def run_ours(a, sigmas, w, h, l):
    import torch
    import torch.distributed as dist

    import pnr_b200
    from pnr_b200.frangi import FrangiPlan, PinnedBuffer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lib = pnr_b200.load_library()
    flags = 0 if a.exact else pnr_b200.FLAG_FMA_SMOOTHING
    extra = [x for x in a.outputs.split(",") if x]
    if "dir" in extra:
        flags |= pnr_b200.FLAG_DIR_F32
    if "scale" in extra:
        flags |= pnr_b200.FLAG_SCALE_IDX
    z0, z1 = l * rank // world, l * (rank + 1) // world
    nz = z1 - z0
    own_vox = w * h * nz
    total_vox = w * h * l

    # NCCL id of the library's own communicator (halo send/recv), from rank 0
    uid = None
    if world > 1:
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            import ctypes as C
            buf = C.create_string_buffer(128)
            rc = lib.frangi_gpu_nccl_unique_id(buf)
            if rc:
                raise SystemExit("nccl unique id: " + lib.frangi_gpu_last_error().decode())
            t.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
        dist.broadcast(t, 0)
        uid = bytes(t.cpu().numpy().tobytes())

    # host buffers (pinned): the slab's input and the reference interface's outputs
    hI = PinnedBuffer((nz, h, w), np.uint8)
    workload_slab(w, h, l, z0, z1, out=hI.array)
    plan = FrangiPlan(sigmas, ZDIST, ALPHA, BETA, CC, False, w, h, l, flags=flags,
                      slab=(z0, z1, rank, world, uid, local_rank))
    plan.upload(hI.array)
    plan.sync()
    stream = torch.cuda.ExternalStream(plan.stream(0), device=dev)

    # ---- device-resident timing -------------------------------------------------
    for _ in range(max(a.warmup, 3)):
        plan.run_resident(sync=False)
    plan.sync()
    plan.timing_depth(a.steps)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    barrier()
    launches0 = pnr_b200.launch_count()
    t_wall0 = time.time()
    ev0.record(stream)
    for _ in range(a.steps):
        plan.run_resident(sync=False)
    ev1.record(stream)
    plan.sync()
    barrier()
    t_wall1 = time.time()
    launches = pnr_b200.launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_total = ev0.elapsed_time(ev1)
    tm = plan.timings()                      # mean per-class device ms over the K timed runs (this rank)
    jmin, jmax = plan.run_resident()         # one more run to read the scalars (outside the timed region)
    tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    ll = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ll, op=dist.ReduceOp.SUM)
    ms_total = float(tt.item())
    launches = int(ll.item())
    ms_per_step = ms_total / a.steps
    value = total_vox / (ms_per_step * 1e-3)

    # ---- N-slab result == independent one-GPU result, bit for bit (every run with N > 1) -------------------------
    # Every rank recomputes ITS planes with a second, single-slab handle over the sub-volume [z0 - H, z1 + H) clipped
    # to the volume, H = ceil(3 sigma_max / zdist) + 2: every stage is z-local with that radius and the clamp /
    # one-sided rules of the sub-volume's ends only touch planes within H of them, so the planes [z0, z1) of the
    # sub-volume run are exactly what a whole-volume run produces there.  No exchange, no NCCL in the
    # checker: it shares nothing with the slab path but the kernels.  Jmin / Jmax (global scalars) are compared
    # through the all-gathered maxima.
    verify = None
    if world > 1 and not a.no_verify:
        import math
        import zlib
        H = max(math.ceil(3 * (sg / ZDIST)) for sg in sigmas) + 2
        za, zb = max(z0 - H, 0), min(z1 + H, l)
        out = plan.download(want_J8=False)
        sub = FrangiPlan(sigmas, ZDIST, ALPHA, BETA, CC, False, w, h, zb - za, flags=flags, devices=(local_rank,))
        sub.upload(workload_slab(w, h, l, za, zb))
        sub_lo, sub_hi = sub.run_resident()
        ref = sub.download(want_J8=False)
        sub.close()
        keys = [k for k in ("J", "Vx", "Vy", "Vz", "scale", "dir") if out.get(k) is not None]
        bad_keys = []
        for k in keys:
            got = out[k]
            want = ref[k][:, z0 - za:z1 - za] if k == "dir" else ref[k][z0 - za:z1 - za]
            if not np.array_equal(got, want):
                bad_keys.append(k)
        own_max = float(ref["J"][z0 - za:z1 - za].max())
        flag = torch.tensor([1 if bad_keys else 0, zlib.crc32(out["J"].tobytes())], dtype=torch.int64, device=dev)
        mx = torch.tensor([own_max], dtype=torch.float64, device=dev)
        allf = [torch.zeros_like(flag) for _ in range(world)]
        dist.all_gather(allf, flag)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        verify = {"method": "every rank's planes recomputed by an independent single-slab handle over its sub-volume "
                            f"[z0-{H}, z1+{H}); J, V (and scale / dir when kept) compared bit for bit",
                  "slabs": world, "mismatching_ranks": [r_ for r_, t_ in enumerate(allf) if int(t_[0].item())],
                  "slab_crc32_J": [int(t_[1].item()) for t_ in allf],
                  "jmax_slabs": jmax, "jmax_independent": float(mx.item())}
        if float(mx.item()) != jmax and 0 not in verify["mismatching_ranks"]:
            verify["mismatching_ranks"].append(-1)         # the global maximum disagrees
        del out, ref

    # ---- the bit-exact smoothing mode (separately rounded multiply and add), timed beside the FMA mode -----------
    exact_ms = None
    if world == 1 and not a.exact and not a.no_exact:
        pe = FrangiPlan(sigmas, ZDIST, ALPHA, BETA, CC, False, w, h, l, flags=flags & ~pnr_b200.FLAG_FMA_SMOOTHING,
                        devices=(local_rank,))
        pe.upload(hI.array)
        for _ in range(3):
            pe.run_resident(sync=False)
        pe.sync()
        se = torch.cuda.ExternalStream(pe.stream(0), device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(se)
        for _ in range(a.steps):
            pe.run_resident(sync=False)
        e1.record(se)
        pe.sync()
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        exact_ms = float(te.item()) / a.steps
        pe.close()

    # ---- end to end through the C-ABI call with host buffers ----------------------
    e2e = None
    if not a.no_e2e:
        if a.pageable:
            class _Plain:
                def __init__(self, shape, dt):
                    self.array = np.empty(shape, dt)
                    self.array.fill(0)          # touch the pages
                def free(self):
                    self.array = None
            hJ = _Plain((nz, h, w), np.float32)
            hV = [_Plain((nz, h, w), np.uint8) for _ in range(3)]
            hIe = np.array(hI.array)
        else:
            hJ = PinnedBuffer((nz, h, w), np.float32)
            hV = [PinnedBuffer((nz, h, w), np.uint8) for _ in range(3)]
            hIe = hI.array
        k_e2e = a.e2e_steps or min(a.steps, 5)
        # --outputs: the extra arrays the handle keeps come back too, into buffers of the same kind (without these the
        # Python mirror would allocate fresh pageable arrays inside every timed call)
        Buf = _Plain if a.pageable else PinnedBuffer
        hS = Buf((nz, h, w), np.uint8) if (flags & pnr_b200.FLAG_SCALE_IDX) else None
        hD = Buf((3, nz, h, w), np.float32) if (flags & pnr_b200.FLAG_DIR_F32) else None
        extra_out = dict(scale=hS.array if hS else None, direction=hD.array if hD else None)
        extra_bytes = (1 if hS else 0) + (12 if hD else 0)
        plan.run(hIe, J=hJ.array, Vx=hV[0].array, Vy=hV[1].array, Vz=hV[2].array, **extra_out)   # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            r = plan.run(hIe, J=hJ.array, Vx=hV[0].array, Vy=hV[1].array, Vz=hV[2].array, **extra_out)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te.item())
        e2e = {"value": total_vox * k_e2e / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(total_vox), "d2h_bytes_per_step": int(total_vox * (7 + extra_bytes) + 8 * world),
               "steps": k_e2e, "ms_per_step": 1e3 * dt / k_e2e,
               "call": "frangi_gpu_run(I_host -> J_host f32, Jmin, Jmax, Vx, Vy, Vz host u8), "
                       + ("pageable" if a.pageable else "pinned") + " host buffers",
               "jmax": float(r["Jmax"])}
        # the same call as the caller really needs it (J is freed at once, Advantra_plugin.cpp:2514): J8 + V only
        if world == 1 and not a.pageable:
            hJ8 = PinnedBuffer((nz, h, w), np.uint8)
            plan.run(hI.array, J=None, Vx=hV[0].array, Vy=hV[1].array, Vz=hV[2].array, J8=hJ8.array, want_J=False, **extra_out)
            barrier()
            t0 = time.perf_counter()
            for _ in range(k_e2e):
                plan.run(hI.array, J=None, Vx=hV[0].array, Vy=hV[1].array, Vz=hV[2].array, J8=hJ8.array, want_J=False, **extra_out)
            torch.cuda.synchronize()
            dt8 = time.perf_counter() - t0
            e2e["j8_variant"] = {"value": total_vox * k_e2e / dt8, "unit": UNIT, "ms_per_step": 1e3 * dt8 / k_e2e,
                                 "d2h_bytes_per_step": int(total_vox * (4 + extra_bytes) + 8),
                                 "call": "frangi_gpu_run(I_host -> J8, Vx, Vy, Vz host u8; J_host = NULL)"}
            hJ8.free()
        # the buffers as the unmodified call site passes them (new float[size], Advantra_plugin.cpp:2490-2494): pageable
        if world == 1 and not a.pageable:
            pJ = np.zeros((nz, h, w), np.float32)
            pV = [np.zeros((nz, h, w), np.uint8) for _ in range(3)]
            pI = np.array(hI.array)
            p_extra = dict(scale=np.zeros((nz, h, w), np.uint8) if hS else None,
                           direction=np.zeros((3, nz, h, w), np.float32) if hD else None)
            plan.run(pI, J=pJ, Vx=pV[0], Vy=pV[1], Vz=pV[2], **p_extra)
            t0 = time.perf_counter()
            kp = min(k_e2e, 2)
            for _ in range(kp):
                plan.run(pI, J=pJ, Vx=pV[0], Vy=pV[1], Vz=pV[2], **p_extra)
            torch.cuda.synchronize()
            dtp = time.perf_counter() - t0
            e2e["pageable"] = {"value": total_vox * kp / dtp, "unit": UNIT, "ms_per_step": 1e3 * dtp / kp, "steps": kp,
                               "call": "the same call with ordinary (pageable) host buffers"}
            del pJ, pV, pI, p_extra
        for b in [hJ] + hV + [x for x in (hS, hD) if x is not None]:
            b.free()

    # ---- the extractSeeds pre-pass on the J8 volume the last run left on the device ----------
    seed_prepass = None
    if world == 1 and not a.no_e2e:
        c = plan.seed_candidates()                                   # warm-up, and the capacity for the timed call
        t0 = time.perf_counter()
        c = plan.seed_candidates(cap=max(1, len(c["keys"])))
        dt_s = time.perf_counter() - t0
        seed_prepass = {"ms": 1e3 * dt_s, "voxel_per_s": total_vox / dt_s, "candidates": int(len(c["keys"])),
                        "call": "frangi_gpu_seed_candidates (layer range + candidate maxima + ranked keys, keys to the host)"}
        if not a.no_cpu_baseline:
            from oracle import Oracle
            blk = min(l, 32)
            j8 = plan.download(want_J8=True, want_J=False, want_V=False)["J8"][:blk].copy()
            t0 = time.perf_counter()
            oc = Oracle().seed_candidates(j8)
            dt_c = time.perf_counter() - t0
            n_blk = int(c["n_max"][:blk].sum())
            seed_prepass["cpu_port"] = {"voxel_per_s": j8.size / dt_c, "cores": 1, "sample": f"first {blk} layers",
                                        "identical": bool(np.array_equal(oc["keys"], c["keys"][:n_blk]))}
            del j8

    if rank != 0:
        plan.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel -------------------------------------------
    S = len(sigmas)
    hbm_peak, peak_src = load_peaks()
    kernels = {
        # algorithmic bytes per OWN voxel, summed over the S launches of a step
        # (DESIGN.md section 4): xy: u8 in + f32 out; z: f32 in + f32 out;
        # voxel kernel: F in + J out + V out on scale 0, F in + J in + J out after
        "gauss_xy": 5 * S, "gauss_z": 8 * S, "hessian_eigen": 11 + 12 * (S - 1),
    }
    dom = max(kernels, key=lambda k: tm[k])
    dom_ms_per_launch = tm[dom] / S
    dom_bytes_per_launch = kernels[dom] / S * own_vox
    achieved = dom_bytes_per_launch / (dom_ms_per_launch * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get(f"{dom}:{w}x{h}x{l}")
    pipe_bytes = pipeline_bytes_per_voxel(S) * total_vox
    pipe_flops = pipeline_flops_per_voxel(sigmas, ZDIST) * total_vox
    fp32_peak = 148 * 128 * 2 * 1.965e9
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
        "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
        "kernel_ms_per_launch": dom_ms_per_launch, "kernel_bytes_per_launch": dom_bytes_per_launch,
        "kernel_share_of_step": tm[dom] / tm["total"],
        "per_class_ms_per_step": {k: tm[k] for k in ("gauss_xy", "gauss_z", "hessian_eigen", "j8", "halo_wait", "total")},
        "pipeline_hbm_frac": pipe_bytes / (ms_per_step * 1e-3) / 1e9 / (hbm_peak * world),
        "pipeline_bytes_per_voxel": pipeline_bytes_per_voxel(S),
        "pipeline_fp32_frac": pipe_flops / (ms_per_step * 1e-3) / (fp32_peak * world),
        "pipeline_flops_per_voxel": pipeline_flops_per_voxel(sigmas, ZDIST),
        "fp32_peak_tflops_nominal": fp32_peak / 1e12,
    }

    # ---- CPU baseline (bounded sample, rank 0 at N = 1 only) -------------------------
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        try:
            arm = CpuArm(sigmas, w)
            t = [arm.step() for _ in range(2)]
            arm.close()
            cpu = {"value": arm.voxels_per_step / min(t), "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                   "sample": arm.sample_text() + "; best of 2 steps"}
        except Exception as e:  # the checker is optional equipment of the bench, the product is not
            cpu = {"value": None, "unit": UNIT, "cores": cpu_cores(), "kind": "unavailable", "sample": repr(e)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"frangi3d sigma={a.sigmas} zdist=2 alpha=beta=0.5 C=500 on {w}x{h}x{l} uint8 "
                        f"(BASELINE.json {config_name(w, h, l, sigmas)}), z-slabs over {world} GPU(s)",
            "input": f"seeded {BASE_BLOCK[0]}x{BASE_BLOCK[1]}x{BASE_BLOCK[2]} synthetic neuron block tiled to the volume",
            "smoothing": "exact (rounded mul+add, bit-identical to the reference)" if a.exact
                         else "fma (within BASELINE tolerance; --exact for the bit-identical mode)",
            "l2": "inputs larger than L2 (per-scale working set >= 9 B/voxel x own voxels)",
            "timing": "CUDA events on the library stream around K back-to-back runs, max over ranks",
        },
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        "jmin": jmin, "jmax": jmax,
    }
    if exact_ms is not None:
        line["exact_ms_per_step"] = exact_ms
    if seed_prepass is not None:
        line["seed_prepass"] = seed_prepass
    if verify is not None:
        line["verify"] = verify
    emit(line)
    plan.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def _claim_stdout():
    """Libraries in this process (NCCL prints its version banner to stdout) must not add lines
    to the one-JSON-line contract: fd 1 is pointed at stderr and the JSON goes to the saved fd."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


_JSON_OUT = None


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    a = parse_args()
    w, h, l = (int(x) for x in a.workload.lower().split("x"))
    sigmas = [float(x) for x in a.sigmas.split(",")]
    if a.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: relaunch under torchrun the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    _JSON_OUT = _claim_stdout()
    if a.impl == "reference":
        return run_reference_arm(a, sigmas, w, h, l)
    return run_ours(a, sigmas, w, h, l)


if __name__ == "__main__":
    sys.exit(main())
