"""Debug aid: per-plane comparison of an N-rank NCCL slab job against a one-GPU run (torchrun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import pnr_b200
from pnr_b200 import FrangiPlan
import bench

def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    w, h, l = (int(v) for v in sys.argv[1].split("x"))
    reps = int(sys.argv[2]); sync = int(sys.argv[3])
    sig = [2., 4., 6.]
    lib = pnr_b200.load_library()
    import ctypes as C
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        buf = C.create_string_buffer(128); lib.frangi_gpu_nccl_unique_id(buf)
        t.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
    dist.broadcast(t, 0)
    uid = bytes(t.cpu().numpy().tobytes())
    z0, z1 = l * rank // world, l * (rank + 1) // world
    I = bench.workload_slab(w, h, l, z0, z1)
    plan = FrangiPlan(sig, 2.0, .5, .5, 500., False, w, h, l, flags=1, slab=(z0, z1, rank, world, uid, lr))
    plan.upload(I)
    for _ in range(reps):
        plan.run_resident(sync=bool(sync))
    plan.sync(); dist.barrier()
    out = plan.download()
    whole = FrangiPlan(sig, 2.0, .5, .5, 500., False, w, h, l, flags=1, devices=(lr,))
    whole.upload(bench.workload_slab(w, h, l, 0, l)); whole.run_resident()
    ref = whole.download(); whole.close()
    for k in ("J", "Vx"):
        d = out[k] != ref[k][z0:z1]
        per = d.reshape(d.shape[0], -1).sum(1)
        bad = np.nonzero(per)[0]
        msg = f"rank {rank} {k}: bad planes (local z: count) " + ", ".join(f"{z}:{per[z]}" for z in bad[:24])
        if len(bad):
            zz = bad[0]; ys, xs = np.nonzero(d[zz])
            msg += f" | first bad plane {zz}: y {ys.min()}..{ys.max()} x {xs.min()}..{xs.max()} n={len(ys)}; x%128 hist {np.bincount(xs % 128 // 32, minlength=4)}"
        print(msg, flush=True)
    plan.close(); dist.destroy_process_group()

main()
