"""world_size-2 (and 3) CPU tests of the multi-GPU host logic over torch.distributed/gloo:
the slab partition, the per-scale halo plan (which xy-smoothed planes go where) and the
Jmin/Jmax reduction -- the same choreography frangi_gpu.cu runs over NCCL send/recv, with
the CPU oracle standing in for the kernels.  The assembled result must equal the oracle on
the whole volume bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pnr_b200.slabs import exchange_plan, halo_planes, max_slabs, slab_ranges, z_radius

SIG = [2.0, 4.0, 6.0]
ZD = 2.0
SHAPE = (36, 24, 28)   # l, h, w


def test_plan_arithmetic():
    assert [z_radius(s, 2.0) for s in (1, 2, 3, 4, 5, 6)] == [2, 3, 5, 6, 8, 9]
    assert [halo_planes(s, 2.0) for s in (2, 4, 6)] == [5, 8, 11]
    assert halo_planes(6, 1.0) == 20                                                  # "3 sigma_max" is not enough at zdist 1
    assert slab_ranges(512, 8) == [(64 * k, 64 * k + 64) for k in range(8)]
    assert slab_ranges(10, 3) == [(0, 3), (3, 6), (6, 10)]
    assert max_slabs(512, SIG, 2.0) == 46 and max_slabs(20, SIG, 2.0) == 1
    p = exchange_plan(512, 8, 3, 6.0, 2.0)
    assert p == dict(send_down=(192, 203), recv_down=(181, 192), send_up=(245, 256), recv_up=(256, 267))
    assert exchange_plan(512, 8, 0, 6.0, 2.0)["send_down"] is None
    assert exchange_plan(512, 8, 7, 6.0, 2.0)["recv_up"] is None
    with pytest.raises(ValueError):
        exchange_plan(40, 8, 1, 6.0, 2.0)                                             # 5-plane slabs cannot feed an 11-plane halo


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import Oracle
    from pnr_b200.synth import make_volume
    o = Oracle()
    l, h, w = SHAPE
    z0, z1 = slab_ranges(l, world)[rank]
    own = make_volume(w, h, l, seed=21, n_neurites=4, z_range=(z0, z1))      # a rank generates only its slab
    Hmax = max(halo_planes(s, ZD) for s in SIG)
    lo, hi = max(z0 - Hmax, 0), min(z1 + Hmax, l)
    ext = np.zeros((hi - lo, h, w), np.uint8)
    ext[z0 - lo:z1 - lo] = own
    # halo exchange with the z neighbours, as planned per scale (here once, for the widest scale, on the u8 planes)
    plan = exchange_plan(l, world, rank, max(SIG), ZD)
    reqs = []
    if plan["send_down"]:
        a, b = plan["send_down"]
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(own[a - z0:b - z0])), rank - 1))
    if plan["send_up"]:
        a, b = plan["send_up"]
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(own[a - z0:b - z0])), rank + 1))
    if plan["recv_down"]:
        a, b = plan["recv_down"]
        t = torch.empty((b - a, h, w), dtype=torch.uint8)
        dist.recv(t, rank - 1)
        ext[a - lo:b - lo] = t.numpy()
    if plan["recv_up"]:
        a, b = plan["recv_up"]
        t = torch.empty((b - a, h, w), dtype=torch.uint8)
        dist.recv(t, rank + 1)
        ext[a - lo:b - lo] = t.numpy()
    for r in reqs:
        r.wait()
    res = o.frangi3d(ext, SIG, ZD, want_scale=False, want_dir=False)
    J = res["J"][z0 - lo:z1 - lo]
    # global Jmin / Jmax: the all-reduce that follows the last scale (frangi.cpp:237-238,257-258)
    mm = torch.tensor([float(J.min()), -float(J.max())], dtype=torch.float64)
    dist.all_reduce(mm, op=dist.ReduceOp.MIN)
    q.put((rank, z0, z1, J.copy(), res["Vx"][z0 - lo:z1 - lo].copy(), float(mm[0]), -float(mm[1])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_choreography_matches_whole_volume(world, oracle):
    from pnr_b200.synth import make_volume
    l, h, w = SHAPE
    I = make_volume(w, h, l, seed=21, n_neurites=4)
    whole = oracle.frangi3d(I, SIG, ZD, want_scale=False, want_dir=False)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    J = np.empty_like(whole["J"])
    Vx = np.empty_like(whole["Vx"])
    for rank, z0, z1, Js, Vs, jmin, jmax in got:
        J[z0:z1] = Js
        Vx[z0:z1] = Vs
        assert jmax == whole["Jmax"] and jmin == whole["Jmin"]
    assert np.array_equal(J, whole["J"]) and np.array_equal(Vx, whole["Vx"])
