"""What the box's PCIe / host memory gives when every rank copies at once: each of the N ranks moves its share of one
step's end-to-end traffic of the 2048 x 2048 x 512 workload (2 GiB up, 14 GiB down in total) between its GPU and pinned
host memory, with nothing else running.  The slowest rank's time is the floor of `e2e` at N GPUs on this box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29577 tools/d2h_concurrent.py

Prints one JSON line on rank 0.  Ad hoc measurement tool (bench.py is the contract)."""
import json, os, time
import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    up = (2 << 30) // world
    down = (14 << 30) // world
    h_up = torch.empty(up, dtype=torch.uint8).pin_memory()
    h_down = torch.empty(down, dtype=torch.uint8).pin_memory()
    d_up = torch.empty(up, dtype=torch.uint8, device="cuda")
    d_down = torch.empty(down, dtype=torch.uint8, device="cuda")
    h_up.fill_(1); h_down.fill_(1); d_down.fill_(2)
    s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
    times = []
    for it in range(6):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s_up):             # both directions at once, as the chunk-pipelined call moves them
            d_up.copy_(h_up, non_blocking=True)
        with torch.cuda.stream(s_down):
            h_down.copy_(d_down, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    if rank == 0:
        best = min(times[1:])
        print(json.dumps({"n_gpus": world, "bytes_up_total": up * world, "bytes_down_total": down * world,
                          "ms_slowest_rank_best_of_5": 1e3 * best, "aggregate_GB_per_s": (up + down) * world / best / 1e9,
                          "all_ms": [round(1e3 * x, 1) for x in times]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
