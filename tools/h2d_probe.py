"""Concurrent pinned-host copy bandwidth of every rank of one box (one process per GPU, torchrun), with and without the
rank-to-core / NUMA binding bench.py applies (bind_rank_to_host).  Answers VERDICT r01 item 4: does the 8-rank collapse of
H2D bandwidth (12.5 GB/s per rank) come from all ranks sharing one socket's cores and memory?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe.py
Prints one JSON line per mode on rank 0: per-rank H2D / D2H GB/s (both directions at once, 256 MiB each, best of 5)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench


def measure(dev, nb=256 << 20, reps=5):
    hb, hb2 = torch.empty(nb, dtype=torch.uint8).pin_memory(), torch.empty(nb, dtype=torch.uint8).pin_memory()
    hb.fill_(1)
    db, db2 = torch.empty(nb, dtype=torch.uint8, device=dev), torch.empty(nb, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    e = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(2)]
    best = [0.0, 0.0]
    for _ in range(reps):
        if dist.is_initialized():
            dist.barrier()
        torch.cuda.synchronize()
        with torch.cuda.stream(s1):
            e[0][0].record()
            db.copy_(hb, non_blocking=True)
            e[0][1].record()
        with torch.cuda.stream(s2):
            e[1][0].record()
            hb2.copy_(db2, non_blocking=True)
            e[1][1].record()
        torch.cuda.synchronize()
        for k in range(2):
            best[k] = max(best[k], nb / (e[k][0].elapsed_time(e[k][1]) * 1e-3) / 1e9)
    return best


def main():
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    for mode in ("unbound", "bound"):
        info = None
        if mode == "bound":
            info = bench.bind_rank_to_host(local, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
        h2d, d2h = measure(dev)
        t = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
        if world > 1:
            allv = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allv, t)
        else:
            allv = [t]
        infos = [None] * world
        if world > 1:
            dist.all_gather_object(infos, info)
        else:
            infos = [info]
        if rank == 0:
            print(json.dumps({"mode": mode, "ranks": world, "h2d_gbs": [round(float(v[0]), 1) for v in allv],
                              "d2h_gbs": [round(float(v[1]), 1) for v in allv],
                              "h2d_total": round(sum(float(v[0]) for v in allv), 1), "binding": infos,
                              "cpus_allowed": len(os.sched_getaffinity(0)),
                              "mems_allowed": [l.split(":")[1].strip() for l in open("/proc/self/status") if l.startswith("Mems_allowed_list")]}),
                  flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
