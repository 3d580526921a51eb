#!/usr/bin/env python3
"""Host-to-device bandwidth per rank, alone and all ranks together (torchrun, one process per GPU): where the end-to-end
path of scpd_decode_host saturates on this box.  Prints one JSON line from rank 0."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def numa_of_gpu(i):
    try:
        bus = torch.cuda.get_device_properties(i).pci_bus_id if hasattr(torch.cuda.get_device_properties(i), "pci_bus_id") else None
    except Exception:
        bus = None
    try:
        import subprocess
        out = subprocess.run(["nvidia-smi", "-i", str(i), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             stdout=subprocess.PIPE, text=True, timeout=20).stdout.strip().lower()
        dom = out[4:] if out.startswith("0000") else out
        return int(open(f"/sys/bus/pci/devices/{dom}/numa_node").read())
    except Exception:
        return None


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = 1 << 30
    h = torch.empty(nbytes, dtype=torch.int8, pin_memory=True)
    h.fill_(1)
    d = torch.empty(nbytes, dtype=torch.int8, device="cuda")
    o = torch.empty(nbytes // 8, dtype=torch.int8, pin_memory=True)

    def h2d(reps=4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        return reps * nbytes / (time.perf_counter() - t0) / 1e9

    def duplex(reps=4):
        s2 = torch.cuda.Stream()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
            with torch.cuda.stream(s2):
                o.copy_(d[: nbytes // 8], non_blocking=True)
        torch.cuda.synchronize()
        return reps * nbytes / (time.perf_counter() - t0) / 1e9

    h2d(1)
    alone = [0.0] * world
    for r in range(world):  # one rank at a time
        if world > 1:
            dist.barrier()
        if r == rank:
            alone[r] = h2d()
    if world > 1:
        dist.barrier()
    together = h2d()
    if world > 1:
        dist.barrier()
    dup = duplex()
    info = {"rank": rank, "alone_gbs": alone[rank], "together_gbs": together, "together_duplex_gbs": dup,
            "gpu_numa_node": numa_of_gpu(local), "cpu_affinity": sorted(os.sched_getaffinity(0))[:4] + ["..."] + [len(os.sched_getaffinity(0))]}
    if world > 1:
        allinfo = [None] * world
        dist.all_gather_object(allinfo, info)
    else:
        allinfo = [info]
    if rank == 0:
        nodes = []
        try:
            nodes = sorted(x for x in os.listdir("/sys/devices/system/node") if x.startswith("node"))
        except OSError:
            pass
        print(json.dumps({"world": world, "bytes_per_copy": nbytes, "host_numa_nodes": nodes,
                          "sum_alone_gbs": sum(x["alone_gbs"] for x in allinfo),
                          "sum_together_gbs": sum(x["together_gbs"] for x in allinfo), "ranks": allinfo}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
