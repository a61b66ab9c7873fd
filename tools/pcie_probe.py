#!/usr/bin/env python3
"""Host<->device copy ceiling of this box with every rank copying at once (explains bench.py's e2e at N > 1):
each rank times H2D, D2H and both directions together for 176 MB pinned buffers (the C3 state arrays)."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 176 * 1024 * 1024
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(mode, iters=10):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return n * iters / dt / 1e9


for mode in ("h2d", "d2h", "both"):
    run(mode, 2)
    g = run(mode)
    t = torch.tensor([g], device="cuda")
    if world > 1:
        lst = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(lst, t)
    else:
        lst = [t]
    if rank == 0:
        print(mode, "GB/s per direction per rank:", " ".join("%.1f" % x.item() for x in lst), flush=True)
if rank == 0:
    try:
        p = torch.cuda.get_device_properties(0)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        print("gpu0 numa_node:", open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
    except Exception as e:
        print("numa_node unavailable:", e)
    print("cpus:", os.cpu_count(), "affinity:", len(os.sched_getaffinity(0)))
    os.system("ls /sys/devices/system/node/ | head; nvidia-smi topo -m 2>&1 | head -14")
if world > 1:
    dist.destroy_process_group()
