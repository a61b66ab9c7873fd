#!/usr/bin/env python
"""Probe: what would sorting the rays of C5 (direction octant, then Morton code of the origin) buy k_raycast_phased?

The rays are sorted on the HOST here, outside every timed region -- this measures the ceiling of a device-side ray sort before
writing one.  Prints the kernel time of cq_raycast_device for the generated order and the sorted order, both order rules."""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

cq = importlib.import_module("swift-game-engine_b200")
sc = cq.scenes


def part1by2(v):
    v = v.astype(np.uint64) & 0x3FF
    v = (v | (v << 16)) & 0x30000FF
    v = (v | (v << 8)) & 0x300F00F
    v = (v | (v << 4)) & 0x30C30C3
    v = (v | (v << 2)) & 0x9249249
    return v


def sort_key(rays, lo, hi):
    o = rays["origin"]
    d = rays["direction"]
    q = np.clip((o - lo) / np.maximum(hi - lo, 1e-6), 0, 1)
    q = np.minimum((q * 1024).astype(np.int64), 1023)
    m = part1by2(q[:, 0]) | (part1by2(q[:, 1]) << 1) | (part1by2(q[:, 2]) << 2)
    octant = ((d[:, 0] < 0).astype(np.uint64)) | ((d[:, 1] < 0).astype(np.uint64) << 1) | ((d[:, 2] < 0).astype(np.uint64) << 2)
    return (octant << 30) | m


def main():
    n = int(os.environ.get("N_RAYS", str(1 << 24)))
    dev = torch.device("cuda:0")
    parts = sc.merged_scene(mirror_dynamic=True)
    lo, hi = sc.scene_aabb(parts[1:])
    rays = sc.gen_rays(n, lo, hi, seed=0xC0111DE5, max_distance=100.0, expand=5.0, y_range=(0.0, 12.0))
    print("ray dtype:", rays.dtype)
    o = rays["origin"]
    key = sort_key(rays, o.min(0), o.max(0))
    t0 = time.perf_counter()
    perm = np.argsort(key, kind="stable")
    print("host argsort %.2f s" % (time.perf_counter() - t0))
    srt = np.ascontiguousarray(rays[perm])
    for order in (cq.ORDER_REFERENCE, cq.ORDER_CANONICAL):
        world = cq.CollisionQuery(parts, order=order)
        outs = {}
        for name, r in (("generated", rays), ("sorted", srt)):
            d_r = torch.from_numpy(r.view(np.uint8).reshape(-1)).to(dev)
            d_out = torch.empty(n * cq.RAY_HIT.itemsize, dtype=torch.uint8, device=dev)
            ts = torch.cuda.Stream(device=dev)
            torch.cuda.set_stream(ts)
            s = ts.cuda_stream
            for _ in range(3):
                world.raycast_device(d_r.data_ptr(), n, d_out.data_ptr(), s)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                world.raycast_device(d_r.data_ptr(), n, d_out.data_ptr(), s)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            outs[name] = d_out.cpu().numpy().view(cq.RAY_HIT)
            print("order=%d %-9s %.3f ms  %.2f G rays/s" % (order, name, ms, n / ms / 1e6))
        a, b = outs["generated"][perm], outs["sorted"]
        print("  same answers after the permutation:", a.tobytes() == b.tobytes())
        del world


if __name__ == "__main__":
    main()
