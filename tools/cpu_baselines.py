#!/usr/bin/env python
"""CPU-baseline table of BASELINE.md §4: the restated reference (oracle/, reference BVH + DFS order) on THIS host, with
1 thread (the reference is single-threaded, MainActor) and with all host cores (independent queries partitioned over
std::threads), for the five configurations of BASELINE.json at bounded sample sizes.  No GPU, no CUDA call.

    python tools/cpu_baselines.py [--quick] > profiles/rN_cpu_baselines_<host>.txt

This is measurement tooling on top of the oracle (like bench.py's cpu_baseline leg); nothing in the product imports it.
"""
import argparse
import importlib
import os
import platform
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, repeat=1):
    best = None
    for _ in range(repeat):
        t0 = time.perf_counter()
        out = fn()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="small terrain (cells 512) and smaller samples")
    a = ap.parse_args()
    from oracle import oracle as orc
    orc.build()
    cq = importlib.import_module("swift-game-engine_b200")  # scenes only (numpy)
    sc = cq.scenes
    cores = os.cpu_count() or 1
    print(f"# restated reference (C++, g++ -O2 -ffp-contract=off), host: {platform.node()} {platform.processor() or ''} "
          f"{cores} logical cores; every figure = units / wall seconds of the call")
    rows = []

    # C1: the golden-trajectory run (4 characters x 600 frames over the demo static world), 1 thread; the time of the
    # oracle calls alone (the scripted intent around them is Python and is not counted)
    parts = sc.c1_scene()
    w = orc.OracleWorld(parts)
    states = orc.init_states(sc.C1_STARTS)
    params = orc.default_params()
    spent = [0.0]

    def c1_step(s):
        t0 = time.perf_counter()
        w.move_and_slide(s, params, order=orc.ORDER_REFERENCE)
        spent[0] += time.perf_counter() - t0
    sc.c1_run(c1_step, states, 600, sc.C1_SPEEDS)
    n_char = len(sc.C1_STARTS)
    rows.append((f"C1 demo world, {n_char} characters x 600 frames", 1, f"{spent[0] / (600 * n_char) * 1e6:.1f} us per character-frame "
                 f"({600 * n_char / spent[0]:,.0f} queries/s)"))
    w.close()

    # C2: capsule sweeps vs Semla render mesh at its demo placement
    parts = sc.semla_scene(use_hulls=False)
    w = orc.OracleWorld(parts)
    lo, hi = sc.scene_aabb(parts[1:])
    for radius, hh, n1, nall in ((1.5, 1.0, 2048, 16384), (0.4, 0.5, 8192, 65536)):
        q = sc.gen_casts(nall, lo, hi, seed=0xC0111DE2, radius=radius, half_height=hh)
        if a.quick:
            n1, nall = n1 // 4, nall // 4
        t1, _ = timed(lambda: w.capsule_cast(q[:n1], 0, orc.ORDER_REFERENCE, 1))
        ta, _ = timed(lambda: w.capsule_cast(q[:nall], 0, orc.ORDER_REFERENCE, cores))
        rows.append((f"C2 capsuleCast vs Semla render mesh (50 k tris), r={radius} hh={hh}", 1, f"{n1 / t1:,.0f} sweeps/s ({n1} sweeps)"))
        rows.append(("", cores, f"{nall / ta:,.0f} sweeps/s ({nall} sweeps)"))
    w.close()

    # C3: move-and-slide step, mirror hulls / render mesh + ground plane
    for mesh, n1, nall in (("hulls", 65536, 1 << 20), ("render", 2048, 32768)):
        if a.quick:
            n1, nall = n1 // 4, nall // 8
        parts = sc.mirror_scene(use_hulls=mesh == "hulls")
        w = orc.OracleWorld(parts)
        pos, vel = sc.gen_c3_characters(nall, seed=0xC0111DE3)
        s_all = orc.init_states(pos, vel)
        for _ in range(2):  # the bench's two warm-up steps: characters land before the timed step
            w.move_and_slide(s_all, params, order=orc.ORDER_REFERENCE, n_threads=cores)
        s1 = s_all[:n1].copy()
        t1, _ = timed(lambda: w.move_and_slide(s1, params, order=orc.ORDER_REFERENCE, n_threads=1))
        ta, _ = timed(lambda: w.move_and_slide(s_all, params, order=orc.ORDER_REFERENCE, n_threads=cores))
        rows.append((f"C3 move-and-slide step, ornate_mirror {mesh} + ground", 1, f"{n1 / t1:,.0f} queries/s ({n1} characters)"))
        rows.append(("", cores, f"{nall / ta:,.0f} queries/s ({nall} characters)"))
        w.close()

    # C4: blocking sweeps over the procedural terrain (reference BVH build timed separately)
    cells = 512 if a.quick else 2236
    parts, half = sc.terrain_scene(cells=cells, cell=2.0)
    tb, w = timed(lambda: orc.OracleWorld(parts))
    ntri = w.counts(0)["triangles"]
    rows.append((f"C4 reference BVH build (top-down median split), {ntri:,} triangles", 1, f"{tb:.1f} s"))
    for radius, hh in ((0.4, 0.5), (1.5, 1.0)):
        q = sc.gen_c4_casts(65536, half, seed=0xC0111DE4, radius=radius, half_height=hh)
        t1, _ = timed(lambda: w.capsule_cast(q[:16384], 1, orc.ORDER_REFERENCE, 1))
        ta, _ = timed(lambda: w.capsule_cast(q, 1, orc.ORDER_REFERENCE, cores))
        rows.append((f"C4 capsuleCastBlocking over the terrain, r={radius} hh={hh}", 1, f"{16384 / t1:,.0f} sweeps/s (16384 sweeps)"))
        rows.append(("", cores, f"{65536 / ta:,.0f} sweeps/s (65536 sweeps)"))
    w.close()
    del w, parts

    # C5: rays + refit of the spinning mirror, three meshes merged
    parts = sc.merged_scene(mirror_dynamic=True)
    w = orc.OracleWorld(parts)
    lo, hi = sc.scene_aabb(parts[1:])
    rays = sc.gen_rays(262144, lo, hi, seed=0xC0111DE5, max_distance=100.0, expand=5.0, y_range=(0.0, 12.0))
    fix = sc.load_mirror_fixture()
    base_t, base_q, base_s = sc.transform_from_matrix(sc.mirror_model(fix["transform"]))
    mirror_id = parts[-1]["entity_id"]
    rot = sc.quat_mul(sc.quat_angle_axis(np.radians(1.0), (0, 1, 0)), base_q)
    tr, _ = timed(lambda: w.update_transforms([mirror_id], [sc.trs_model(base_t, rot, base_s)]), repeat=3)
    t1, _ = timed(lambda: w.raycast(rays[:65536], orc.ORDER_REFERENCE, 1))
    ta, _ = timed(lambda: w.raycast(rays, orc.ORDER_REFERENCE, cores))
    rows.append((f"C5 refit of the mirror part ({w.counts(1)['triangles']:,} of {w.counts(0)['triangles'] + w.counts(1)['triangles']:,} triangles)",
                 1, f"{tr * 1e3:.2f} ms"))
    rows.append(("C5 raycast, three meshes merged", 1, f"{65536 / t1:,.0f} rays/s (65536 rays)"))
    rows.append(("", cores, f"{262144 / ta:,.0f} rays/s (262144 rays)"))
    w.close()

    width = max(len(r[0]) for r in rows)
    for name, threads, value in rows:
        print(f"{name:<{width}}  {threads:>3} thr  {value}")


if __name__ == "__main__":
    main()
