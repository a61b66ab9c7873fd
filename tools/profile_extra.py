#!/usr/bin/env python
"""Workloads for the kernels bench.py has no line for, so that every kernel of the path has an ncu capture under profiles/:

    python tools/profile_extra.py overlap      # k_capsule_overlap_pool<ALL>: 1,048,576 capsuleOverlapAll(8) vs the mirror render mesh
    python tools/profile_extra.py separation   # k_sep_turns / k_sep_post: AgentSeparationSystem, 262,144 agents at 30% coverage on a terrain
    python tools/profile_extra.py build        # k_os_pass (onesweep), k_karras, k_fit ...: cq_world_create of the 10 M-triangle terrain (canonical order)

Each mode runs its work three times (warm-up included) and prints one line with CUDA-event timings."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

cq = importlib.import_module("swift-game-engine_b200")
sc = cq.scenes


def main(mode):
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)

    def up(a):
        return torch.from_numpy(np.frombuffer(a.tobytes(), np.uint8).copy()).to(dev)

    if mode == "overlap":
        parts = sc.mirror_scene(use_hulls=False)
        g = cq.CollisionQuery(parts)
        lo, hi = sc.scene_aabb(parts[1:])
        n = 1 << 20
        c = sc.gen_capsules(n, lo, hi, seed=8)
        d_c = up(c)
        d_out = torch.empty(n * 8 * cq.OVERLAP_HIT.itemsize, dtype=torch.uint8, device=dev)
        d_cnt = torch.empty(n, dtype=torch.int32, device=dev)
        d_ov = torch.empty(n, dtype=torch.uint8, device=dev)
        L = cq.lib()
        L.cq_capsule_overlap_all_device.argtypes = [cq.C.c_void_p] * 2 + [cq.C.c_int32] * 2 + [cq.C.c_void_p] * 4
        ms = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = L.cq_capsule_overlap_all_device(g.handle, d_c.data_ptr(), n, 8, d_out.data_ptr(), d_cnt.data_ptr(), d_ov.data_ptr(), st.cuda_stream)
            assert rc == 0
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        print(f"overlap-all(8): {n} capsules vs {g.info()['n_static_triangles']} triangles: {ms} ms; overflow {float(d_ov.float().mean()):.3f}, "
              f"mean hits {float(d_cnt.float().mean()):.2f}")
    elif mode == "separation":
        parts, half = sc.terrain_scene(cells=700, cell=2.0)
        g = cq.CollisionQuery(parts)
        n = 1 << 18
        rng = np.random.default_rng(3)
        span = float(np.sqrt(n * np.pi * 0.4 ** 2 / 0.3)) / 2
        x = rng.uniform(-span, span, (n, 2))
        y = sc.terrain_height(x[:, 0], x[:, 1]) + np.float32(1.1)
        pos = np.stack([x[:, 0], y, x[:, 1]], axis=1).astype(np.float32)
        h, s = rng.uniform(0, 2 * np.pi, n), rng.uniform(0, 4.5, n)
        vel = np.stack([np.cos(h) * s, np.zeros(n), np.sin(h) * s], axis=1).astype(np.float32)
        p = cq.default_params(radius=0.4, half_height=0.5, skin_width=0.08)
        d = up(cq.init_states(pos, vel))
        ms = []
        for _ in range(3):
            g.move_and_slide_device(d.data_ptr(), n, p, flags=3, stream=st.cuda_stream)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            g.agent_separation_device(d.data_ptr(), n, p, stream=st.cuda_stream)
            torch.cuda.synchronize()
            ms.append((time.perf_counter() - t0) * 1e3)
        print(f"agent separation: {n} agents at 30% coverage on {g.info()['n_static_triangles']} triangles: {ms} ms per call (2 sweeps + post)")
    elif mode == "build":
        parts, half = sc.terrain_scene()
        ms = []
        for _ in range(3):
            t0 = time.perf_counter()
            g = cq.CollisionQuery(parts, order=cq.ORDER_CANONICAL)
            wall = time.perf_counter() - t0
            ms.append((g.info()["build_ms"], wall))
            g.close()
        print(f"cq_world_create of {len(parts[0]['indices']) // 3} triangles (canonical order): (kernel ms, wall s) {ms}")
    else:
        raise SystemExit(__doc__)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "")
