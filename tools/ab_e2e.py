"""A/B of two builds of libcq.so on the host-pointer (e2e) move-and-slide call, same box, interleaved processes.

    python tools/ab_e2e.py swift-game-engine_b200/csrc/libcq.so swift-game-engine_b200/csrc/libcq_prev.so

Each library runs in its own subprocess (no torch): C3 hulls scene, 1,048,576 characters in pinned host memory, 3 warm-up
calls, then `--steps` timed calls of cq_move_and_slide_batch with the state carried; the libraries alternate `--rounds`
times so that box-level drift (PCIe / NUMA placement, clocks) shows up as spread within a library rather than as a
difference between them.  Prints one line per run and the medians.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(lib_path, n, steps):
    sys.path.insert(0, ROOT)
    cq = importlib.import_module("swift-game-engine_b200")
    cq.LIB_PATH = os.path.abspath(lib_path)  # before the first cq.lib() call
    sc = cq.scenes
    world = cq.CollisionQuery(sc.mirror_scene(use_hulls=True))
    pos, vel = sc.gen_c3_characters(n, seed=0xC0111DE3)
    pinned = cq.PinnedArray((n,), cq.STATE)
    pinned.array[:] = cq.init_states(pos, vel)
    params = cq.default_params()
    for _ in range(3):
        world.move_and_slide(pinned.array, params)
    per_call = []
    for _ in range(steps):
        t0 = time.perf_counter()
        world.move_and_slide(pinned.array, params)
        per_call.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps({"lib": lib_path, "version": cq.lib().cq_version().decode(), "ms_median": statistics.median(per_call),
                      "ms_min": min(per_call), "ms_mean": sum(per_call) / len(per_call),
                      "grounded": float(pinned.array["grounded"].mean())}))
    world.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+")
    ap.add_argument("--chars", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--worker", default=None)
    a = ap.parse_args()
    if a.worker:
        worker(a.worker, a.chars, a.steps)
        return
    results = {lib: [] for lib in a.libs}
    for _ in range(a.rounds):
        for lib in a.libs:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", lib, "--chars", str(a.chars),
                                  "--steps", str(a.steps), lib], capture_output=True, text=True, timeout=300)
            if out.returncode != 0:
                print(json.dumps({"lib": lib, "error": out.stderr[-500:]}))
                continue
            line = out.stdout.strip().splitlines()[-1]
            print(line, flush=True)
            results[lib].append(json.loads(line)["ms_median"])
    for lib, v in results.items():
        if v:
            print(json.dumps({"lib": lib, "median_of_medians_ms": statistics.median(v), "runs": v}))


if __name__ == "__main__":
    main()
