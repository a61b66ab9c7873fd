#!/usr/bin/env python
"""export_inputs.py — inputs of the Swift golden-vector harness (oracle/swift_ref/main.swift), one file per world:

    python oracle/swift_ref/export_inputs.py /tmp/cq_inputs          ->  /tmp/cq_inputs_{c1,c2,c3}.bin + .json manifests

  c1  the demo's static world; walk: 4 characters x 600 frames with the scripted intent of tests/golden/make_c1_trajectory.py
  c2  Semla render mesh + ground; 16,384 sweeps in each of the three modes, 8,192 overlaps / overlap-alls (8 and 3), 65,536 rays
  c3  ornate_mirror hulls + ground; walk: 4,096 characters x 6 steps (gravity on), plus 4,096 overlap-alls

Every part is written with WORLD-space vertices (computed by the C++ oracle's TriangleMeshSet.rebuild, bit-identical to
what the library uploads) and gets an identity transform on the Swift side.  TEST INFRASTRUCTURE (oracle/)."""
import importlib
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

cq = importlib.import_module("swift-game-engine_b200")
sc = cq.scenes


def world_space(part):
    w = orc.OracleWorld([dict(part, is_dynamic=False)])
    pos = w.read_soup(0)["positions"].copy()
    w.close()
    return pos  # every vertex of the part, in order (the degenerate filter drops triangles, never vertices)


def write_world(f, parts):
    f.write(struct.pack("<4sII", b"CQSW", 1, len(parts)))
    for p in parts:
        pos = world_space(p).astype("<f4")
        idx = np.ascontiguousarray(p["indices"], "<u4").reshape(-1)
        assert len(pos) == len(np.asarray(p["positions"]).reshape(-1, 3))
        f.write(struct.pack("<IIffBBHII", int(p["entity_id"]), int(p.get("layer", 1)), float(p.get("mu_s", 0.8)), float(p.get("mu_k", 0.6)),
                            int(bool(p.get("flatten_ground", False))), int(bool(p.get("is_dynamic", False))), 0, len(pos), len(idx)))
        f.write(pos.tobytes())
        f.write(idx.tobytes())


def walk(f, manifest, name, pos, vel, steps, params, intent=None, accel=20.0, decel=36.0):
    n = len(pos)
    f.write(struct.pack("<IIIf3fIffI", 5, n, steps, np.float32(1.0 / 60.0), 0.0, -98.0, 0.0, 1, accel, decel, 0 if intent is None else 1))
    f.write(np.ascontiguousarray(params, cq.PARAMS).tobytes())
    f.write(np.concatenate([np.asarray(pos, "<f4"), np.asarray(vel, "<f4")], axis=1).tobytes())
    if intent is not None:
        f.write(np.ascontiguousarray(intent, "<f4").tobytes())  # steps x n x 3
    manifest.append({"kind": "walk", "name": name, "n": n, "steps": steps})


def export(prefix):
    # ---- c1
    man = []
    with open(prefix + "_c1.bin", "wb") as f:
        parts = sc.c1_scene()
        write_world(f, parts)
        f.write(struct.pack("<I", 1))
        frames, n = 600, len(sc.C1_STARTS)
        yaw = 2.0 * np.pi * np.arange(frames) / frames
        spd = np.asarray(sc.C1_SPEEDS, np.float64)
        intent = np.zeros((frames, n, 3), np.float32)
        intent[:, :, 0] = np.float32(np.cos(yaw)[:, None] * spd[None, :])
        intent[:, :, 2] = np.float32(np.sin(yaw)[:, None] * spd[None, :])
        walk(f, man, "c1", sc.C1_STARTS, np.zeros_like(sc.C1_STARTS), frames, cq.default_params(), intent)
    json.dump({"parts": [int(p["entity_id"]) for p in parts], "scenarios": man}, open(prefix + "_c1.json", "w"))
    # ---- c2
    man = []
    with open(prefix + "_c2.bin", "wb") as f:
        parts = sc.semla_scene(use_hulls=False)
        write_world(f, parts)
        lo, hi = sc.scene_aabb(parts[1:])
        f.write(struct.pack("<I", 7))
        q = sc.gen_casts(16384, lo, hi, seed=0xC0111DE2)
        for mode in (0, 1, 2):
            f.write(struct.pack("<III", 1, mode, len(q)))
            f.write(q.tobytes())
            man.append({"kind": "casts", "mode": mode, "n": len(q), "seed": 0xC0111DE2})
        c = sc.gen_capsules(8192, lo, hi, seed=8, expand=0.5)
        f.write(struct.pack("<II", 2, len(c)))
        f.write(c.tobytes())
        man.append({"kind": "overlap", "n": len(c)})
        for mh in (8, 3):
            f.write(struct.pack("<III", 3, mh, len(c)))
            f.write(c.tobytes())
            man.append({"kind": "overlap_all", "max_hits": mh, "n": len(c)})
        r = sc.gen_rays(65536, lo, hi, seed=0xC0111DE5, expand=5.0)
        f.write(struct.pack("<II", 4, len(r)))
        f.write(r.tobytes())
        man.append({"kind": "rays", "n": len(r)})
    json.dump({"parts": [int(p["entity_id"]) for p in parts], "scenarios": man}, open(prefix + "_c2.json", "w"))
    # ---- c3
    man = []
    with open(prefix + "_c3.bin", "wb") as f:
        parts = sc.mirror_scene(use_hulls=True)
        write_world(f, parts)
        lo, hi = sc.scene_aabb(parts[1:])
        f.write(struct.pack("<I", 2))
        pos, vel = sc.gen_c3_characters(4096, seed=0xC0111DE3)
        pos[:512, 1] += 3.0   # some start in the air
        pos[512:1024, 1] -= 0.4  # some start inside the ground: depenetration
        walk(f, man, "c3", pos, vel, 6, cq.default_params())
        c = sc.gen_capsules(4096, lo, hi, seed=8)
        f.write(struct.pack("<III", 3, 8, len(c)))
        f.write(c.tobytes())
        man.append({"kind": "overlap_all", "max_hits": 8, "n": len(c)})
    json.dump({"parts": [int(p["entity_id"]) for p in parts], "scenarios": man}, open(prefix + "_c3.json", "w"))


if __name__ == "__main__":
    export(sys.argv[1] if len(sys.argv) > 1 else "/tmp/cq_inputs")
