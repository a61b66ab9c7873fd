"""Regenerates tests/golden/ornate_mirror.npz from the reference's asset.

Run in the build container only (it reads /root/reference, which does not exist on
the GPU box):  python tests/golden/make_mesh_fixture.py
The fixture holds the DATA of Game/ornate_mirror.static.json (schema:
StaticMeshLoader.swift:168-197) narrowed to float32 exactly as the reference's
loader narrows it (JSON double -> Float), so tests and bench.py can build the
mirror scene without the reference tree.
"""
import json
import os
import sys

import numpy as np

SRC = "/root/reference/Game/ornate_mirror.static.json"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ornate_mirror.npz")


def main():
    d = json.load(open(SRC))
    assert d["version"] == 1 and len(d["meshes"]) == 1
    m = d["meshes"][0]
    out = {
        "name": np.array(m["name"]),
        "transform_rowmajor": np.asarray(m["transform"], np.float64).astype(np.float32),
        "positions": np.asarray(m["mesh"]["positions"], np.float64).astype(np.float32).reshape(-1, 3),
        "indices": np.asarray(m["mesh"]["indices"], np.uint32),
        "n_hulls": np.array(len(m.get("collisionHulls") or [])),
    }
    for i, h in enumerate(m.get("collisionHulls") or []):
        out[f"hull{i}_positions"] = np.asarray(h["positions"], np.float64).astype(np.float32).reshape(-1, 3)
        out[f"hull{i}_indices"] = np.asarray(h["indices"], np.uint32)
    if out["indices"].max() <= 65535:
        out["indices"] = out["indices"].astype(np.uint16)
    np.savez_compressed(DST, **out)
    print("wrote", DST, os.path.getsize(DST), "bytes;", out["positions"].shape[0], "verts,",
          out["indices"].shape[0] // 3, "tris,", int(out["n_hulls"]), "hulls")


if __name__ == "__main__":
    sys.exit(main())
