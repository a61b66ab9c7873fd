#!/usr/bin/env python
"""Blender-free regenerator of the reference's `*.static.json` assets.

The reference produces its collision/render assets with a Blender script
(Tools/FbxToStaticMeshJson/export_static_mesh_json.py:131-239).  Two of the three named assets
(Game/17-Cheese.static.json, Game/Semla.static.json) are missing large blobs, but their FBX sources are
present under ExternalResources/.  This tool reads binary FBX (7.x) directly and emits the same schema
(StaticMeshLoader.swift:168-197):

  * mesh.positions / indices : FBX-space vertices verbatim, polygons triangulated (quads by Blender's own
    quad-flip rule: 14,220 of the 14,246 triangles of the shipped mirror asset come out as the same position
    triples; n-gons as a fan)
  * transform (row-major)    : Rx(+90 deg) . T . R . S with the centimetre -> metre factor, i.e. what Blender's
    importer puts in matrix_world for a Y-up file (validated against the shipped ornate_mirror.static.json:
    same triangle count, local AABB and transform)
  * collisionHulls           : STAND-INS — convex hulls (scipy) of the <= 2 largest connected components,
    greedily reduced to <= 24 faces.  Blender's convex-hull + decimate output cannot be reproduced bit for bit.

Regenerated assets are stand-ins, not the original blobs: vertex order (hence triangle numbering) differs
from what Blender's exporter would write because it re-indexes on (position, normal, uv).

  python tools/fbx_to_static_mesh.py <in.fbx> <out.static.json>            # the reference's schema
  python tools/fbx_to_static_mesh.py --fixtures [blender|shorter]           # tests/golden/{semla,cheese}.npz (default: Blender's quad rule)
"""
import json
import os
import struct
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/ExternalResources"


# ---------------------------------------------------------------- binary FBX reader
class Node:
    def __init__(self, name, props, children):
        self.name, self.props, self.children = name, props, children

    def find(self, name):
        return [c for c in self.children if c.name == name]

    def first(self, name):
        r = self.find(name)
        return r[0] if r else None


def _read_prop(buf, off):
    t = chr(buf[off])
    off += 1
    scal = {"Y": ("<h", 2), "C": ("<?", 1), "I": ("<i", 4), "F": ("<f", 4), "D": ("<d", 8), "L": ("<q", 8)}
    if t in scal:
        fmt, sz = scal[t]
        return struct.unpack_from(fmt, buf, off)[0], off + sz
    arr = {"f": np.float32, "d": np.float64, "l": np.int64, "i": np.int32, "b": np.uint8}
    if t in arr:
        n, enc, clen = struct.unpack_from("<III", buf, off)
        off += 12
        raw = bytes(buf[off:off + clen])
        if enc == 1:
            raw = zlib.decompress(raw)
        return np.frombuffer(raw, dtype=arr[t], count=n).copy(), off + clen
    if t in "SR":
        n = struct.unpack_from("<I", buf, off)[0]
        off += 4
        data = bytes(buf[off:off + n])
        return (data.decode("utf-8", "replace") if t == "S" else data), off + n
    raise ValueError(f"unknown FBX property type {t!r}")


def _read_node(buf, off, wide):
    if wide:
        end, nprops, plen = struct.unpack_from("<QQQ", buf, off)
        off += 24
    else:
        end, nprops, plen = struct.unpack_from("<III", buf, off)
        off += 12
    nlen = buf[off]
    off += 1
    if end == 0:
        return None, off + nlen
    name = bytes(buf[off:off + nlen]).decode("ascii", "replace")
    off += nlen
    props = []
    pend = off + plen
    for _ in range(nprops):
        p, off = _read_prop(buf, off)
        props.append(p)
    off = pend
    children = []
    sentinel = 25 if wide else 13
    while off < end:
        if end - off == sentinel and not any(buf[off:end]):
            break
        child, off = _read_node(buf, off, wide)
        if child is None:
            break
        children.append(child)
    return Node(name, props, children), end


def read_fbx(path):
    buf = memoryview(open(path, "rb").read())
    assert bytes(buf[:20]) == b"Kaydara FBX Binary  ", "not a binary FBX file"
    version = struct.unpack_from("<I", buf, 23)[0]
    wide = version >= 7500
    off = 27
    nodes = []
    while off < len(buf) - (25 if wide else 13):
        n, off = _read_node(buf, off, wide)
        if n is None:
            break
        nodes.append(n)
    return Node("root", [], nodes), version


def _props70(node):
    out = {}
    p = node.first("Properties70") if node else None
    if p:
        for c in p.find("P"):
            out[c.props[0]] = c.props[4:]
    return out


# ---------------------------------------------------------------- geometry
def quad_flips(v0, v1, v2, v3):
    """Blender's loop-triangle rule for quads (BLI math_geom `is_quad_flip_v3_first_third_fast`, float32): a quad is cut along
    its first diagonal v0-v2 — triangles (0,1,2) (0,2,3) — unless v1 and v3 lie on the same side of that diagonal
    (dot(cross(v1-v0, v2-v0), cross(v3-v0, v2-v0)) > 0), in which case it is cut along v1-v3: (0,1,3) (1,2,3)."""
    v0, v1, v2, v3 = (np.asarray(x, np.float32) for x in (v0, v1, v2, v3))
    d02 = v2 - v0
    return bool(np.dot(np.cross(v1 - v0, d02), np.cross(v3 - v0, d02)) > 0)


def triangulate(verts, pvi, quad_rule="blender"):
    """Polygons -> triangles.  quad_rule "blender": the rule above, which reproduces 14,220 of the 14,246 triangles of the
    shipped ornate_mirror.static.json as the same position triples (the plain first-vertex fan: 14,188; the other 26 are
    n-gons, which Blender poly-fills).  quad_rule "shorter": the shorter diagonal — what the committed Semla / 17-Cheese
    fixtures of round 1 were generated with (7,532 of 14,246 on the mirror: same vertices, same surface up to the quads'
    diagonals); kept so that `--fixtures` reproduces those files byte for byte."""
    tris, poly = [], []
    for idx in pvi:
        if idx < 0:
            poly.append(~int(idx))
            n = len(poly)
            if n == 3:
                tris.append(poly)
            elif n == 4:
                a, b, c, d = poly
                if quad_rule == "shorter":
                    first = np.sum((verts[a] - verts[c]) ** 2) <= np.sum((verts[b] - verts[d]) ** 2)
                else:
                    first = not quad_flips(verts[a], verts[b], verts[c], verts[d])
                if first:
                    tris += [[a, b, c], [a, c, d]]
                else:
                    tris += [[a, b, d], [b, c, d]]
            elif n > 4:
                for k in range(1, n - 1):
                    tris.append([poly[0], poly[k], poly[k + 1]])
            poly = []
        else:
            poly.append(int(idx))
    return np.asarray(tris, np.uint32)


def euler_xyz_deg(r):
    rx, ry, rz = np.radians(np.asarray(r, np.float64))
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx  # FBX default rotation order eEulerXYZ


def load_geometry(path, quad_rule="blender"):
    root, version = read_fbx(path)
    gs = _props70(root.first("GlobalSettings"))
    up_axis = int(gs.get("UpAxis", [1])[0])
    unit = float(gs.get("UnitScaleFactor", [1.0])[0])
    objects = root.first("Objects")
    geoms = [g for g in objects.find("Geometry") if g.first("Vertices") is not None]
    models = {m.props[0]: m for m in objects.find("Model")}
    links = {}
    conns = root.first("Connections")
    if conns:
        for c in conns.find("C"):
            if c.props[0] == "OO":
                links[c.props[1]] = c.props[2]
    parts = []
    for g in geoms:
        verts = np.asarray(g.first("Vertices").props[0], np.float64).reshape(-1, 3)
        pvi = np.asarray(g.first("PolygonVertexIndex").props[0], np.int64)
        model = models.get(links.get(g.props[0]))
        p70 = _props70(model)
        T = np.asarray(p70.get("Lcl Translation", [0, 0, 0])[:3], np.float64)
        R = np.asarray(p70.get("Lcl Rotation", [0, 0, 0])[:3], np.float64)
        S = np.asarray(p70.get("Lcl Scaling", [1, 1, 1])[:3], np.float64)
        name = (model.props[1].split("\x00")[0] if model else "mesh")
        M = np.eye(4)
        M[:3, :3] = euler_xyz_deg(R) * S[None, :]
        M[:3, 3] = T
        k = unit * 0.01  # FBX units are centimetres x UnitScaleFactor; Blender imports in metres
        U = np.diag([k, k, k, 1.0])
        if up_axis == 1:  # Y-up file -> Blender Z-up world: rotate +90 deg about X
            A = np.array([[1, 0, 0, 0], [0, 0, -1, 0], [0, 1, 0, 0], [0, 0, 0, 1.0]])
        else:
            A = np.eye(4)
        world = A @ U @ M
        parts.append({"name": name, "positions": verts, "indices": triangulate(verts, pvi, quad_rule).reshape(-1),
                      "transform": world, "up_axis": up_axis, "unit_scale": unit, "fbx_version": version})
    return parts


# ---------------------------------------------------------------- hull stand-ins
def _components(n_verts, tris):
    parent = np.arange(n_verts)

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for a, b, c in tris:
        ra, rb, rc = find(a), find(b), find(c)
        parent[rb] = ra
        parent[rc] = ra
    roots = np.array([find(i) for i in range(n_verts)])
    return roots


def hull_standins(positions, indices, max_hulls=2, max_faces=24):
    from scipy.spatial import ConvexHull
    tris = np.asarray(indices).reshape(-1, 3)
    # weld coincident vertices first (FBX splits vertices far less than the exporter, but be safe)
    key = np.round(positions, 6)
    _, inv = np.unique(key, axis=0, return_inverse=True)
    roots = _components(inv.max() + 1, inv[tris])
    comp_of_vert = roots[inv]
    used = np.unique(comp_of_vert[np.unique(tris)])
    sizes = sorted(((np.count_nonzero(comp_of_vert == c), c) for c in used), reverse=True)[:max_hulls]
    hulls = []
    for _, c in sizes:
        pts = positions[comp_of_vert == c]
        if len(pts) < 4:
            continue
        h = ConvexHull(pts)
        hv = pts[h.vertices]
        # greedy reduction: farthest-point subsets of the hull vertices until the hull has <= max_faces
        k = len(hv)
        while True:
            sub = hv if k >= len(hv) else _farthest_points(hv, k)
            hh = ConvexHull(sub)
            if len(hh.simplices) <= max_faces or k <= 4:
                break
            k = max(4, int(k * 0.8))
        centre = sub.mean(0)
        faces = []
        for s in hh.simplices:  # orient outward
            a, b, c = sub[s]
            if np.dot(np.cross(b - a, c - a), a - centre) < 0:
                s = s[[0, 2, 1]]
            faces.append(s)
        remap = {v: i for i, v in enumerate(np.unique(np.asarray(faces)))}
        hp = sub[list(remap.keys())]
        hi = np.array([[remap[v] for v in f] for f in faces], np.uint32)
        hulls.append((hp.astype(np.float32), hi.reshape(-1)))
    return hulls


def _farthest_points(pts, k):
    chosen = [int(np.argmax(np.linalg.norm(pts - pts.mean(0), axis=1)))]
    d = np.linalg.norm(pts - pts[chosen[0]], axis=1)
    while len(chosen) < k:
        i = int(np.argmax(d))
        chosen.append(i)
        d = np.minimum(d, np.linalg.norm(pts - pts[i], axis=1))
    return pts[chosen]


# ---------------------------------------------------------------- outputs
def to_static_json(parts):
    meshes = []
    for p in parts:
        hulls = hull_standins(p["positions"], p["indices"])
        meshes.append({
            "name": p["name"], "transform": [float(x) for x in np.asarray(p["transform"], np.float32).reshape(16)],
            "mesh": {"positions": [float(x) for x in p["positions"].astype(np.float32).reshape(-1)], "normals": [], "uvs": [],
                     "indices": [int(i) for i in p["indices"]],
                     "submeshes": [{"start": 0, "count": int(len(p["indices"])), "material": "Default"}]},
            "collisionHulls": [{"positions": [float(x) for x in hp.reshape(-1)], "indices": [int(i) for i in hi]} for hp, hi in hulls]})
    return {"version": 1, "meshes": meshes}


def save_fixture(parts, dst):
    assert len(parts) == 1, "fixtures hold single-part assets"
    p = parts[0]
    hulls = hull_standins(p["positions"], p["indices"])
    out = {"name": np.array(p["name"]), "transform_rowmajor": np.asarray(p["transform"], np.float64).astype(np.float32).reshape(16),
           "positions": p["positions"].astype(np.float32), "indices": np.asarray(p["indices"], np.uint32),
           "n_hulls": np.array(len(hulls))}
    if out["indices"].max() <= 65535:
        out["indices"] = out["indices"].astype(np.uint16)
    for i, (hp, hi) in enumerate(hulls):
        out[f"hull{i}_positions"], out[f"hull{i}_indices"] = hp, hi
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes:", len(p["positions"]), "verts,", len(p["indices"]) // 3, "tris,",
          [len(h[1]) // 3 for h in hulls], "hull faces")


def validate_against_mirror():
    """The one asset whose shipped JSON exists: triangle count, local AABB and transform must match."""
    parts = load_geometry(os.path.join(REF, "ornate-mirror/source/ornate_mirror.fbx"))
    ref = json.load(open("/root/reference/Game/ornate_mirror.static.json"))["meshes"][0]
    p = parts[0]
    rp = np.asarray(ref["mesh"]["positions"]).reshape(-1, 3)
    ok_count = len(p["indices"]) // 3 == len(ref["mesh"]["indices"]) // 3
    ok_aabb = np.allclose(p["positions"].min(0), rp.min(0), atol=1e-5) and np.allclose(p["positions"].max(0), rp.max(0), atol=1e-5)
    ok_tr = np.allclose(np.asarray(p["transform"]).reshape(16), np.asarray(ref["transform"]), atol=1e-5)
    print("mirror validation: triangles", len(p["indices"]) // 3, "(shipped", len(ref["mesh"]["indices"]) // 3, ")",
          "count", ok_count, "aabb", ok_aabb, "transform", ok_tr)
    return ok_count and ok_aabb and ok_tr


def main(argv):
    if argv[:1] == ["--fixtures"]:
        rule = argv[1] if len(argv) > 1 else "blender"  # the committed fixtures follow Blender's quad rule (round 2);
        assert rule in ("shorter", "blender")           # `--fixtures shorter` reproduces the round-1 files
        assert validate_against_mirror(), "FBX reader does not reproduce the shipped ornate_mirror asset"
        save_fixture(load_geometry(os.path.join(REF, "semla/source/Semla.fbx"), rule), os.path.join(ROOT, "tests/golden/semla.npz"))
        save_fixture(load_geometry(os.path.join(REF, "17-Cheese.fbx"), rule), os.path.join(ROOT, "tests/golden/cheese.npz"))
        return 0
    if len(argv) != 2:
        print(__doc__)
        return 2
    json.dump(to_static_json(load_geometry(argv[0])), open(argv[1], "w"))
    print("wrote", argv[1])
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
