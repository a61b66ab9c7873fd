"""Golden vectors from the reference's OWN Swift code (oracle/swift_ref: the unmodified Game/*.swift sources compiled
against a simd shim) — the route that pins the C++ oracle, and through it the CUDA path, to the reference itself.

No Swift toolchain exists in this repository's build image, so tests/golden/swift_{c1,c2,c3}.npz are absent until
somebody runs oracle/swift_ref/build.sh where swiftc exists; the comparisons below then switch on by themselves
(CPU: the oracle in ORDER_REFERENCE against the goldens; GPU: the library in CQ_ORDER_REFERENCE against the goldens).
What always runs is the plumbing: the exporter's input files parse, an output file in the harness' format (emulated here
with the C++ oracle, in a temporary directory, never committed as a golden) imports into the .npz layout, and the
comparison code accepts it."""
import importlib
import json
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
SWIFT_REF = os.path.join(ROOT, "oracle", "swift_ref")


def _scene(scenes, tag):
    return {"c1": scenes.c1_scene, "c2": lambda: scenes.semla_scene(use_hulls=False), "c3": lambda: scenes.mirror_scene(use_hulls=True)}[tag]()


def _parts_in_order(parts, z):
    """The world with its parts in the entity order the Swift process saw (statics, then dynamics)."""
    by_id = {int(p["entity_id"]): p for p in parts}
    order = [int(t) for t in z["static_order"]] + [int(t) for t in z["dynamic_order"]]
    assert sorted(order) == sorted(by_id)
    return [by_id[t] for t in order]


def _read_inputs(path, cq):
    """Parse an exporter file back into (queries per scenario): proves the format the Swift side reads."""
    blob = open(path, "rb").read()
    magic, version, n_parts = struct.unpack_from("<4sII", blob, 0)
    assert magic == b"CQSW" and version == 1
    at = 12
    for _ in range(n_parts):
        tag, layer, mus, muk, fl, dyn, _z, nv, ni = struct.unpack_from("<IIffBBHII", blob, at)
        at += 28 + 12 * nv + 4 * ni
    (n_sc,) = struct.unpack_from("<I", blob, at)
    at += 4
    out = []
    for _ in range(n_sc):
        (kind,) = struct.unpack_from("<I", blob, at)
        at += 4
        if kind == 1:
            mode, n = struct.unpack_from("<II", blob, at)
            out.append(("casts", mode, np.frombuffer(blob, cq.CAST, n, at + 8).copy()))
            at += 8 + n * cq.CAST.itemsize
        elif kind in (2, 3):
            mh = None
            if kind == 3:
                (mh,) = struct.unpack_from("<I", blob, at)
                at += 4
            (n,) = struct.unpack_from("<I", blob, at)
            out.append(("overlap" if kind == 2 else "overlap_all", mh, np.frombuffer(blob, cq.CAPSULE, n, at + 4).copy()))
            at += 4 + n * cq.CAPSULE.itemsize
        elif kind == 4:
            (n,) = struct.unpack_from("<I", blob, at)
            out.append(("rays", None, np.frombuffer(blob, cq.RAY, n, at + 4).copy()))
            at += 4 + n * cq.RAY.itemsize
        elif kind == 5:
            n, steps, dt, gx, gy, gz, grav, accel, decel, has_intent = struct.unpack_from("<IIf3fIffI", blob, at)
            at += 40
            params = np.frombuffer(blob, cq.PARAMS, 1, at).copy()
            at += cq.PARAMS.itemsize
            pv = np.frombuffer(blob, "<f4", n * 6, at).reshape(n, 6).copy()
            at += n * 24
            intent = None
            if has_intent:
                intent = np.frombuffer(blob, "<f4", steps * n * 3, at).reshape(steps, n, 3).copy()
                at += steps * n * 12
            out.append(("walk", dict(steps=steps, dt=dt, gravity=(gx, gy, gz), accel=accel, decel=decel, intent=intent, params=params), pv))
        else:
            raise AssertionError(kind)
    assert at == len(blob)
    return out


def _intent(states, desired, accel, decel, dt):
    """PhysicsIntentSystem for a character-controller body (Systems.swift:228-233, approachVecD :419-426), Double."""
    v = states["velocity"]
    tgt = np.stack([desired[:, 0].astype(np.float64), np.zeros(len(v)), desired[:, 2].astype(np.float64)], axis=1)
    cur = np.stack([v[:, 0], np.zeros(len(v)), v[:, 2]], axis=1)
    ln_t = np.sqrt((tgt[:, 0] * tgt[:, 0] + tgt[:, 1] * tgt[:, 1]) + tgt[:, 2] * tgt[:, 2])
    ln_c = np.sqrt((cur[:, 0] * cur[:, 0] + cur[:, 1] * cur[:, 1]) + cur[:, 2] * cur[:, 2])
    max_delta = np.where(ln_t >= ln_c, np.float32(accel), np.float32(decel)).astype(np.float64) * np.float64(np.float32(dt))
    delta = tgt - cur
    ln = np.sqrt((delta[:, 0] * delta[:, 0] + delta[:, 1] * delta[:, 1]) + delta[:, 2] * delta[:, 2])
    reach = (ln <= max_delta) | (ln < 0.00001)
    nxt = np.where(reach[:, None], tgt, cur + delta / np.where(ln > 0, ln, 1.0)[:, None] * max_delta[:, None])
    v[:, 0], v[:, 2] = nxt[:, 0], nxt[:, 2]


def _run(world_kind, world, mod, scenarios, order):
    """Every scenario of an input file through `world` (the oracle or the library); returns {name: array} like the .npz."""
    out = {}
    for kind, arg, q in scenarios:
        if kind == "casts":
            out[f"casts_mode{arg}"] = world.capsule_cast(q, arg, order) if world_kind == "oracle" else \
                [world.capsuleCast, world.capsuleCastBlocking, world.capsuleCastGround][arg](q)
        elif kind == "overlap":
            out["overlap"] = world.capsule_overlap(q, order) if world_kind == "oracle" else world.capsuleOverlap(q)
        elif kind == "overlap_all":
            if world_kind == "oracle":  # the C wrapper emits caller order (stable sort by depth): undo nothing, compare sorted
                hits, counts, _ = world.capsule_overlap_all(q, arg, order)
            else:
                hits, counts, _ = world.capsuleOverlapAll(q, arg)
            out[f"overlap_all_{arg}"], out[f"overlap_all_{arg}_counts"] = hits, counts
        elif kind == "rays":
            out["rays"] = world.raycast(q, order) if world_kind == "oracle" else world.raycast(q)
        elif kind == "walk":
            s = mod.init_states(q[:, :3].copy(), q[:, 3:].copy())
            rec = np.zeros((arg["steps"], len(s)), s.dtype)
            for f in range(arg["steps"]):
                if arg["intent"] is not None:
                    _intent(s, arg["intent"][f], arg["accel"], arg["decel"], arg["dt"])
                if world_kind == "oracle":
                    world.move_and_slide(s, arg["params"], arg["dt"], arg["gravity"], 1, order)
                else:
                    world.move_and_slide(s, arg["params"], arg["dt"], arg["gravity"], 1)
                rec[f] = s
            out["walk"] = rec
    return out


def _caller_order(hits, counts):
    depth = np.where(np.arange(hits.shape[1])[None, :] < counts[:, None], hits["depth"], -np.inf)
    return np.take_along_axis(hits, np.argsort(-depth, axis=1, kind="stable"), axis=1)


def _compare(got, z, sorted_overlaps):
    checked = 0
    for name, a in got.items():
        key = name if name in z else ("walk_" + [k for k in z.files if k.startswith("walk_")][0][5:] if name == "walk" else None)
        assert key in z, name
        want = z[key]
        if name.startswith("overlap_all") and not name.endswith("counts"):
            cnt = z[name + "_counts"]
            want = _caller_order(want, cnt)  # the Swift API returns visiting order; its callers sort by depth (SYS:759)
            if not sorted_overlaps:
                a = _caller_order(a, cnt)
        if a.dtype.names and "_pad" in a.dtype.names:
            assert all(np.array_equal(a[f], want[f]) for f in a.dtype.names if f != "_pad"), name
        else:
            assert a.tobytes() == want.tobytes(), name
        checked += 1
    return checked


def test_harness_plumbing_round_trip(cq, orc, scenes, tmp_path):
    """export_inputs.py -> (harness output emulated with the oracle) -> import_goldens.py -> the comparison: the file
    formats the Swift side reads and writes are exactly what the Python side writes and reads."""
    prefix = str(tmp_path / "in")
    subprocess.check_call([sys.executable, os.path.join(SWIFT_REF, "export_inputs.py"), prefix])
    tag = "c3"
    man = json.load(open(f"{prefix}_{tag}.json"))
    scenarios = _read_inputs(f"{prefix}_{tag}.bin", cq)
    assert [s[0] for s in scenarios] == [m["kind"] for m in man["scenarios"]]
    parts = _scene(scenes, tag)
    o = orc.OracleWorld(parts)
    res = _run("oracle", o, orc, scenarios, orc.ORDER_REFERENCE)
    o.close()
    blob = struct.pack("<4sI", b"CQSO", 1)
    ids = [int(p["entity_id"]) for p in parts]
    blob += struct.pack("<I", len(ids)) + np.asarray(ids, "<u4").tobytes() + struct.pack("<I", 0)
    for kind, arg, q in scenarios:  # the order the harness writes in
        if kind == "walk":
            blob += res["walk"].tobytes()
        elif kind == "overlap_all":
            blob += res[f"overlap_all_{arg}"].tobytes() + res[f"overlap_all_{arg}_counts"].astype("<i4").tobytes()
    open(str(tmp_path / f"out_{tag}.bin"), "wb").write(blob)
    sys.path.insert(0, SWIFT_REF)
    imp = importlib.import_module("import_goldens")
    old = imp.ROOT
    imp.ROOT = str(tmp_path)
    os.makedirs(tmp_path / "tests" / "golden")
    try:
        imp.convert(prefix, str(tmp_path / "out"), tag)
    finally:
        imp.ROOT = old
    z = np.load(str(tmp_path / "tests" / "golden" / f"swift_{tag}.npz"))
    assert [int(t) for t in z["static_order"]] == ids and len(z["dynamic_order"]) == 0
    assert _compare(res, z, sorted_overlaps=True) == 3
    for f in ("build.sh", "main.swift", "simd_shim.swift", "Stubs.swift"):
        assert os.path.exists(os.path.join(SWIFT_REF, f))


def _goldens(tag):
    path = os.path.join(GOLDEN, f"swift_{tag}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{os.path.basename(path)} absent: no Swift toolchain has produced it yet (oracle/swift_ref/build.sh)")
    return np.load(path)


@pytest.mark.parametrize("tag", ["c1", "c2", "c3"])
def test_oracle_equals_the_swift_reference(cq, orc, scenes, tag, tmp_path):
    z = _goldens(tag)
    prefix = str(tmp_path / "in")
    subprocess.check_call([sys.executable, os.path.join(SWIFT_REF, "export_inputs.py"), prefix])
    o = orc.OracleWorld(_parts_in_order(_scene(scenes, tag), z))
    got = _run("oracle", o, orc, _read_inputs(f"{prefix}_{tag}.bin", cq), orc.ORDER_REFERENCE)
    o.close()
    assert _compare(got, z, sorted_overlaps=True) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["c1", "c2", "c3"])
def test_library_equals_the_swift_reference(cq, scenes, tag, tmp_path):
    z = _goldens(tag)
    prefix = str(tmp_path / "in")
    subprocess.check_call([sys.executable, os.path.join(SWIFT_REF, "export_inputs.py"), prefix])
    g = cq.CollisionQuery(_parts_in_order(_scene(scenes, tag), z), order=cq.ORDER_REFERENCE)
    got = _run("library", g, cq, _read_inputs(f"{prefix}_{tag}.bin", cq), None)
    g.close()
    assert _compare(got, z, sorted_overlaps=False) > 0
