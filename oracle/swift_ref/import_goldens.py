#!/usr/bin/env python
"""import_goldens.py — outputs of the Swift harness -> tests/golden/swift_{c1,c2,c3}.npz:

    python oracle/swift_ref/import_goldens.py /tmp/cq_inputs /tmp/cq_outputs
        reads  /tmp/cq_inputs_{c1,c2,c3}.json (manifests) and /tmp/cq_outputs_{c1,c2,c3}.bin (harness results)

Each .npz holds `static_order` / `dynamic_order` (the entity order the reference's Dictionary produced in that process:
tests build their worlds with the parts in this order so that triangle numbering agrees) and one array per scenario in the
record dtypes of include/cq.h.  TEST INFRASTRUCTURE (oracle/)."""
import importlib
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
cq = importlib.import_module("swift-game-engine_b200")


def convert(inputs_prefix, outputs_prefix, tag):
    man = json.load(open(f"{inputs_prefix}_{tag}.json"))
    blob = open(f"{outputs_prefix}_{tag}.bin", "rb").read()
    magic, version = struct.unpack_from("<4sI", blob, 0)
    assert magic == b"CQSO" and version == 1
    at = 8
    out = {}
    for name in ("static_order", "dynamic_order"):
        (k,) = struct.unpack_from("<I", blob, at)
        out[name] = np.frombuffer(blob, "<u4", k, at + 4).copy()
        at += 4 + 4 * k

    def take(dtype, count):
        nonlocal at
        a = np.frombuffer(blob, dtype, count, at).copy()
        at += a.nbytes
        return a

    for i, s in enumerate(man["scenarios"]):
        if s["kind"] == "casts":
            out[f"casts_mode{s['mode']}"] = take(cq.CAST_HIT, s["n"])
        elif s["kind"] == "overlap":
            out["overlap"] = take(cq.OVERLAP_HIT, s["n"])
        elif s["kind"] == "overlap_all":
            mh = max(1, s["max_hits"])
            out[f"overlap_all_{mh}"] = take(cq.OVERLAP_HIT, s["n"] * mh).reshape(s["n"], mh)
            out[f"overlap_all_{mh}_counts"] = take("<i4", s["n"])
        elif s["kind"] == "rays":
            out["rays"] = take(cq.RAY_HIT, s["n"])
        elif s["kind"] == "walk":
            out[f"walk_{s['name']}"] = take(cq.STATE, s["steps"] * s["n"]).reshape(s["steps"], s["n"])
    assert at == len(blob), "trailing bytes in the harness output"
    path = os.path.join(ROOT, "tests", "golden", f"swift_{tag}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    for tag in ("c1", "c2", "c3"):
        if os.path.exists(f"{sys.argv[2]}_{tag}.bin"):
            convert(sys.argv[1], sys.argv[2], tag)
