"""Regenerates tests/golden/c1_trajectory.npz: config C1 of BASELINE.json (a capsule character driven for 600
fixed steps over the demo's static world) as computed by the CPU oracle in the REFERENCE's own visiting
order.  The oracle itself is pinned by tests/test_oracle_known_answers.py; this file pins the oracle's
behaviour over time (regression) and is what the GPU trajectory is compared with.
    python tests/golden/make_c1_trajectory.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
DST = os.path.join(ROOT, "tests", "golden", "c1_trajectory.npz")

# 4 characters (scenes.C1_STARTS / C1_SPEEDS): the demo player start (0, 7.5, 0) at run speed 12.5
# (CharacterFactory.swift:77-78, InputSystem.swift:121-123) + 3 slower walkers whose 600-frame loops cross the
# mirror hulls, the step box and the ramp


def main():
    sc = importlib.import_module("swift-game-engine_b200.scenes")
    from oracle import oracle as orc
    parts = sc.c1_scene()
    STARTS, SPEEDS = sc.C1_STARTS, sc.C1_SPEEDS
    out = {"starts": STARTS, "speeds": SPEEDS}
    for name, order in (("reference", orc.ORDER_REFERENCE), ("canonical", orc.ORDER_CANONICAL)):
        w = orc.OracleWorld(parts)
        s = orc.init_states(STARTS)
        p = orc.default_params()
        rec = sc.c1_run(lambda st: w.move_and_slide(st, p, order=order), s, 600, SPEEDS)
        for k in rec.dtype.names:
            out[f"{name}_{k}"] = rec[k]
        print(name, "final positions:\n", rec["position"][-1], "\ngrounded frames:", rec["grounded"].sum(0))
    np.savez_compressed(DST, **out)
    print("wrote", DST, os.path.getsize(DST), "bytes")


if __name__ == "__main__":
    main()
