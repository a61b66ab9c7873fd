#!/usr/bin/env python
"""Rebuild profiles/r2_traffic.json from the committed ncu summaries (profiles/r2_*_summary.txt): DRAM bytes read + written
by ONE launch of each kernel on its BASELINE workload, divided by the units that launch processed.  bench.py scales the
per-unit figure into `roofline.traffic`."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROWS = [  # kernel key, config key, summary file, units, workload
    ("k_move_and_slide", "hulls", "r2_mas_hulls_summary.txt", 1 << 20, "C3 hulls, 1,048,576 characters"),
    ("k_move_and_slide", "terrain", "r2_mas_terrain_summary.txt", 1 << 20, "target scene (10 M-triangle terrain), 1,048,576 characters"),
    ("k_move_and_slide", "render", "r2_mas_render_summary.txt", 1 << 20, "C3 render mesh, 1,048,576 characters"),
    ("k_capsule_cast", "c4", "r2_cast_c4_summary.txt", 1 << 23, "C4, 8,388,608 blocking sweeps over the terrain"),
    ("k_capsule_cast", "c2", "r2_cast_c2_summary.txt", 1 << 16, "C2, 65,536 sweeps vs Semla"),
    ("k_raycast", "c5", "r2_ray_c5_ref_summary.txt", 1 << 24, "C5, 16,777,216 rays, reference order (k_raycast_phased<.,REF>)"),
    ("k_raycast", "c5_canonical", "r2_ray_c5_canon_summary.txt", 1 << 24, "C5, 16,777,216 rays, canonical order"),
    ("k_capsule_overlap_pool", "overlap_all", "r2_overlap_all_summary.txt", 1 << 20, "1,048,576 capsuleOverlapAll(8) vs the mirror render mesh"),
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def metric(text, name):
    m = re.search(r"^%s (\w+) ([0-9.eE+-]+)" % re.escape(name), text, re.M)
    return float(m.group(2)) * SCALE[m.group(1)]


def main():
    out = {}
    for kernel, cfg, fn, units, workload in ROWS:
        text = open(os.path.join(ROOT, "profiles", fn)).read()
        rd, wr = metric(text, "dram__bytes_read.sum"), metric(text, "dram__bytes_write.sum")
        out.setdefault(kernel, {})[cfg] = {
            "workload": workload, "units": units, "dram_bytes_read": rd, "dram_bytes_write": wr,
            "dram_bytes_per_unit": (rd + wr) / units,
            "source": "ncu --set full --clock-control none, one launch: profiles/%s (tools/r2_evidence.sh, final call of round 2)" % fn}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
    for k, v in out.items():
        for c, x in v.items():
            print("%-24s %-14s %10.1f B/unit" % (k, c, x["dram_bytes_per_unit"]))


if __name__ == "__main__":
    main()
