#!/usr/bin/env python
"""Static code-size breakdown of one kernel by source file / function region (needs -lineinfo, which the Makefile passes).

    python tools/sass_breakdown.py swift-game-engine_b200/csrc/cq_mas.o 'k_move_and_slideILb0ELb0ELb0E'

Extracts the sm_100a cubin from the object (cuobjdump -xelf), disassembles it with `nvdisasm --print-line-info`, and
attributes every SASS instruction of the kernel whose mangled name contains the pattern to the innermost frame of its
inline chain that is not a vector-algebra helper or a CUDA header.  Prints instructions and bytes (16 B each) per source file, per named line range (REGIONS
below: the pieces DESIGN.md §5.1 talks about), and the opcode mix of each region.  No GPU needed.
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

HELPER_LINES = 52  # cq_math.cuh:1-52 = f3 / d3 algebra (dot, cross, normalize ...): charged to the caller
# (file suffix, first line, last line, label): line ranges of the current sources; refresh when the files move
REGIONS = [
    ("cq_math.cuh", 53, 150, "narrow phase: closest_point_on_triangle, segment_segment_dist2"),
    ("cq_math.cuh", 151, 234, "narrow phase: vertical-axis specialisations (vseg_segment_dist2, segment-triangle intersect)"),
    ("cq_math.cuh", 235, 285, "narrow phase: segment_triangle_distance body"),
    ("cq_math.cuh", 286, 305, "ray_triangle"),
    ("cq_math.cuh", 306, 10 ** 9, "capsule-capsule CCD (agents)"),
    ("cq_pool.cuh", 188, 296, "pool: cooperative walk (pool_walk_round)"),
    ("cq_pool.cuh", 297, 332, "pool: job pickup (pool_take_jobs)"),
    ("cq_pool.cuh", 333, 410, "pool: pair state machine (pool_eval, without the distance function)"),
    ("cq_pool.cuh", 411, 470, "pool: commit (pool_commit)"),
    ("cq_pool.cuh", 471, 10 ** 9, "pool: main loop (pool_run)"),
    ("cq_pool.cuh", 1, 187, "pool: posting queries (pool_post_*, roots)"),
    ("cq_world.cuh", 1, 10 ** 9, "cq_world.cuh helpers"),
    ("cq_mas.cu", 1, 10 ** 9, "controller logic (cq_mas.cu)"),
]


def main(obj, pattern):
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
        cubins = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
        text = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cubins[0])], capture_output=True, text=True,
                              check=True).stdout
    in_kernel, chain, fresh = False, [("?", 0)], True
    per_file, per_region, mix = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    total = 0

    def region_of(frames):
        # innermost frame first; helper frames (vector algebra, CUDA headers) are charged to the function that called them
        for f, ln in frames:
            if f.endswith("cq_math.cuh") and ln <= HELPER_LINES:
                continue
            if not f.startswith("cq_"):
                continue
            return f, next((lab for suf, lo, hi, lab in REGIONS if f.endswith(suf) and lo <= ln <= hi), "other: " + f)
        return frames[0][0], "other: " + frames[0][0]

    for line in text.splitlines():
        if line.startswith(".text."):
            in_kernel = pattern in line
            continue
        if line.startswith(".section") or line.startswith("\t.section"):
            in_kernel = False
        if not in_kernel:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            # -gi prints the inline chain innermost first, one frame per line, the outermost last; a chain belongs to the
            # instructions that follow it until the next chain starts
            if fresh:
                chain, fresh = [], False
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        fresh = True
        total += 1
        op = m.group(2).split(".")[0]
        f, label = region_of(chain)
        per_file[f] += 1
        per_region[label] += 1
        mix[label][op] += 1
    print(f"kernel matching {pattern!r}: {total} instructions = {total * 16 / 1024:.1f} KB")
    print("\nby source file:")
    for f, n in per_file.most_common():
        print(f"  {n:6d}  {n * 16 / 1024:6.1f} KB  {f}")
    print("\nby region:")
    for lab, n in per_region.most_common():
        top = ", ".join(f"{op} {c}" for op, c in mix[lab].most_common(8))
        print(f"  {n:6d}  {n * 16 / 1024:6.1f} KB  {lab}\n          {top}")


if __name__ == "__main__":
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    main(sys.argv[1], sys.argv[2])
