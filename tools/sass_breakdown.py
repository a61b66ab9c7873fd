#!/usr/bin/env python
"""Static code-size breakdown of one kernel by source file / function region (needs -lineinfo, which the Makefile passes).

    python tools/sass_breakdown.py swift-game-engine_b200/csrc/cq_mas.o 'k_move_and_slideILb0ELb0ELb0E' [ncu_source.csv]

Extracts the sm_100a cubin from the object (cuobjdump -xelf), disassembles it with `nvdisasm --print-line-info`, and
attributes every SASS instruction of the kernel whose mangled name contains the pattern to the innermost frame of its
inline chain that is not a vector-algebra helper or a CUDA header.  With a third argument — the CSV of
`ncu -i report.ncu-rep --page source --csv` for the same kernel of the same binary — the per-instruction samples and
executed counts of the report are joined on and summed per region (the join checks every opcode).  Prints instructions and bytes (16 B each) per source file, per named line range (REGIONS
below: the pieces DESIGN.md §5.1 talks about), and the opcode mix of each region.  No GPU needed.
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

# the sources the object was built from (CQ_CSRC overrides: an object of an older commit needs that commit's line numbers)
CSRC = os.environ.get("CQ_CSRC") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                 "swift-game-engine_b200", "csrc")


def _line_of(file, needle):
    """1-based line of the first line of csrc/<file> that contains `needle` (a function definition), or None."""
    try:
        for i, line in enumerate(open(os.path.join(CSRC, file)), 1):
            if needle in line:
                return i
    except OSError:
        pass
    return None


def _regions():
    """(file suffix, first line, last line, label): the pieces DESIGN.md §5.1 talks about, located by the definitions
    that open them so that the table follows the sources as they move."""
    def cuts(file, marks):  # marks: (needle, label) in file order; a region runs up to the next mark that was found
        found = [(ln, lab) for ln, lab in ((_line_of(file, nd), lab) for nd, lab in marks) if ln is not None]
        found.sort()
        out = []
        for k, (ln, lab) in enumerate(found):
            hi = found[k + 1][0] - 1 if k + 1 < len(found) else 10 ** 9
            out.append((file, ln, hi, lab))
        if found:
            out.append((file, 1, found[0][0] - 1, "(head of %s)" % file))
        return out
    math = cuts("cq_math.cuh", [
        ("float closest_point_on_triangle(", "narrow phase: closest_point_on_triangle, segment_segment_dist2"),
        ("float vseg_segment_dist2(", "narrow phase: vertical-axis specialisations (vseg_segment_dist2, segment-triangle intersect)"),
        ("float segment_triangle_distance(", "narrow phase: segment_triangle_distance body"),
        ("bool ray_triangle(", "ray_triangle"),
        ("bool clamp_interval(", "capsule-capsule CCD (agents)")])
    pool = cuts("cq_pool.cuh", [
        ("void pool_push_roots(", "pool: posting queries (pool_post_*, roots)"),
        ("bool sweep_reach(", "pool: sweeps holding a hit: reach box / cannot-matter tests (called from the walk and the pickup)"),
        ("void pool_walk_round(", "pool: cooperative walk (pool_walk_round)"),
        ("void pool_take_jobs(", "pool: job pickup (pool_take_jobs)"),
        ("void pool_eval(", "pool: pair state machine (pool_eval, without the distance function)"),
        ("struct OverlapTop2", "pool: commit (pool_commit)"),
        ("void pool_run(", "pool: main loop (pool_run)")])
    rest = [("cq_world.cuh", 1, 10 ** 9, "cq_world.cuh helpers"), ("cq_mas.cu", 1, 10 ** 9, "controller logic (cq_mas.cu)")]
    return math + pool + rest


REGIONS = _regions()
# cq_math.cuh up to its first function region = f3 / d3 algebra (dot, cross, normalize ...): charged to the caller
HELPER_LINES = (_line_of("cq_math.cuh", "float closest_point_on_triangle(") or 53) - 1


def load_ncu(path):
    """Rows of `ncu -i X.ncu-rep --page source --csv` (SASS view) of ONE kernel: per instruction, in address order."""
    import csv
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        op = r[col["Source"]].split()
        op = [t for t in op if not t.startswith("@")]
        out.append({"op": op[0].split(".")[0] if op else "?", "samples": int(r[col["# Samples"]]),
                    "inst": int(r[col["Instructions Executed"]]), "thread_inst": int(r[col["Thread Instructions Executed"]]),
                    "no_inst": int(r[col["stall_no_inst"]]), "long_sb": int(r[col["stall_long_sb"]]),
                    "wait": int(r[col["stall_wait"]]), "branch": int(r[col["stall_branch_resolving"]])})
    return out


def main(obj, pattern, ncu_csv=None):
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
        cubins = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
        text = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cubins[0])], capture_output=True, text=True,
                              check=True).stdout
    ncu = load_ncu(ncu_csv) if ncu_csv else None
    dyn = collections.defaultdict(lambda: collections.Counter())
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
        if ncu is not None:
            if total > len(ncu) or ncu[total - 1]["op"] != op:
                sys.exit(f"the profile is not of this binary: instruction {total - 1} is {op} here, "
                         f"{ncu[total - 1]['op'] if total <= len(ncu) else 'missing'} in the report")
            for k, v in ncu[total - 1].items():
                if k != "op":
                    dyn[label][k] += v
    print(f"kernel matching {pattern!r}: {total} instructions = {total * 16 / 1024:.1f} KB")
    print("\nby source file:")
    for f, n in per_file.most_common():
        print(f"  {n:6d}  {n * 16 / 1024:6.1f} KB  {f}")
    print("\nby region:")
    for lab, n in per_region.most_common():
        top = ", ".join(f"{op} {c}" for op, c in mix[lab].most_common(8))
        print(f"  {n:6d}  {n * 16 / 1024:6.1f} KB  {lab}\n          {top}")


    if ncu is not None:
        if len(ncu) != total:
            sys.exit(f"the profile is not of this binary: {len(ncu)} instructions in the report, {total} here")
        tot = collections.Counter()
        for d in dyn.values():
            tot.update(d)
        print("\ndynamic picture from the ncu report (same binary, joined instruction by instruction):")
        print("  region: share of warp-stall samples | share of executed warp instructions | active lanes per instruction | "
              "top stall reasons among its samples")
        for lab, d in sorted(dyn.items(), key=lambda kv: -kv[1]["samples"]):
            lanes = d["thread_inst"] / d["inst"] if d["inst"] else 0.0
            s_ = max(d["samples"], 1)
            print(f"  {100 * d['samples'] / tot['samples']:5.1f}%  {100 * d['inst'] / tot['inst']:5.1f}%  {lanes:5.1f}  {lab}\n"
                  f"          no_instruction {100 * d['no_inst'] / s_:.0f}%, long_scoreboard {100 * d['long_sb'] / s_:.0f}%, "
                  f"wait {100 * d['wait'] / s_:.0f}%, branch_resolving {100 * d['branch'] / s_:.0f}%")


if __name__ == "__main__":
    if len(sys.argv) not in (3, 4):
        sys.exit(__doc__)
    main(*sys.argv[1:])
