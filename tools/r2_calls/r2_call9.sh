#!/bin/bash
# Round-2 GPU call 9: owner-grouped commit against the lane-at-a-time commit (same box), full parity incl. the new
# full-size C2 / C5 tests.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
D=swift-game-engine_b200/csrc
timeout 1800 python -m pytest tests -m gpu -q -rf --no-header > $O/r2c9_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/r2c9_pytest.log; tail -12 $O/r2c9_pytest.log
run() { local tag=$1 lib=$2; shift 2; CQ_LIB=$D/$lib.so timeout 300 python bench.py "$@" --no-cpu-baseline --no-extras > $O/r2c9_ab_${tag}_$lib.json 2> $O/r2c9_ab_${tag}_$lib.err; }
for L in libcq libcq_serial; do
  run hulls $L --mesh hulls --steps 20 --warmup 5
  run terrain $L --mesh terrain --steps 10 --warmup 3
  run render $L --mesh render --steps 5 --warmup 3
  run c2 $L --only c2 --steps 3 --warmup 3
  run c4 $L --only c4 --steps 5 --warmup 3
  CQ_LIB=$D/$L.so timeout 200 python tools/profile_extra.py overlap > $O/r2c9_overlap_$L.txt 2>&1
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c9_ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        print(f, "%.1f M/s  %.3f ms/step  e2e %.1f M/s (%.2f ms)" % (d["value"] / 1e6, d["ms_per_step"], e.get("value", 0) / 1e6, e.get("ms_per_step", 0)))
    except Exception as ex:
        print(f, "ERR", ex)
PY
cat $O/r2c9_overlap_*.txt
