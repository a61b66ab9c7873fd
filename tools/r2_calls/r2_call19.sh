#!/bin/bash
# Round-2 GPU call 19: per-triangle materials (cq_world_options.triangle_materials) — full parity run + the headline numbers
# of the build with the material row lookup.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -rf --no-header > $O/r2c19_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/r2c19_pytest.log; tail -15 $O/r2c19_pytest.log
run() { local tag=$1; shift; timeout 300 python bench.py "$@" --no-cpu-baseline --no-extras > $O/r2c19_$tag.json 2> $O/r2c19_$tag.err; }
run hulls --mesh hulls --steps 20 --warmup 5
run terrain --mesh terrain --steps 10 --warmup 3
run render --mesh render --steps 5 --warmup 3
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c19_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.1f M/s  %.3f ms/step  e2e %.1f M/s" % (d["value"] / 1e6, d["ms_per_step"], (d.get("e2e") or {}).get("value", 0) / 1e6))
    except Exception as ex:
        print(f, "ERR", ex)
PY
