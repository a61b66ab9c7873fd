#!/bin/bash
# Round-2 GPU call 20: front-end thresholds after the slim pair state — idle lanes that open the front end of the
# move-and-slide kernel (16 shipped; 20, 24) and how far the walk fills the pair ring (96 shipped; 64, 160).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
D=swift-game-engine_b200/csrc
run() { local tag=$1 lib=$2; shift 2; CQ_LIB=$D/$lib.so timeout 300 python bench.py "$@" --no-cpu-baseline --no-extras > $O/r2c20_ab_${tag}_$lib.json 2> $O/r2c20_ab_${tag}_$lib.err; }
for L in libcq libcq_fe20 libcq_fe24 libcq_wf64 libcq_wf160; do
  run hulls $L --mesh hulls --steps 20 --warmup 5
  run terrain $L --mesh terrain --steps 10 --warmup 3
  run render $L --mesh render --steps 5 --warmup 3
done
for L in libcq libcq_wf64 libcq_wf160; do
  run c4 $L --only c4 --steps 5 --warmup 3
  run c2 $L --only c2 --steps 3 --warmup 3
done
run hulls2 libcq --mesh hulls --steps 20 --warmup 5
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c20_ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.1f M/s  %.3f ms/step  e2e %.1f M/s" % (d["value"] / 1e6, d["ms_per_step"], (d.get("e2e") or {}).get("value", 0) / 1e6))
    except Exception as ex:
        print(f, "ERR", ex)
PY
