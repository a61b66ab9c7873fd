#!/bin/bash
# Round-2 GPU call 13: overlap drops at pickup in the staged move-and-slide kernel (libcq) against the previous commit
# (libcq_prev); counting modes (reference stats / path counters); parity.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
D=swift-game-engine_b200/csrc
timeout 1800 python -m pytest tests -m gpu -q -rf --no-header > $O/r2c13_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/r2c13_pytest.log; tail -12 $O/r2c13_pytest.log
run() { local tag=$1 lib=$2; shift 2; CQ_LIB=$D/$lib.so timeout 300 python bench.py "$@" --no-cpu-baseline --no-extras > $O/r2c13_ab_${tag}_$lib.json 2> $O/r2c13_ab_${tag}_$lib.err; }
for L in libcq libcq_prev; do
  run render $L --mesh render --steps 5 --warmup 3
  run terrain $L --mesh terrain --steps 10 --warmup 3
  run hulls $L --mesh hulls --steps 20 --warmup 5
done
run c2 libcq --only c2 --steps 3 --warmup 3
run c4 libcq --only c4 --steps 5 --warmup 3
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c13_ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        pq = d["roofline"].get("per_query", {})
        print(f, "%.1f M/s  %.3f ms/step  e2e %.1f M/s (%.2f ms)  evals/q %s cands %s" % (d["value"] / 1e6, d["ms_per_step"], e.get("value", 0) / 1e6,
              e.get("ms_per_step", 0), pq.get("distance_evals"), pq.get("candidates")))
    except Exception as ex:
        print(f, "ERR", ex)
PY
