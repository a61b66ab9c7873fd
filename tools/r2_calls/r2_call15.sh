#!/bin/bash
# Round-2 GPU call 15: ray traversal stack in shared memory (libcq) against the local-memory stack (libcq_prev); overlap drop
# without the retry loop on terrain / render; parity.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
D=swift-game-engine_b200/csrc
timeout 1800 python -m pytest tests -m gpu -q -rf --no-header > $O/r2c15_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/r2c15_pytest.log; tail -12 $O/r2c15_pytest.log
run() { local tag=$1 lib=$2; shift 2; CQ_LIB=$D/$lib.so timeout 300 python bench.py "$@" --no-cpu-baseline --no-extras > $O/r2c15_ab_${tag}_$lib.json 2> $O/r2c15_ab_${tag}_$lib.err; }
for L in libcq libcq_prev; do
  run c5ref $L --only c5 --steps 5 --warmup 3
  run c5can $L --only c5 --order canonical --steps 5 --warmup 3
  run terrain $L --mesh terrain --steps 10 --warmup 3
  run render $L --mesh render --steps 5 --warmup 3
done
run c5refb libcq --only c5 --steps 5 --warmup 3
run c5refb libcq_prev --only c5 --steps 5 --warmup 3
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c15_ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        pq = d["roofline"].get("per_query", {})
        print(f, "%.1f M/s  %.3f ms/step  kernel %.3f ms  e2e %.1f M/s  evals/q %s cands %s" % (d["value"] / 1e6, d["ms_per_step"], d["roofline"].get("kernel_ms", 0), e.get("value", 0) / 1e6,
              pq.get("distance_evals"), pq.get("candidates")))
    except Exception as ex:
        print(f, "ERR", ex)
PY
