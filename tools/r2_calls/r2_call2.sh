#!/bin/bash
# Round-2 GPU call 2: reference-order parity (default build), look-ahead prune variant (parity + A/B).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
D=swift-game-engine_b200/csrc
timeout 1200 python -m pytest tests -m gpu -q -rf --no-header > $O/r2c2_pytest.log 2>&1
echo "pytest default rc=$?" | tee -a $O/r2c2_pytest.log
tail -25 $O/r2c2_pytest.log
CQ_LIB=$D/libcq_la.so timeout 1200 python -m pytest tests -m gpu -q -rf --no-header > $O/r2c2_pytest_la.log 2>&1
echo "pytest lookahead rc=$?" | tee -a $O/r2c2_pytest_la.log
tail -15 $O/r2c2_pytest_la.log
run() { # tag, lib, args...
  local tag=$1 lib=$2; shift 2
  CQ_LIB=$D/$lib timeout 300 python bench.py "$@" --no-cpu-baseline > $O/r2c2_ab_${tag}.json 2> $O/r2c2_ab_${tag}.err
}
for L in libcq libcq_la; do
  run hulls_$L $L.so --mesh hulls --steps 20 --warmup 5
  run terrain_$L $L.so --mesh terrain --steps 10 --warmup 3
  run render_$L $L.so --mesh render --steps 5 --warmup 3
  run c2_$L $L.so --workload c2 --steps 3 --warmup 3
  run c4_$L $L.so --workload c4 --steps 5 --warmup 3
  run c5_$L $L.so --workload c5 --steps 5 --warmup 3
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c2_ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        pq = d["roofline"].get("per_query", {})
        print(f, "%.1f M/s  %.3f ms/step  e2e %.1f M/s  evals/q %s" % (d["value"] / 1e6, d["ms_per_step"], e.get("value", 0) / 1e6,
              pq.get("distance_evals")))
    except Exception as ex:
        print(f, "ERR", ex)
PY
