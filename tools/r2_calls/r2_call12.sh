#!/bin/bash
# Round-2 GPU call 12: slim pair state (libcq) against the build before it (libcq_fat); walk-level bestT culling (both builds);
# agent-separation rounds as a device-side WHILE graph against the host-driven schedule (CQ_SEP_HOST_ROUNDS=1); parity.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
D=swift-game-engine_b200/csrc
timeout 1800 python -m pytest tests -m gpu -q -rf --no-header > $O/r2c12_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/r2c12_pytest.log; tail -12 $O/r2c12_pytest.log
run() { local tag=$1 lib=$2; shift 2; CQ_LIB=$D/$lib.so timeout 300 python bench.py "$@" --no-cpu-baseline --no-extras > $O/r2c12_ab_${tag}_$lib.json 2> $O/r2c12_ab_${tag}_$lib.err; }
for L in libcq libcq_fat; do
  run hulls $L --mesh hulls --steps 20 --warmup 5
  run terrain $L --mesh terrain --steps 10 --warmup 3
  run render $L --mesh render --steps 5 --warmup 3
  run c2 $L --only c2 --steps 3 --warmup 3
  run c4 $L --only c4 --steps 5 --warmup 3
  run c5 $L --only c5 --steps 5 --warmup 3
done
run hulls2 libcq --mesh hulls --steps 20 --warmup 5
run hulls2 libcq_fat --mesh hulls --steps 20 --warmup 5
run sepdev libcq --mesh terrain --agents 0.1 --separation --steps 5 --warmup 3
CQ_SEP_HOST_ROUNDS=1 run sephost libcq --mesh terrain --agents 0.1 --separation --steps 5 --warmup 3
run sepdev30 libcq --mesh terrain --agents 0.3 --separation --steps 5 --warmup 3
CQ_SEP_HOST_ROUNDS=1 run sephost30 libcq --mesh terrain --agents 0.3 --separation --steps 5 --warmup 3
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c12_ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        pq = d["roofline"].get("per_query", {})
        print(f, "%.1f M/s  %.3f ms/step  e2e %.1f M/s (%.2f ms)  evals/q %s cands %s" % (d["value"] / 1e6, d["ms_per_step"], e.get("value", 0) / 1e6,
              e.get("ms_per_step", 0), pq.get("distance_evals"), pq.get("candidates")))
    except Exception as ex:
        print(f, "ERR", ex)
PY
