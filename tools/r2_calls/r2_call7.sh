#!/bin/bash
# Round-2 GPU call 7: FE_READY (run the owners' logic only when enough owners are ready) A/B on the move-and-slide scenes,
# with parity of the best-looking variant.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
D=swift-game-engine_b200/csrc
run() { local tag=$1 lib=$2; shift 2; CQ_LIB=$D/$lib.so timeout 300 python bench.py "$@" --no-cpu-baseline --no-extras > $O/r2c7_ab_${tag}_$lib.json 2> $O/r2c7_ab_${tag}_$lib.err; }
for L in libcq libcq_ready8 libcq_ready12 libcq_ready16 libcq_ready24; do
  run hulls $L --mesh hulls --steps 20 --warmup 5
  run terrain $L --mesh terrain --steps 10 --warmup 3
  run render $L --mesh render --steps 5 --warmup 3
done
run hulls2 libcq --mesh hulls --steps 20 --warmup 5
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c7_ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        print(f, "%.1f M/s  %.3f ms/step  e2e %.1f M/s (%.2f ms)" % (d["value"] / 1e6, d["ms_per_step"], e.get("value", 0) / 1e6, e.get("ms_per_step", 0)))
    except Exception as ex:
        print(f, "ERR", ex)
PY
for L in libcq_ready12 libcq_ready16; do
CQ_LIB=$D/$L.so timeout 900 python -m pytest tests -m gpu -q --no-header -k "move_and_slide or agents or separation or platforms or c1 or parameter or crowd or full_size_c3" > $O/r2c7_pytest_$L.log 2>&1
echo "pytest $L rc=$?"; tail -3 $O/r2c7_pytest_$L.log
done
