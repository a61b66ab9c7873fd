#!/bin/bash
# Round-2 GPU call 6: parity of the default build and of the plane-separation variant; same-box A/B of the remaining
# suspects for the 3-5% the step lost against round 1 (tie flag, record stride, inlined overlap commit) and of the variant.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
D=swift-game-engine_b200/csrc
timeout 1500 python -m pytest tests -m gpu -q -rf --no-header > $O/r2c6_pytest.log 2>&1
echo "pytest default rc=$?" | tee -a $O/r2c6_pytest.log; tail -4 $O/r2c6_pytest.log
CQ_LIB=$D/libcq_pc.so timeout 1500 python -m pytest tests -m gpu -q -rf --no-header > $O/r2c6_pytest_pc.log 2>&1
echo "pytest plane-cull rc=$?" | tee -a $O/r2c6_pytest_pc.log; tail -6 $O/r2c6_pytest_pc.log
run() { local tag=$1 lib=$2; shift 2; CQ_LIB=$D/$lib.so timeout 300 python bench.py "$@" --no-cpu-baseline --no-extras > $O/r2c6_ab_${tag}_$lib.json 2> $O/r2c6_ab_${tag}_$lib.err; }
old() { local tag=$1; shift; (cd .ab_old && timeout 300 python bench.py "$@" --no-cpu-baseline > ../$O/r2c6_ab_${tag}_old.json 2> ../$O/r2c6_ab_${tag}_old.err); }
old hulls --mesh hulls --steps 20 --warmup 5
for L in libcq libcq_notie libcq_pad39 libcq_inl libcq_pc; do run hulls $L --mesh hulls --steps 20 --warmup 5; done
old hulls2 --mesh hulls --steps 20 --warmup 5
old terrain --mesh terrain --steps 10 --warmup 3
for L in libcq libcq_notie libcq_pad39 libcq_inl libcq_pc; do run terrain $L --mesh terrain --steps 10 --warmup 3; done
for L in libcq libcq_inl libcq_pc; do run render $L --mesh render --steps 5 --warmup 3; done
for L in libcq libcq_pad39 libcq_pc; do run c4 $L --only c4 --steps 5 --warmup 3; run c2 $L --only c2 --steps 3 --warmup 3; done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c6_ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        pq = d["roofline"].get("per_query", {})
        print(f, "%.1f M/s  %.3f ms/step  e2e %.1f M/s (%.2f ms)  evals/q %s" % (d["value"] / 1e6, d["ms_per_step"], e.get("value", 0) / 1e6,
              e.get("ms_per_step", 0), pq.get("distance_evals")))
    except Exception as ex:
        print(f, "ERR", ex)
PY
