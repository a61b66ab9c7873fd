#!/bin/bash
# Round-2 GPU call 1: merged resident crowd + compile-time variants A/B + the ncu captures round 1 left out
# (k_capsule_cast on C2, k_raycast on C5).  Everything lands in gpurun_out/r2c1_*.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
D=swift-game-engine_b200/csrc
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r2c1_gpu.txt 2>&1
lscpu | head -20 >> $O/r2c1_gpu.txt

timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c1_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2c1_pytest.log
tail -3 $O/r2c1_pytest.log

timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2c1_bench_default.json 2> $O/r2c1_bench_default.err
echo "bench default rc=$?"

run() { # tag, lib, args...
  local tag=$1 lib=$2; shift 2
  CQ_LIB=$D/$lib timeout 300 python bench.py "$@" --no-cpu-baseline > $O/r2c1_ab_${tag}.json 2> $O/r2c1_ab_${tag}.err
}
for L in libcq libcq_sp libcq_ep libcq_spep; do
  run hulls_$L $L.so --mesh hulls --steps 20 --warmup 5
  run terrain_$L $L.so --mesh terrain --steps 10 --warmup 3
done
for L in libcq libcq_ep libcq_ub; do
  run c4_$L $L.so --workload c4 --steps 5 --warmup 3
done
run render_libcq libcq.so --mesh render --steps 5 --warmup 3
run render_libcq_ep libcq_ep.so --mesh render --steps 5 --warmup 3

timeout 600 python bench.py --workload c2 --steps 3 --warmup 3 > $O/r2c1_c2.json 2> $O/r2c1_c2.err
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 > $O/r2c1_c5.json 2> $O/r2c1_c5.err

python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c1_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        rc = (e.get("resident_crowd") or {})
        print(f, "%.1f M/s  %.3f ms/step  e2e %.1f M/s  crowd-e2e %s" % (d["value"] / 1e6, d["ms_per_step"], e.get("value", 0) / 1e6,
              ("%.1f M/s %.2f ms" % (rc["value"] / 1e6, rc["ms_per_step"])) if rc else "-"))
    except Exception as ex:
        print(f, "ERR", ex)
PY

# ncu: one full capture each of the C2 cast kernel and the C5 raycast kernel (plain run first, same command)
C2="python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline"
$C2 > $O/r2c1_c2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_capsule_cast -s 3 -c 1 -f -o $O/r2c1_c2_cast $C2 > $O/r2c1_c2_ncu.log 2>&1
echo "ncu c2 rc=$?"
C5="python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline"
$C5 > $O/r2c1_c5_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_raycast -s 3 -c 1 -f -o $O/r2c1_c5_ray $C5 > $O/r2c1_c5_ncu.log 2>&1
echo "ncu c5 rc=$?"
ls -la $O | tail -30
