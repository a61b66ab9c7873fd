#!/bin/bash
# Round-2 GPU call 4: where did the hulls step lose 10%?  Same-box A/B of: the round-1 code (+SINGLE_POST) as control,
# the current default, and builds without the tie flag / without the stack guards / with the look-ahead prune; the
# phase-voted ray kernel against the stepwise ones in both orders; parity of everything the ray kernel touches.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
D=swift-game-engine_b200/csrc
timeout 900 python -m pytest tests -m gpu -q -rf --no-header -k "ray or c5 or terrain or refit or streams or empty or group" > $O/r2c4_pytest_rays.log 2>&1
echo "pytest rays rc=$?" | tee -a $O/r2c4_pytest_rays.log
tail -8 $O/r2c4_pytest_rays.log
run() { # tag, lib, args...
  local tag=$1 lib=$2; shift 2
  CQ_LIB=$D/$lib timeout 300 python bench.py "$@" --no-cpu-baseline --no-extras > $O/r2c4_ab_${tag}.json 2> $O/r2c4_ab_${tag}.err
}
for rep in 1 2; do
  (cd .ab_old && timeout 120 python bench.py --mesh hulls --steps 20 --warmup 5 --no-cpu-baseline > ../$O/r2c4_ab_hulls_old_$rep.json 2> ../$O/r2c4_ab_hulls_old_$rep.err)
  for L in libcq libcq_notie libcq_noguard libcq_la; do
    run hulls_${L}_$rep $L.so --mesh hulls --steps 20 --warmup 5
  done
done
(cd .ab_old && timeout 200 python bench.py --mesh terrain --steps 10 --warmup 3 --no-cpu-baseline > ../$O/r2c4_ab_terrain_old.json 2> ../$O/r2c4_ab_terrain_old.err)
(cd .ab_old && timeout 200 python bench.py --mesh render --steps 5 --warmup 3 --no-cpu-baseline > ../$O/r2c4_ab_render_old.json 2> ../$O/r2c4_ab_render_old.err)
(cd .ab_old && timeout 200 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > ../$O/r2c4_ab_c2_old.json 2> ../$O/r2c4_ab_c2_old.err)
for L in libcq libcq_notie libcq_noguard libcq_la; do
  run terrain_$L $L.so --mesh terrain --steps 10 --warmup 3
done
for L in libcq libcq_la; do
  run render_$L $L.so --mesh render --steps 5 --warmup 3
  run c2_$L $L.so --only c2 --steps 3 --warmup 3
done
run c5_phased libcq.so --only c5 --steps 5 --warmup 3
run c5canon_phased libcq.so --only c5 --steps 5 --warmup 3 --order canonical
CQ_RAY_STEPWISE=1 run c5_stepwise libcq.so --only c5 --steps 5 --warmup 3
CQ_RAY_STEPWISE=1 run c5canon_stepwise libcq.so --only c5 --steps 5 --warmup 3 --order canonical
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c4_ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        pq = d["roofline"].get("per_query", {})
        print(f, "%.1f M/s  %.3f ms/step  e2e %.1f M/s (%.2f ms)  evals/q %s" % (d["value"] / 1e6, d["ms_per_step"], e.get("value", 0) / 1e6,
              e.get("ms_per_step", 0), pq.get("distance_evals")))
    except Exception as ex:
        print(f, "ERR", ex)
PY
# ncu: the phase-voted ray kernel in reference order (the default), one launch
C5="python bench.py --only c5 --steps 1 --warmup 3 --no-cpu-baseline"
$C5 > $O/r2c4_c5_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_raycast -s 3 -c 1 -f -o $O/r2c4_c5_ray_phased $C5 > $O/r2c4_c5_ncu.log 2>&1
echo "ncu c5 rc=$?"
