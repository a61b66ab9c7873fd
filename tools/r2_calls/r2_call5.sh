#!/bin/bash
# Round-2 GPU call 5: rank carried in tv2.w + warp-uniform stack guard + look-ahead in the query kernels only.
# Full GPU parity, then the same-box A/B against the round-1 code (.ab_old = round 1 + SINGLE_POST), then the bench line.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rf --no-header > $O/r2c5_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/r2c5_pytest.log
tail -8 $O/r2c5_pytest.log
run() { local tag=$1; shift; timeout 300 python bench.py "$@" --no-cpu-baseline --no-extras > $O/r2c5_ab_${tag}.json 2> $O/r2c5_ab_${tag}.err; }
old() { local tag=$1; shift; (cd .ab_old && timeout 300 python bench.py "$@" --no-cpu-baseline > ../$O/r2c5_ab_${tag}_old.json 2> ../$O/r2c5_ab_${tag}_old.err); }
for rep in 1 2; do
  old hulls_$rep --mesh hulls --steps 20 --warmup 5
  run hulls_$rep --mesh hulls --steps 20 --warmup 5
  run hullscanon_$rep --mesh hulls --steps 20 --warmup 5 --order canonical
done
old terrain --mesh terrain --steps 10 --warmup 3
run terrain --mesh terrain --steps 10 --warmup 3
old render --mesh render --steps 5 --warmup 3
run render --mesh render --steps 5 --warmup 3
run rendercanon --mesh render --steps 5 --warmup 3 --order canonical
old c2 --workload c2 --steps 3 --warmup 3
run c2 --only c2 --steps 3 --warmup 3
run c2canon --only c2 --steps 3 --warmup 3 --order canonical
old c4 --workload c4 --steps 5 --warmup 3
run c4 --only c4 --steps 5 --warmup 3
D=swift-game-engine_b200/csrc
for L in libcq $(cd $D && ls libcq_ray_*.so | sed 's/\.so$//'); do
  CQ_LIB=$D/$L.so timeout 200 python bench.py --only c5 --steps 5 --warmup 3 --no-cpu-baseline > $O/r2c5_ab_c5_$L.json 2> $O/r2c5_ab_c5_$L.err
  CQ_LIB=$D/$L.so timeout 200 python bench.py --only c5 --steps 5 --warmup 3 --no-cpu-baseline --order canonical > $O/r2c5_ab_c5canon_$L.json 2> $O/r2c5_ab_c5canon_$L.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c5_ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        pq = d["roofline"].get("per_query", {})
        print(f, "%.1f M/s  %.3f ms/step  e2e %.1f M/s (%.2f ms)  evals/q %s" % (d["value"] / 1e6, d["ms_per_step"], e.get("value", 0) / 1e6,
              e.get("ms_per_step", 0), pq.get("distance_evals")))
    except Exception as ex:
        print(f, "ERR", ex)
PY
( time timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2c5_bench_default.json 2> $O/r2c5_bench_default.err ) 2> $O/r2c5_bench_default.time
echo "bench default rc=$?"; tail -3 $O/r2c5_bench_default.time
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c5_bench_default.json").read().strip().splitlines()[-1])
print("headline %.1f M/s %.3f ms  e2e %.1f M/s (%.2f ms) match %s  full %.1f M/s  replay %s  cpu %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6,
      d["e2e"]["ms_per_step"], d["e2e"]["matches_device_path"], d["e2e"]["full_record"]["value"]/1e6, d["config"]["counting_replay_matches_timed_run"], d["cpu_baseline"]))
for k, x in d["extra"].items():
    if "error" in x: print("  ", k, "ERROR", x["error"], x.get("trace","")[-300:]); continue
    print("   extra %-8s %.1f M/s  %.3f ms/step  frac %.3f  e2e %.1f M/s  wall %.1f s  cpu %s" % (k, x["value"]/1e6, x["ms_per_step"], x["roofline"]["frac"] if "roofline" in x else -1,
          (x.get("e2e") or {}).get("value", 0)/1e6, x.get("bench_wall_s", 0), {kk: v for kk, v in (x.get("cpu_baseline") or {}).items() if kk != "sample"}))
PY
