#!/bin/bash
# Round-2 GPU call 3: parity after the commit / second-pass rework, the new bench line with extras, A/B of the look-ahead
# prune and of the two order rules, the multi-GPU example on one GPU, launch list of the headline command.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
D=swift-game-engine_b200/csrc
timeout 1200 python -m pytest tests -m gpu -q -rf --no-header > $O/r2c3_pytest.log 2>&1
echo "pytest default rc=$?" | tee -a $O/r2c3_pytest.log
tail -12 $O/r2c3_pytest.log
CQ_LIB=$D/libcq_la.so timeout 1200 python -m pytest tests -m gpu -q -rf --no-header > $O/r2c3_pytest_la.log 2>&1
echo "pytest lookahead rc=$?" | tee -a $O/r2c3_pytest_la.log
tail -5 $O/r2c3_pytest_la.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2c3_bench_default.json 2> $O/r2c3_bench_default.err ) 2> $O/r2c3_bench_default.time
echo "bench default rc=$?"; cat $O/r2c3_bench_default.time
run() { # tag, lib, args...
  local tag=$1 lib=$2; shift 2
  CQ_LIB=$D/$lib timeout 300 python bench.py "$@" --no-cpu-baseline --no-extras > $O/r2c3_ab_${tag}.json 2> $O/r2c3_ab_${tag}.err
}
for L in libcq libcq_la; do
  run hulls_$L $L.so --mesh hulls --steps 20 --warmup 5
  run hullscanon_$L $L.so --mesh hulls --steps 20 --warmup 5 --order canonical
  run terrain_$L $L.so --mesh terrain --steps 10 --warmup 3
  run render_$L $L.so --mesh render --steps 5 --warmup 3
  run c2_$L $L.so --only c2 --steps 3 --warmup 3
  run c4_$L $L.so --only c4 --steps 5 --warmup 3
done
run c2canon_libcq libcq.so --only c2 --steps 3 --warmup 3 --order canonical
run c5_libcq libcq.so --only c5 --steps 5 --warmup 3
run c5canon_libcq libcq.so --only c5 --steps 5 --warmup 3 --order canonical
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c3_ab_*.json")) + ["gpurun_out/r2c3_bench_default.json"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        pq = d["roofline"].get("per_query", {})
        print(f, "%.1f M/s  %.3f ms/step  e2e %.1f M/s (%.2f ms)  evals/q %s" % (d["value"] / 1e6, d["ms_per_step"], e.get("value", 0) / 1e6,
              e.get("ms_per_step", 0), pq.get("distance_evals")))
        for k, x in (d.get("extra") or {}).items():
            if "error" in x:
                print("   extra", k, "ERROR", x["error"]); continue
            print("   extra %-8s %.1f M/s  %.3f ms/step  frac %.3f  e2e %.1f M/s  wall %.1f s  parity %s" % (k, x["value"] / 1e6, x["ms_per_step"],
                  x["roofline"]["frac"] if "roofline" in x else -1, (x.get("e2e") or {}).get("value", 0) / 1e6, x.get("bench_wall_s", 0),
                  {kk: v for kk, v in (x.get("cpu_baseline") or {}).items() if "identical" in kk}))
    except Exception as ex:
        print(f, "ERR", ex)
PY
timeout 300 ./examples/bin/cq_multi_gpu 1 1000 4194304 3 > $O/r2c3_multi_gpu_example.txt 2>&1; echo "example rc=$?"; cat $O/r2c3_multi_gpu_example.txt
# launch list of the headline command (shares, not absolutes)
H="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
$H > $O/r2c3_headline_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2c3_launches_headline.csv $H > $O/r2c3_launches_ncu.log 2>&1
echo "ncu launches rc=$?"
