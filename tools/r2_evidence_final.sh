#!/bin/bash
# Last GPU call of round 2: the shipped build's smoke, default bench line, reference arm, launch list, and fresh ncu captures of
# the kernels that changed after tools/r2_evidence.sh ran (same recipe; the other summaries under profiles/ are from that call).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -rf --no-header > $O/r2fin_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2fin_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/r2fin_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2fin_smoke.log | cut -c1-200
( time timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2fin_bench_default.json 2> $O/r2fin_bench_default.err ) 2> $O/r2fin_bench_default.time
echo "bench default rc=$?"; tail -3 $O/r2fin_bench_default.time
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2fin_bench_reference.json 2>/dev/null; echo "reference arm rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2fin_bench_default.json").read().strip().splitlines()[-1])
print("headline %.1f M/s %.3f ms  e2e %.1f M/s (%.2f ms) match %s  frac %.3f" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6,
      d["e2e"]["ms_per_step"], d["e2e"]["matches_device_path"], d["roofline"]["frac"]))
for k, x in d["extra"].items():
    if "error" in x: print("  ", k, "ERROR", x["error"]); continue
    print("   extra %-8s %.1f M/s  %.3f ms/step  frac %.3f  e2e %.1f M/s" % (k, x["value"]/1e6, x["ms_per_step"], x["roofline"]["frac"] if "roofline" in x else -1, (x.get("e2e") or {}).get("value", 0)/1e6))
PY
cap() { local tag=$1 rx=$2 skip=$3 cnt=$4; shift 4
  "$@" > $O/r2fin_${tag}_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o $O/r2fin_$tag "$@" > $O/r2fin_${tag}_ncu.log 2>&1
  echo "ncu $tag rc=$?"; }
B="python bench.py --no-cpu-baseline --no-extras"
cap mas_hulls k_move_and_slide 3 1 $B --steps 2 --warmup 3
cap cast_c2 k_capsule_cast 3 1 python bench.py --no-cpu-baseline --only c2 --steps 1 --warmup 3
cap overlap_all k_capsule_overlap_pool 1 1 python tools/profile_extra.py overlap
H="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
$H > $O/r2fin_headline_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2fin_launches_headline.csv $H > $O/r2fin_launches_ncu.log 2>&1
echo "ncu launches rc=$?"
for r in $O/r2fin_*.ncu-rep; do t=$(basename $r .ncu-rep); python tools/ncu_summary.py $r "$t" > $O/${t}_summary.txt 2>/dev/null; done
rm -f $O/r2fin_*.ncu-rep
ls $O | grep r2fin | head -40
