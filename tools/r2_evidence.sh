#!/bin/bash
# Round-2 evidence call: one `ncu --set full` capture per kernel of the path on its BASELINE workload (each after the same
# command exited 0 without the profiler) + the launch list of the headline bench command.  Reports land in gpurun_out/;
# tools/ncu_summary.py condenses them into profiles/r2_*_summary.txt.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python __graft_entry__.py smoke > $O/r2ev_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2ev_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2ev_bench_default.json 2> $O/r2ev_bench_default.err ) 2> $O/r2ev_bench_default.time
echo "bench default rc=$?"; tail -3 $O/r2ev_bench_default.time
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2ev_bench_reference.json 2>/dev/null; echo "reference arm rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2ev_bench_default.json").read().strip().splitlines()[-1])
print("headline %.1f M/s %.3f ms  e2e %.1f M/s (%.2f ms) match %s  full %.1f M/s (%.2f ms)  replay %s  frac %.3f fp32 %.3f" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6,
      d["e2e"]["ms_per_step"], d["e2e"]["matches_device_path"], d["e2e"]["full_record"]["value"]/1e6, d["e2e"]["full_record"]["ms_per_step"], d["config"]["counting_replay_matches_timed_run"],
      d["roofline"]["frac"], d["roofline"]["fp32_secondary"]["frac"]))
for k, x in d["extra"].items():
    if "error" in x: print("  ", k, "ERROR", x["error"], x.get("trace","")[-300:]); continue
    print("   extra %-8s %.1f M/s  %.3f ms/step  frac %.3f  e2e %.1f M/s  wall %.1f s  cpu %s" % (k, x["value"]/1e6, x["ms_per_step"], x["roofline"]["frac"] if "roofline" in x else -1,
          (x.get("e2e") or {}).get("value", 0)/1e6, x.get("bench_wall_s", 0), {kk: v for kk, v in (x.get("cpu_baseline") or {}).items() if kk not in ("sample", "unit", "kind")}))
PY
cap() { # tag, kernel regex, skip, count, command...
  local tag=$1 rx=$2 skip=$3 cnt=$4; shift 4
  "$@" > $O/r2ev_${tag}_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o $O/r2ev_$tag "$@" > $O/r2ev_${tag}_ncu.log 2>&1
  echo "ncu $tag rc=$?"
}
B="python bench.py --no-cpu-baseline --no-extras"
cap mas_hulls k_move_and_slide 3 1 $B --steps 2 --warmup 3
cap mas_terrain k_move_and_slide 3 1 $B --mesh terrain --steps 2 --warmup 3
cap mas_render k_move_and_slide 3 1 $B --mesh render --steps 2 --warmup 3
cap cast_c4 k_capsule_cast 3 1 python bench.py --no-cpu-baseline --only c4 --steps 1 --warmup 3
cap cast_c2 k_capsule_cast 3 1 python bench.py --no-cpu-baseline --only c2 --steps 1 --warmup 3
cap ray_c5_ref k_raycast 3 1 python bench.py --no-cpu-baseline --only c5 --steps 1 --warmup 3
cap ray_c5_canon k_raycast 3 1 python bench.py --no-cpu-baseline --only c5 --steps 1 --warmup 3 --order canonical
cap overlap_all k_capsule_overlap_pool 1 1 python tools/profile_extra.py overlap
# (the rounds of a sweep normally run inside a CUDA-graph WHILE node; for the profiler they are launched from the host — same kernels)
cap sep_turns k_sep_turns 40 2 env CQ_SEP_HOST_ROUNDS=1 python tools/profile_extra.py separation
cap sep_post k_sep_post 1 1 env CQ_SEP_HOST_ROUNDS=1 python tools/profile_extra.py separation
cap build_onesweep k_os_pass 4 2 python tools/profile_extra.py build
cap build_karras "k_karras|k_fit|k_collapse4|k_morton" 4 4 python tools/profile_extra.py build
H="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
$H > $O/r2ev_headline_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2ev_launches_headline.csv $H > $O/r2ev_launches_ncu.log 2>&1
echo "ncu launches rc=$?"
# condense on the box (the merge-back limit is 64 MiB): text summaries for every report, the reports themselves only for
# the three kernels whose per-region tables are rebuilt at home
for r in $O/r2ev_*.ncu-rep; do
  t=$(basename $r .ncu-rep)
  python tools/ncu_summary.py $r "$t" > $O/${t}_summary.txt 2>/dev/null
  ncu -i $r --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
if len(rows)>=3:
    h=rows[0]
    for k in range(2,len(rows)):
        d=dict(zip(h,rows[k]))
        print(d.get('Kernel Name','?')[:60], '| dur ns', d.get('gpu__time_duration.sum'), '| dram rd', d.get('dram__bytes_read.sum'), d.get('dram__bytes_read.sum [Mbyte]'), '| wr', d.get('dram__bytes_write.sum'), '| lanes', d.get('smsp__thread_inst_executed_per_inst_executed.ratio'))
" > $O/${t}_raw.txt 2>/dev/null
done
for keep in mas_hulls mas_terrain cast_c2 ray_c5_ref; do mv $O/r2ev_$keep.ncu-rep $O/keep_r2ev_$keep.ncu-rep 2>/dev/null; done
rm -f $O/r2ev_*.ncu-rep
ls -la $O | grep r2ev | head -60
