#!/bin/bash
# Round-2 GPU call 11: drops at pickup (tie-hopeless / box lower bound) in the query kernels against the build without them,
# parity, the ray-sort probe, and a probe of CUDA-graph WHILE nodes (device-side round loop for agent separation).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
D=swift-game-engine_b200/csrc
timeout 60 examples/bin/cond_graph_probe > $O/r2c11_cond_graph_probe.txt 2>&1; cat $O/r2c11_cond_graph_probe.txt
timeout 1800 python -m pytest tests -m gpu -q -rf --no-header > $O/r2c11_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/r2c11_pytest.log; tail -12 $O/r2c11_pytest.log
run() { local tag=$1 lib=$2; shift 2; CQ_LIB=$D/$lib.so timeout 300 python bench.py "$@" --no-cpu-baseline --no-extras > $O/r2c11_ab_${tag}_$lib.json 2> $O/r2c11_ab_${tag}_$lib.err; }
for L in libcq libcq_nodrop; do
  run c2 $L --only c2 --steps 3 --warmup 3
  run c4 $L --only c4 --steps 5 --warmup 3
done
run c2b libcq --only c2 --steps 3 --warmup 3
run c2b libcq_nodrop --only c2 --steps 3 --warmup 3
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c11_ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        pq = d["roofline"].get("per_query", {})
        print(f, "%.1f M/s  %.3f ms/step  e2e %.1f M/s (%.2f ms)  evals/q %s cands %s" % (d["value"] / 1e6, d["ms_per_step"], e.get("value", 0) / 1e6,
              e.get("ms_per_step", 0), pq.get("distance_evals"), pq.get("candidates")))
    except Exception as ex:
        print(f, "ERR", ex)
PY
timeout 600 python tools/ray_sort_probe.py > $O/r2c11_ray_sort_probe.txt 2>&1; cat $O/r2c11_ray_sort_probe.txt
