#!/bin/bash
# Round-2 GPU call 8: parity incl. the dirty-subtree refit; default build after the readiness rule; refit cost on a big set.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rf --no-header > $O/r2c8_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/r2c8_pytest.log; tail -15 $O/r2c8_pytest.log
for M in hulls terrain render; do
  timeout 300 python bench.py --mesh $M --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/r2c8_ab_$M.json 2> $O/r2c8_ab_$M.err
done
timeout 300 python bench.py --only c5 --steps 5 --warmup 3 --no-cpu-baseline > $O/r2c8_ab_c5.json 2> $O/r2c8_ab_c5.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c8_ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        print(f, "%.1f M/s  %.3f ms/step  e2e %.1f M/s (%.2f ms)" % (d["value"] / 1e6, d["ms_per_step"], e.get("value", 0) / 1e6, e.get("ms_per_step", 0)), d["config"].get("refit_ms_mean"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
# refit cost: one 12-triangle box moving inside the 10 M-triangle static set, dirty path vs full path, both orders
python - <<'PY' > gpurun_out/r2c8_refit_cost.txt 2>&1
import importlib, os, subprocess, sys, json
code = r"""
import importlib, sys, time, numpy as np
cq = importlib.import_module("swift-game-engine_b200")
sc = cq.scenes
order = cq.ORDER_REFERENCE if sys.argv[1] == "reference" else cq.ORDER_CANONICAL
parts, half = sc.terrain_scene()
bv, bi = sc.box_mesh(3.0)
parts = parts + [sc.part(bv, bi, sc.trs_model((10, 20, 10)), entity_id=7)]
g = cq.CollisionQuery(parts, order=order)
ms = []
for k in range(6):
    g.update_transforms([7], [sc.trs_model((10 + k, 20, 10))])
    ms.append(g.info()["refit_ms"])
print(sys.argv[1], "full" if "CQ_REFIT_FULL" in __import__("os").environ else "dirty", "refit_ms of one 12-triangle static part in a 10 M-triangle set:", [round(x, 4) for x in ms])
"""
for order in ("reference", "canonical"):
    for full in (False, True):
        env = dict(os.environ)
        if full: env["CQ_REFIT_FULL"] = "1"
        print(subprocess.run([sys.executable, "-c", code, order], env=env, capture_output=True, text=True).stdout.strip())
PY
cat gpurun_out/r2c8_refit_cost.txt
