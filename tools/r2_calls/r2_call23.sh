#!/bin/bash
# Round-2 GPU call 23: chunking of cq_crowd_step (e2e): chunk size (CQ_CHUNK; shipped 131072 = 8 chunks per 1 M characters)
# and the "compute bound -> 2 chunks" rule (CQ_CROWD_HINT; shipped 3.0) on the hulls and terrain scenes.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
run() { local tag=$1; shift; timeout 200 python bench.py "$@" --no-cpu-baseline --no-extras --steps 10 --warmup 3 > $O/r2c23_$tag.json 2> $O/r2c23_$tag.err; }
run hulls_default --mesh hulls
CQ_CHUNK=65536 run hulls_c64k --mesh hulls
CQ_CHUNK=262144 run hulls_c256k --mesh hulls
run terrain_default --mesh terrain
CQ_CROWD_HINT=100 run terrain_nohint --mesh terrain
CQ_CROWD_HINT=100 CQ_CHUNK=262144 run terrain_nohint_c256k --mesh terrain
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c23_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); e = d["e2e"]
        print(f, "device %.3f ms  e2e %.1f M/s (%.3f ms) match %s" % (d["ms_per_step"], e["value"] / 1e6, e["ms_per_step"], e.get("matches_device_path")))
    except Exception as ex:
        print(f, "ERR", ex)
PY
