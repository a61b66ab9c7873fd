#!/bin/bash
# Round-2 multi-GPU call: bench.py under torchrun on the N visible GPUs (headline + strong-scaling extras with the
# library's NCCL gather in the timed region), the C++ example on all GPUs, the group test.  usage: r2_multi.sh N
cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r2m${N}_gpus.txt 2>&1
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > $O/r2m${N}_bench.json 2> $O/r2m${N}_bench.err ) 2> $O/r2m${N}_bench.time
echo "bench N=$N rc=$?"; tail -3 $O/r2m${N}_bench.time
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2m{n}_bench.json").read().strip().splitlines()[-1])
    print("headline %.1f M/s %.3f ms  e2e %.1f M/s (%.2f ms)  full %.1f M/s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6,
          d["e2e"]["ms_per_step"], d["e2e"]["full_record"]["value"]/1e6))
    for k, x in d["extra"].items():
        if "error" in x: print("  ", k, "ERROR", x["error"], x.get("trace", "")[-400:]); continue
        print("   extra %-9s %9.1f M/s  %7.3f ms/step  %s  collective %s ms  kernel alone %s ms  e2e %.1f M/s  wall %.1f s" % (k, x["value"]/1e6, x["ms_per_step"],
              x.get("scaling"), x.get("collective_ms"), x.get("kernel_ms_alone"), (x.get("e2e") or {}).get("value", 0)/1e6, x.get("bench_wall_s", 0)))
except Exception as e:
    print("ERR", e); print(open(f"gpurun_out/r2m{n}_bench.err").read()[-2000:])
PY
timeout 600 ./examples/bin/cq_multi_gpu $N 1500 8388608 5 > $O/r2m${N}_example.txt 2>&1; echo "example rc=$?"; cat $O/r2m${N}_example.txt
timeout 600 python -m pytest tests -m gpu -q --no-header -k "group" > $O/r2m${N}_pytest_group.log 2>&1; echo "pytest group rc=$?"; tail -3 $O/r2m${N}_pytest_group.log
timeout 300 python bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $O/r2m${N}_bench_reference.json 2>/dev/null; echo "reference arm rc=$?"; cut -c1-200 $O/r2m${N}_bench_reference.json
