#!/bin/bash
# Round-2 GPU call 21: capsuleOverlapAll drops candidates visited after the kept ones once the overflow is established;
# refused options; full parity; overlap-all timing (1,048,576 capsules pressed into the render mesh).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -rf --no-header > $O/r2c21_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/r2c21_pytest.log; tail -15 $O/r2c21_pytest.log
timeout 300 python tools/profile_extra.py overlap > $O/r2c21_overlap.txt 2>&1; cat $O/r2c21_overlap.txt
timeout 300 python bench.py --only c2 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/r2c21_c2.json 2>/dev/null
timeout 300 python bench.py --only c4 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/r2c21_c4.json 2>/dev/null
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c21_c*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.1f M/s  %.3f ms/step" % (d["value"] / 1e6, d["ms_per_step"]))
    except Exception as ex:
        print(f, "ERR", ex)
PY
