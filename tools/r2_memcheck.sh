#!/bin/bash
# Round-2 GPU call: compute-sanitizer memcheck over the parity tests that exercise the round's new device code (warp-reduced
# commit, drops at pickup / in the walk, slim pair state, WHILE-graph separation rounds, dirty refit, phased rays).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -q -x --no-header -p no:cacheprovider \
  -k "capsule_cast_bit_exact or capsule_overlap or move_and_slide_bit_exact or agent_separation or parameter_edge or raycast_vs_oracle or dirty_subtree or c2_sweeps_against_semla or kinematic_platforms" \
  > $O/r2_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a $O/r2_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" $O/r2_memcheck.log | head -20
tail -5 $O/r2_memcheck.log
