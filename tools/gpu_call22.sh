#!/bin/bash
# round 2, GPU call 22 (1 GPU): ncu --set full of the final build's kernels at the bench size (c3, one timed step)
set -u
O=gpurun_out/call22
mkdir -p $O
CMD="python bench.py --workload c3 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > $O/plain_c3.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'brick_|bin_power' -o $O/prof_c3_final $CMD > $O/ncu_c3.log 2>&1
tail -3 $O/ncu_c3.log | cut -c1-200
