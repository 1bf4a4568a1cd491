#!/bin/bash
# round 2, GPU call 7: ncu evidence at the BENCH size (c3, 1024^3): launch list + --set full of the deposit kernels
set -u
O=gpurun_out/call7
mkdir -p $O
CMD="python bench.py --workload c3 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > $O/plain_c3.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'brick_tile_kernel|brick_scatter_kernel|brick_count_kernel|bin_power' -s 4 -c 5 -o $O/prof_c3 $CMD > $O/ncu_c3.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 80 --csv --log-file $O/launches_c3.csv $CMD > $O/ncu_launches.log 2>&1
tail -3 $O/ncu_c3.log | cut -c1-300
ls -la $O
