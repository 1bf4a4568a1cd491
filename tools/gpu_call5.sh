#!/bin/bash
# round 2, GPU call 5: experiments that decide the next deposit design -- ATOMS with duplicate addresses vs bank
# conflicts, the tile kernel's sensitivity to particle order, one c2c against two r2c transforms
set -u
O=gpurun_out/call5
mkdir -p $O
timeout 120 tools/ubench/atoms > $O/atoms.txt 2>&1
for ord in input cell random; do
  timeout 300 python bench.py --workload c3s --order $ord --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_c3s_$ord.json 2> $O/bench_c3s_$ord.err
done
timeout 300 python tools/fft_probe2.py > $O/fft_probe2.txt 2>&1
tail -n 30 $O/atoms.txt; cat $O/fft_probe2.txt
