#!/bin/bash
# round 2, GPU call 2: paged one-pass partition -- tests, A/B against the two-pass partition, ncu of the new kernels
set -u
O=gpurun_out/call2
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest.txt
for wl in c3 c2 c2u; do
  for k in paged twopass; do
    APK_PARTITION=$k timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_${wl}_${k}.json 2> $O/bench_${wl}_${k}.err
  done
done
APK_BIN_TABLE=1 timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
CMD="python bench.py --workload c3s --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > $O/plain_c3s.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'brick_tile_kernel|brick_partition_kernel' -s 3 -c 3 -o $O/prof_c3s $CMD > $O/ncu_c3s.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file $O/launches_c3s.csv $CMD > $O/ncu_launches.log 2>&1
ls -la $O
