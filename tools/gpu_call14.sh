#!/bin/bash
# round 2, GPU call 14 (1 GPU): partition with 4 consecutive particles per thread (16-byte loads, warp runs over 128
# particles) -- tests, benches, ncu of the bench step
set -u
O=gpurun_out/call14
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest.txt
for wl in c3 c2 c4 c2u; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_$wl.json 2> $O/bench_$wl.err
done
timeout 300 python bench.py --workload c3s --order random --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_c3s_random.json 2> $O/bench_c3s_random.err
CMD="python bench.py --workload c3 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > $O/plain_c3.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'brick_|bin_power' -o $O/prof_c3 $CMD > $O/ncu_c3.log 2>&1
cat $O/pytest.txt
