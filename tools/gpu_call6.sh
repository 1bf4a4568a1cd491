#!/bin/bash
# round 2, GPU call 6: lean partition (one filing for the interlaced pair) + subnormal fixed-point tile kernel
set -u
O=gpurun_out/call6
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest.txt
for wl in c3 c2 c3s c4; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_$wl.json 2> $O/bench_$wl.err
done
timeout 300 python bench.py --workload c2u --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_c2u.json 2> $O/bench_c2u.err
cat $O/pytest.txt
