#!/bin/bash
# round 2, GPU call 3: consolidated code -- test-suite, brick-shape / tile-stride A/B, full default bench
set -u
O=gpurun_out/call3
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest.txt
for v in zs32 zs33 b16x8 b16x16 b24x12; do
  for wl in c3 c2; do
    ASTRILD_PK_LIB=$PWD/build/variants/libapk_$v.so timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_${wl}_$v.json 2> $O/bench_${wl}_$v.err
  done
done
ASTRILD_PK_LIB=$PWD/build/variants/libapk_zs33.so timeout 300 python bench.py --workload c2u --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_c2u_zs33.json 2> $O/bench_c2u_zs33.err
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
ls -la $O
