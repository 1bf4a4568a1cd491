#!/bin/bash
# round 2, GPU call 1: box facts, ATOMS microbenchmark, GPU test-suite, tile-kernel A/B, fixtures
set -u
mkdir -p gpurun_out
O=gpurun_out/call1
mkdir -p $O
{ nproc; free -g; nvidia-smi -L; } > $O/box.txt 2>&1
( cd tools/ubench && timeout 120 ./atoms ) > $O/atoms.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > $O/pytest.txt
for wl in c3 c2; do
  for k in pp cell; do
    APK_TILE_KERNEL=$k timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_${wl}_${k}.json 2> $O/bench_${wl}_${k}.err
  done
done
APK_BIN_TABLE=1 timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_c3_bintable.json 2> $O/bench_c3_bintable.err
APK_BIN_TABLE=1 timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_c2_bintable.json 2> $O/bench_c2_bintable.err
timeout 1500 python tools/make_fixtures.py c2 c3s c4s c3 c4 > $O/fixtures.txt 2>&1
ls -la gpurun_out >> $O/box.txt
