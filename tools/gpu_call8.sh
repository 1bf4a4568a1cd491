#!/bin/bash
# round 2, GPU call 8: bank-queue tile kernel against the particle-parallel one; scatter at 4 CTAs / SM
set -u
O=gpurun_out/call8
mkdir -p $O
APK_TILE=queue timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py tests/test_gpu_zz_small_mesh.py -m gpu -x -q 2>&1 | tail -8 > $O/pytest_queue.txt
for wl in c3 c2 c3s; do
  for v in pp queue; do
    APK_TILE=$v timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_${wl}_$v.json 2> $O/bench_${wl}_$v.err
  done
done
APK_TILE=queue ASTRILD_PK_LIB=$PWD/build/variants/libapk_q3.so timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_c3_queue3.json 2> $O/bench_c3_queue3.err
ASTRILD_PK_LIB=$PWD/build/variants/libapk_sc4.so timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_c3_sc4.json 2> $O/bench_c3_sc4.err
APK_TILE=queue timeout 300 python bench.py --workload c3s --order random --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_c3s_queue_random.json 2> $O/bench_c3s_queue_random.err
cat $O/pytest_queue.txt
