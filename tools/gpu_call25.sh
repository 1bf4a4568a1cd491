#!/bin/bash
# round 2, GPU call 25 (1 GPU): occupancy of the rotated tile kernel (8 CTAs/SM at 32 registers; 128 / 512 threads per
# brick) and an ncu capture of it at 512^3
set -u
O=gpurun_out/call25
mkdir -p $O
for v in rot1 rot1c8 rot1t128 rot1t512 rot0c8; do
  export ASTRILD_PK_LIB=$PWD/build/variants/libapk_$v.so
  timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_c3_$v.json 2> $O/bench_c3_$v.err
  python - $O/bench_c3_$v.json $v <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], 'ms/step', round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['stages']['ms'].items()}, 'check', d['check']['ok'], d['check']['max_rel_P'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
done
export ASTRILD_PK_LIB=$PWD/build/variants/libapk_rot1.so
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:brick_tile -c 2 -o $O/prof_c3s_rot1 python bench.py --workload c3s --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_rot1.log 2>&1
tail -2 $O/ncu_rot1.log
