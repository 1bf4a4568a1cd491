#!/bin/bash
# round 2, GPU call 26 (1 GPU): rotated tile kernel -- x-plane skew, y skew, the rule for the k-th lane of a group,
# and the flush with its mesh offsets computed in registers (one LDS per column instead of four)
set -u
O=gpurun_out/call26
mkdir -p $O
for v in r0f r1f r1f_s5 r1f_s13 r1f_s3 r1f_y3x9 r3f_y3x9 r1f_rule1 r1f_rule2; do
  export ASTRILD_PK_LIB=$PWD/build/variants/libapk_$v.so
  timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_c3_$v.json 2> $O/bench_c3_$v.err
  python - $O/bench_c3_$v.json $v <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], 'ms/step', round(d['ms_per_step'], 3), 'tile', round(d['stages']['ms']['dep_deposit'], 3), 'check', d['check']['ok'], d['check']['max_rel_P'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
done
