#!/bin/bash
# round 2, GPU call 27 (1 GPU): tile flush by x-planes with a running mesh pointer (one LDS per column); rule "k & 1";
# scatter with its four returning atomics in flight together
set -u
O=gpurun_out/call27
mkdir -p $O
for v in n1 n2 n3 n4 n5 n6; do
  export ASTRILD_PK_LIB=$PWD/build/variants/libapk_$v.so
  timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_c3_$v.json 2> $O/bench_c3_$v.err
  python - $O/bench_c3_$v.json $v <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d['stages']['ms']
    print(sys.argv[2], 'ms/step', round(d['ms_per_step'], 3), 'count', round(m['dep_count'], 3), 'scatter', round(m['dep_scatter'], 3), 'tile', round(m['dep_deposit'], 3), 'check', d['check']['ok'], d['check']['max_rel_P'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
done
for v in n4 n3; do
  export ASTRILD_PK_LIB=$PWD/build/variants/libapk_$v.so
  timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_c2_$v.json 2> $O/bench_c2_$v.err
  python - $O/bench_c2_$v.json c2_$v <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
m = d['stages']['ms']
print(sys.argv[2], 'ms/step', round(d['ms_per_step'], 3), 'count', round(m['dep_count'], 3), 'scatter', round(m['dep_scatter'], 3), 'tile', round(m['dep_deposit'], 3), 'check', d['check']['ok'], d['check']['max_rel_P'])
PY
done
