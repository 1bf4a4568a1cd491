#!/bin/bash
# round 2, GPU call 31 (1 GPU): larger bricks (longer runs of equal keys in the partition, less halo in the flush)
set -u
O=gpurun_out/call31
mkdir -p $O
for v in b12x12 b16x12 b12x16 b24x8 b16x10; do
  export ASTRILD_PK_LIB=$PWD/build/variants/libapk_$v.so
  for wl in c3 c2; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_${wl}_$v.json 2> $O/bench_${wl}_$v.err
  python - $O/bench_${wl}_$v.json ${wl}_$v <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d['stages']['ms']
    print(sys.argv[2], 'ms/step', round(d['ms_per_step'], 3), 'count', round(m['dep_count'], 3), 'scatter', round(m['dep_scatter'], 3), 'tile', round(m['dep_deposit'], 3), 'check', d['check']['ok'], d['check']['max_rel_P'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
  done
done
