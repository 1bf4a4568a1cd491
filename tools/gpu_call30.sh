#!/bin/bash
# round 2, GPU call 30 (1 GPU): brick shapes under the rotated tile kernel with the plane flush (per-brick overhead --
# clear, flush, halo REDs -- against tile size / CTAs per SM)
set -u
O=gpurun_out/call30
mkdir -p $O
for v in main b12x12 b16x8 b12x8 b16x6 b8x6; do
  if [ "$v" != main ]; then export ASTRILD_PK_LIB=$PWD/build/variants/libapk_$v.so; else unset ASTRILD_PK_LIB; fi
  for wl in c3; do
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
unset ASTRILD_PK_LIB
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); m=d['stages']['ms']; print('c2_main', round(d['ms_per_step'],3), 'tile', round(m['dep_deposit'],3), d['check']['ok'])"
