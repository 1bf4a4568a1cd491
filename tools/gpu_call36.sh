#!/bin/bash
# round 2, GPU call 36 (1 GPU): tile kernel with the warp's 96 payload words loaded coalesced and handed out by three
# shuffles (APK_TILE_COOP_LOAD = 1) against the in-tree build -- does SHFL share the load/store data pipe?
set -u
O=gpurun_out/call36
mkdir -p $O
for v in coop main; do
  if [ "$v" != main ]; then export ASTRILD_PK_LIB=$PWD/build/variants/libapk_$v.so; else unset ASTRILD_PK_LIB; fi
  timeout 100 python bench.py --workload c3 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_c3_$v.json 2> $O/bench_c3_$v.err
  python - $O/bench_c3_$v.json c3_$v <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d['stages']['ms']
    print(sys.argv[2], 'ms/step', round(d['ms_per_step'], 3), 'count', round(m['dep_count'], 3), 'scatter', round(m['dep_scatter'], 3), 'tile', round(m['dep_deposit'], 3), 'check', d['check']['ok'], d['check']['max_rel_P'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
done
