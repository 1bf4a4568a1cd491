#!/bin/bash
# round 2, GPU call 24 (1 GPU): tile kernel with rotated x-plane order for lanes that start on the same bank
# (APK_TILE_ROT = 1: MATCH.ANY, 2: ballots; x-plane skew 11 or 9 banks) against the in-tree build (ROT = 0)
set -u
O=gpurun_out/call24
mkdir -p $O
for v in main rot1 rot2 rot1s9; do
  if [ "$v" != main ]; then export ASTRILD_PK_LIB=$PWD/build/variants/libapk_$v.so; else unset ASTRILD_PK_LIB; fi
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
unset ASTRILD_PK_LIB
ASTRILD_PK_LIB=$PWD/build/variants/libapk_rot1.so timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_c2_rot1.json 2> $O/bench_c2_rot1.err
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_c2_main.json 2> $O/bench_c2_main.err
tail -c 600 $O/bench_c2_rot1.json | tr ',' '\n' | grep -E "dep_deposit|ok" ; tail -c 600 $O/bench_c2_main.json | tr ',' '\n' | grep -E "dep_deposit|ok"
