#!/bin/bash
# round 2, GPU call 34 (1 GPU): brick-ordered payload as one array per coordinate (APK_PAYLOAD_SOA = 1) against the
# 12-byte records: deposit parity tests under the variant, then c3 / c2 / c4
set -u
O=gpurun_out/call34
mkdir -p $O
export ASTRILD_PK_LIB=$PWD/build/variants/libapk_soa.so
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py tests/test_gpu_zz_small_mesh.py tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -5 > $O/pytest_soa.txt
cat $O/pytest_soa.txt
for v in soa main; do
  if [ "$v" != main ]; then export ASTRILD_PK_LIB=$PWD/build/variants/libapk_$v.so; else unset ASTRILD_PK_LIB; fi
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
