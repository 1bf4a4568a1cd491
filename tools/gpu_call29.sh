#!/bin/bash
# round 2, GPU call 29 (1 GPU): tile kernel with the twin as a template parameter, plane flush for TSC / table flush for CIC
set -u
O=gpurun_out/call29
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py tests/test_gpu_zz_small_mesh.py tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -5 > $O/pytest.txt
cat $O/pytest.txt
for wl in c3 c2 c4; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_${wl}_main.json 2> $O/bench_${wl}_main.err
  python - $O/bench_${wl}_main.json ${wl}_main <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d['stages']['ms']
    print(sys.argv[2], 'ms/step', round(d['ms_per_step'], 3), 'count', round(m['dep_count'], 3), 'scatter', round(m['dep_scatter'], 3), 'tile', round(m['dep_deposit'], 3), 'check', d['check']['ok'], d['check']['max_rel_P'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
done
