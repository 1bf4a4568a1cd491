#!/bin/bash
# round 2, GPU call 32 (1 GPU): final single-GPU records of the round -- test-suite, default bench with every leg,
# reference arm, c2 / c4, smoke, launch list of the timed steps, ncu --set full of the kernels at the bench size
set -u
O=gpurun_out/call32
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_c4.json 2> $O/bench_c4.err
timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
CMD="python bench.py --workload c3 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_c3.csv $CMD > $O/ncu_launches.log 2>&1
CMD1="python bench.py --workload c3 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'brick_|bin_power' -o $O/prof_c3_final $CMD1 > $O/ncu_c3.log 2>&1
cat $O/pytest.txt $O/smoke.txt
tail -2 $O/ncu_c3.log | cut -c1-200
for f in bench_default bench_c2 bench_c4; do python - $O/$f.json $f <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d['stages']['ms']
    print(sys.argv[2], 'ms/step', round(d['ms_per_step'], 3), 'e2e', (d.get('e2e') or {}).get('ms_per_step'), {k: round(v, 3) for k, v in m.items()}, 'check', d['check']['ok'], d['check']['max_rel_P'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
done
tail -c 700 $O/bench_reference.json
