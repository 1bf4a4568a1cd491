#!/bin/bash
# round 2, GPU call 19 (1 GPU): final single-GPU records -- test-suite, (k, mu) kernel timing, default bench with every leg,
# reference arm, launch list of the timed steps
set -u
O=gpurun_out/call19
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/pytest.txt
timeout 300 python tools/kmu_probe.py > $O/kmu_probe.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
CMD="python bench.py --workload c3 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_c3.csv $CMD > $O/ncu_launches.log 2>&1
cat $O/pytest.txt $O/kmu_probe.txt $O/smoke.txt
