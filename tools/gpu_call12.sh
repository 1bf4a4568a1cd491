#!/bin/bash
# round 2, GPU call 12 (1 GPU): full test-suite with the catalog / folding rows, default bench incl. the PkBatch leg,
# launch list of exactly the timed steps (cudaProfilerStart/Stop)
set -u
O=gpurun_out/call12
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err
timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 > $O/bench_c2.json 2> $O/bench_c2.err
CMD="python bench.py --workload c3 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_c3.csv $CMD > $O/ncu_launches.log 2>&1
cat $O/pytest.txt; tail -c 400 $O/bench_default.err
