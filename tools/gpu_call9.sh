#!/bin/bash
# round 2, GPU call 9: pipelined single-GPU pair (mesh 0's r2c under mesh 1's tile kernel), clears beside the partition;
# full test-suite; default bench (e2e + cpu baseline) and the reference arm; launch list of the bench step
set -u
O=gpurun_out/call9
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest.txt
for wl in c3 c2 c4; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_$wl.json 2> $O/bench_$wl.err
done
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
CMD="python bench.py --workload c3 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'apk|fft' -c 60 --csv --log-file $O/launches_c3.csv $CMD > $O/ncu_launches.log 2>&1
cat $O/pytest.txt
