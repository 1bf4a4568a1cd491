#!/bin/bash
# round 2, GPU call 16 (1 GPU): row N4 ((k, mu) wedges + multipoles) parity and timing; full test-suite
set -u
O=gpurun_out/call16
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest.txt
timeout 300 python tools/kmu_probe.py > $O/kmu_probe.txt 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
cat $O/pytest.txt $O/kmu_probe.txt
