#!/bin/bash
# round 2, GPU call 20 (1 GPU): slab path (k, mu) mode with emulated ranks + the whole test-suite once more
set -u
O=gpurun_out/call20
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $O/pytest.txt
cat $O/pytest.txt
