#!/bin/bash
# round 2, GPU call 37 (1 GPU): the GPU test-suite at the round's final tree
set -u
O=gpurun_out/call37
mkdir -p $O
timeout 70 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > $O/pytest.txt
cat $O/pytest.txt
