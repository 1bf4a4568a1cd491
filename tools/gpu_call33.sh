#!/bin/bash
# round 2, GPU call 33 (1 GPU): the GPU test-suite after the tolerance of the mesh comparison follows the fixed-point
# quantum of 8191-particle chunks
set -u
O=gpurun_out/call33
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/pytest.txt
cat $O/pytest.txt
