#!/bin/bash
# round 2, GPU call 17 (1 GPU): full test-suite after row N4 and the slab path's earlier 1-D transform
set -u
O=gpurun_out/call17
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/pytest.txt
cat $O/pytest.txt
