#!/bin/bash
# round 2, GPU call 23 (1 GPU): compute-sanitizer memcheck over a small run of every hand-written kernel family
set -u
O=gpurun_out/call23
mkdir -p $O
timeout 300 python tools/sanitize_small.py > $O/plain.txt 2>&1 &&
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > $O/memcheck.txt 2>&1
echo "rc=$?" >> $O/memcheck.txt
tail -3 $O/plain.txt; tail -8 $O/memcheck.txt
