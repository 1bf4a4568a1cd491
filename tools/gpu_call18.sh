#!/bin/bash
# round 2, GPU call 18 (8 GPUs): field 0's 1-D transform on the side stream under the twin's deposit -- NCCL parity, c3 at N = 8
set -u
O=gpurun_out/call18
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node=8 --master-port 29621 tests/run_slab_nccl.py > $O/slab_nccl8.txt 2>&1; echo "rc=$?" >> $O/slab_nccl8.txt
timeout 400 $TR --nproc-per-node=8 --master-port 29622 bench.py --gpus 8 --steps 20 --warmup 3 > $O/bench_c3_8gpu.json 2> $O/bench_c3_8gpu.err; echo "rc=$?" >> $O/bench_c3_8gpu.err
timeout 300 $TR --nproc-per-node=4 --master-port 29623 bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e --no-routing-stress > $O/bench_c3_4gpu.json 2> $O/bench_c3_4gpu.err; echo "rc=$?" >> $O/bench_c3_4gpu.err
tail -2 $O/slab_nccl8.txt; tail -c 200 $O/bench_c3_8gpu.err
