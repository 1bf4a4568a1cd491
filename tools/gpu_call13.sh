#!/bin/bash
# round 2, GPU call 13 (8 GPUs): after the route fix -- NCCL parity with heavy routing at 8 and 4 ranks (two calls per
# runner), c3 at N = 8 and 4 with the host-buffer leg and the routing stress
set -u
O=gpurun_out/call13
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node=8 --master-port 29601 tests/run_slab_nccl.py > $O/slab_nccl8.txt 2>&1; echo "rc=$?" >> $O/slab_nccl8.txt
timeout 240 $TR --nproc-per-node=4 --master-port 29602 tests/run_slab_nccl.py > $O/slab_nccl4.txt 2>&1; echo "rc=$?" >> $O/slab_nccl4.txt
timeout 400 $TR --nproc-per-node=8 --master-port 29603 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_c3_8gpu.json 2> $O/bench_c3_8gpu.err; echo "rc=$?" >> $O/bench_c3_8gpu.err
timeout 300 $TR --nproc-per-node=4 --master-port 29604 bench.py --gpus 4 --steps 10 --warmup 3 > $O/bench_c3_4gpu.json 2> $O/bench_c3_4gpu.err; echo "rc=$?" >> $O/bench_c3_4gpu.err
tail -2 $O/slab_nccl8.txt; tail -2 $O/slab_nccl4.txt; tail -c 300 $O/bench_c3_8gpu.err; tail -c 300 $O/bench_c3_4gpu.err
