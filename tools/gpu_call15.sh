#!/bin/bash
# round 2, GPU call 15 (8 GPUs): final multi-GPU records -- c3 at N = 8 and 2 with host buffers bound to each GPU's NUMA
# node, c5 at N = 8
set -u
O=gpurun_out/call15
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node=8 --master-port 29611 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_c3_8gpu.json 2> $O/bench_c3_8gpu.err; echo "rc=$?" >> $O/bench_c3_8gpu.err
APK_BENCH_NUMA=0 timeout 400 $TR --nproc-per-node=8 --master-port 29612 bench.py --gpus 8 --steps 5 --warmup 3 --no-routing-stress > $O/bench_c3_8gpu_nonuma.json 2> $O/bench_c3_8gpu_nonuma.err; echo "rc=$?" >> $O/bench_c3_8gpu_nonuma.err
timeout 240 $TR --nproc-per-node=8 --master-port 29613 bench.py --gpus 8 --workload c5 --steps 5 --warmup 2 --no-e2e --no-routing-stress > $O/bench_c5_8gpu.json 2> $O/bench_c5_8gpu.err; echo "rc=$?" >> $O/bench_c5_8gpu.err
timeout 300 $TR --nproc-per-node=2 --master-port 29614 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_c3_2gpu.json 2> $O/bench_c3_2gpu.err; echo "rc=$?" >> $O/bench_c3_2gpu.err
tail -c 200 $O/bench_c3_8gpu.err; tail -c 200 $O/bench_c5_8gpu.err
