#!/bin/bash
# round 2, GPU call 10 (8 GPUs): slab path under the lean partition -- NCCL parity, c3 at N = 8 / 4 / 2, c5 at N = 8
set -u
O=gpurun_out/call10
mkdir -p $O
nvidia-smi -L > $O/box.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node=8 --master-port 29581 tests/run_slab_nccl.py > $O/slab_nccl8.txt 2>&1; echo "rc=$?" >> $O/slab_nccl8.txt
timeout 400 $TR --nproc-per-node=8 --master-port 29582 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_c3_8gpu.json 2> $O/bench_c3_8gpu.err; echo "rc=$?" >> $O/bench_c3_8gpu.err
timeout 240 $TR --nproc-per-node=8 --master-port 29583 bench.py --gpus 8 --workload c5 --steps 5 --warmup 2 --no-e2e --no-routing-stress > $O/bench_c5_8gpu.json 2> $O/bench_c5_8gpu.err; echo "rc=$?" >> $O/bench_c5_8gpu.err
timeout 300 $TR --nproc-per-node=4 --master-port 29584 bench.py --gpus 4 --steps 10 --warmup 3 > $O/bench_c3_4gpu.json 2> $O/bench_c3_4gpu.err; echo "rc=$?" >> $O/bench_c3_4gpu.err
timeout 300 $TR --nproc-per-node=2 --master-port 29585 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_c3_2gpu.json 2> $O/bench_c3_2gpu.err; echo "rc=$?" >> $O/bench_c3_2gpu.err
tail -3 $O/slab_nccl8.txt; tail -c 600 $O/bench_c3_8gpu.err; tail -c 600 $O/bench_c5_8gpu.err
