#!/bin/bash
# round 2, GPU call 35 (2 GPUs): the NCCL slab path under the final kernels (rotated tile kernel, 24 x 8 bricks):
# parity script against the oracle, config 3 at N = 2 with the golden check (device-resident leg only)
set -u
O=gpurun_out/call35
mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29641 tests/run_slab_nccl.py > $O/slab_nccl2.txt 2>&1; echo "rc=$?" >> $O/slab_nccl2.txt
tail -4 $O/slab_nccl2.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > $O/bench_c3_2gpu.json 2> $O/bench_c3_2gpu.err
echo "rc=$?" >> $O/bench_c3_2gpu.err
tail -c 1500 $O/bench_c3_2gpu.json
