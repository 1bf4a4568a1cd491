#!/bin/bash
# round 2, GPU call 4 (2 GPUs): the NCCL slab path against the oracle, bench at N = 2 with the golden check
set -u
O=gpurun_out/call4
mkdir -p $O
nvidia-smi -L > $O/box.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29571 tests/run_slab_nccl.py > $O/slab_nccl.txt 2>&1
echo "rc=$?" >> $O/slab_nccl.txt
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_ingest.py -m gpu -x -q 2>&1 | tail -15 > $O/pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_c3_2gpu.json 2> $O/bench_c3_2gpu.err
echo "rc=$?" >> $O/bench_c3_2gpu.err
tail -c 3000 $O/bench_c3_2gpu.err
