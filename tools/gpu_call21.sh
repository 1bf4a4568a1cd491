#!/bin/bash
# round 2, GPU call 21 (2 GPUs): NCCL slab script incl. the (k, mu) mode
set -u
O=gpurun_out/call21
mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29631 tests/run_slab_nccl.py > $O/slab_nccl2.txt 2>&1; echo "rc=$?" >> $O/slab_nccl2.txt
tail -5 $O/slab_nccl2.txt
