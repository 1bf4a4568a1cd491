#!/bin/bash
# round 2, GPU call 11 (4 GPUs): which feature of the NCCL slab path breaks with more than two ranks
set -u
O=gpurun_out/call11
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node=4"
timeout 200 $TR --master-port 29591 tools/diag_slab_nccl.py > $O/diag_default.txt 2>&1
APK_SLAB_P2P=0 timeout 200 $TR --master-port 29592 tools/diag_slab_nccl.py > $O/diag_nop2p.txt 2>&1
DIAG_NO_SIDE=1 timeout 200 $TR --master-port 29593 tools/diag_slab_nccl.py > $O/diag_noside.txt 2>&1
DIAG_NO_SIDE=1 APK_SLAB_P2P=0 timeout 200 $TR --master-port 29594 tools/diag_slab_nccl.py > $O/diag_neither.txt 2>&1
grep -h DIAG $O/diag_*.txt | cut -c1-260
