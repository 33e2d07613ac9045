#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job26.log 2>&1
echo "== peer exchange test (2 and 3 processes, one GPU)"; timeout 900 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -15
echo "== nccl mode same harness"; HCU_DIST_EXCHANGE=nccl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py --nside 64 --backend gloo --same-device 2>&1 | grep -E "dist_check|rror" | tail -3
