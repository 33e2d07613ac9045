#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job34.log 2>&1
echo "== tiny first (deadlock guard)"; timeout 120 python tools/prof_sht.py --niter 1 --reps 1 --nside 64 --spin 2 --nmaps 8 2>&1 | tail -1 || { echo "TINY FAILED/HUNG"; exit 1; }
timeout 120 python tools/prof_sht.py --niter 1 --reps 1 --nside 64 --spin 0 --nmaps 10 2>&1 | tail -1 || { echo "TINY FAILED/HUNG"; exit 1; }
P="timeout 300 python tools/prof_sht.py --niter 1 --reps 2 --nside 2048"
echo "== pingpong spin2 8 maps"; $P --spin 2 --nmaps 8 2>&1 | tail -1
echo "== nopp     spin2 8 maps"; HERACLES_CUDA_LIB=$PWD/heracles_b200/lib/exp/lib_nopp.so $P --spin 2 --nmaps 8 2>&1 | tail -1
echo "== pingpong spin2 4 maps"; $P --spin 2 --nmaps 4 2>&1 | tail -1
echo "== nopp     spin2 4 maps"; HERACLES_CUDA_LIB=$PWD/heracles_b200/lib/exp/lib_nopp.so $P --spin 2 --nmaps 4 2>&1 | tail -1
echo "== pingpong spin0 10 maps"; $P --spin 0 --nmaps 10 2>&1 | tail -1
echo "== nopp     spin0 10 maps"; HERACLES_CUDA_LIB=$PWD/heracles_b200/lib/exp/lib_nopp.so $P --spin 0 --nmaps 10 2>&1 | tail -1
echo "== pytest sht"; timeout 1200 python -m pytest tests/test_gpu_sht.py -x -q 2>&1 | tail -3
