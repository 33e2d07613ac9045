#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job40.log 2>&1
T="timeout 300 python tools/fft_ab.py --ncomp 4 --time --reps 3"
echo "== 4096 default (belt 512)"; $T --nside 4096 2>&1 | tail -3
echo "== 4096 belt 256"; HCU_R2_BELT_NT=256 $T --nside 4096 2>&1 | tail -2
echo "== 4096 rest 512"; HCU_R2_REST_NT=512 $T --nside 4096 2>&1 | tail -2
echo "== 4096 rest 128"; HCU_R2_REST_NT=128 $T --nside 4096 2>&1 | tail -2
echo "== 2048 default"; $T --nside 2048 2>&1 | tail -3
echo "== 2048 belt 256"; HCU_R2_BELT_NT=256 $T --nside 2048 2>&1 | tail -2
echo "== 2048 belt 512"; HCU_R2_BELT_NT=512 $T --nside 2048 2>&1 | tail -2
echo "== 2048 rest 256"; HCU_R2_REST_NT=256 $T --nside 2048 2>&1 | tail -2
echo "== 1024 default"; $T --nside 1024 2>&1 | tail -3
echo "== 1024 belt 256"; HCU_R2_BELT_NT=256 $T --nside 1024 2>&1 | tail -2
echo "== 1024 belt 512"; HCU_R2_BELT_NT=512 $T --nside 1024 2>&1 | tail -2
