#!/bin/bash
# GPU session 2: gen2 after latency fixes; wide variant; ablation; nside 8192 + big Bluestein tests; C3 parity
mkdir -p gpurun_out
O=gpurun_out/job2
exec > $O.log 2>&1
echo "== pytest (gen2 NW16 default)"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
P="timeout 300 python tools/prof_sht.py --nside 2048 --niter 1 --reps 2"
for NW in 12 16; do
  export HCU_LEGENDRE_NW=$NW
  echo "== timing gen2 nw $NW"
  $P --nmaps 8 --spin 0 2>&1 | tail -1
  $P --nmaps 8 --spin 2 2>&1 | tail -1
done
export HCU_LEGENDRE_NW=8
echo "== wide (16 comps, analysis only)"
timeout 300 python tools/prof_sht.py --nside 2048 --niter 0 --reps 2 --nmaps 16 --spin 0 2>&1 | tail -1
timeout 300 python tools/prof_sht.py --nside 2048 --niter 0 --reps 2 --nmaps 16 --spin 2 2>&1 | tail -1
export HCU_LEGENDRE_NW=16
echo "== norec ablation nw16"
HERACLES_CUDA_LIB=$PWD/heracles_b200/lib/exp/lib_norec.so $P --nmaps 8 --spin 2 2>&1 | tail -1
unset HCU_LEGENDRE_NW
echo "== ncu gen2 nw16"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:legendre -c 2 -o gpurun_out/job2_leg python tools/prof_sht.py --nside 1024 --nmaps 8 --spin 2 --niter 1 --reps 1 2>&1 | tail -2
ls -la gpurun_out
