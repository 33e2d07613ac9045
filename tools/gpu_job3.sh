#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job3.log 2>&1
P="timeout 300 python tools/prof_sht.py --nside 2048 --niter 0 --reps 2 --nmaps 8 --spin 2"
for NW in 12 16; do
  export HCU_LEGENDRE_NW=$NW
  echo "== nw $NW full";  $P 2>&1 | tail -1
  for v in noflush norec noflushnorec; do
    echo "== nw $NW $v"; HERACLES_CUDA_LIB=$PWD/heracles_b200/lib/exp/lib_$v.so $P 2>&1 | tail -1
  done
done
