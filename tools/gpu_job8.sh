#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job8.log 2>&1
P="timeout 300 python tools/prof_sht.py --nside 2048 --niter 1 --reps 2"
export HCU_LEGENDRE_NW=12
for G in 1 2; do
  export HCU_LEGENDRE_GEN=$G
  for n in 2 4 8; do echo "== gen $G spin 0 nmaps $n"; $P --spin 0 --nmaps $n 2>&1 | tail -1; done
  for n in 2 4 8; do echo "== gen $G spin 2 nmaps $n"; $P --spin 2 --nmaps $n 2>&1 | tail -1; done
done
