#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job6.log 2>&1
echo "== pytest sht"; timeout 900 python -m pytest tests/test_gpu_sht.py -m gpu -q -x --deselect tests/test_gpu_sht.py::test_sparse_map_nside_8192 2>&1 | tail -4
P="timeout 300 python tools/prof_sht.py --nside 2048 --niter 1 --reps 2 --nmaps 8"
for NW in 12 16; do
  export HCU_LEGENDRE_NW=$NW
  echo "== nw $NW"; $P --spin 0 2>&1 | tail -1; $P --spin 2 2>&1 | tail -1
done
export HCU_LEGENDRE_NW=12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:legendre -c 2 -o gpurun_out/job6_leg python tools/prof_sht.py --nside 1024 --nmaps 8 --spin 2 --niter 1 --reps 1 2>&1 | tail -1
