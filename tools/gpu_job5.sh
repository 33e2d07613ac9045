#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job5.log 2>&1
echo "== pytest sht"; timeout 900 python -m pytest tests/test_gpu_sht.py -m gpu -q -x --deselect tests/test_gpu_sht.py::test_sparse_map_nside_8192 2>&1 | tail -4
HCU_LEGENDRE_NW=12 timeout 900 python -m pytest tests/test_gpu_sht.py -m gpu -q -x -k "map2alm or golden or parity" 2>&1 | tail -3
P="timeout 300 python tools/prof_sht.py --nside 2048 --niter 0 --reps 2 --nmaps 8"
for NW in 12 16; do
  export HCU_LEGENDRE_NW=$NW
  echo "== nw $NW"; $P --spin 0 2>&1 | tail -1; $P --spin 2 2>&1 | tail -1
done
