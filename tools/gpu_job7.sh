#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job7.log 2>&1
echo "== pytest sht"; timeout 900 python -m pytest tests/test_gpu_sht.py -m gpu -q -x --deselect tests/test_gpu_sht.py::test_sparse_map_nside_8192 2>&1 | tail -3
P="timeout 300 python tools/prof_sht.py --nside 2048 --niter 1 --reps 2 --nmaps 8"
for NW in 12 16; do
  export HCU_LEGENDRE_NW=$NW
  echo "== nw $NW default (reduce at 6)"; $P --spin 0 2>&1 | tail -1; $P --spin 2 2>&1 | tail -1
  for v in hcu_chore_reduce7 hcu_chore_reduce4 hcu_syn_step4; do
    echo "== nw $NW $v"; HERACLES_CUDA_LIB=$PWD/heracles_b200/lib/exp/lib_$v.so $P --spin 0 2>&1 | tail -1; HERACLES_CUDA_LIB=$PWD/heracles_b200/lib/exp/lib_$v.so $P --spin 2 2>&1 | tail -1
  done
done
echo "== gen1"; HCU_LEGENDRE_GEN=1 $P --spin 0 2>&1 | tail -1;  HCU_LEGENDRE_GEN=1 $P --spin 2 2>&1 | tail -1
