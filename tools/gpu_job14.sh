#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job14.log 2>&1
echo "== pytest sht auto"; timeout 900 python -m pytest tests/test_gpu_sht.py tests/test_gpu_dist.py -m gpu -q --deselect tests/test_gpu_sht.py::test_sparse_map_nside_8192 2>&1 | tail -3
echo "== pytest sht gen2"; HCU_LEGENDRE_GEN=2 timeout 900 python -m pytest tests/test_gpu_sht.py -m gpu -q --deselect tests/test_gpu_sht.py::test_sparse_map_nside_8192 2>&1 | tail -3
export HCU_LEGENDRE_GEN=1
echo "== pytest sht gen1"; timeout 900 python -m pytest tests/test_gpu_sht.py -m gpu -q --deselect tests/test_gpu_sht.py::test_sparse_map_nside_8192 2>&1 | tail -3
P="timeout 300 python tools/prof_sht.py --nside 2048 --niter 1 --reps 2"
echo "== gen1"
$P --spin 0 --nmaps 10 2>&1 | tail -1
$P --spin 0 --nmaps 8 2>&1 | tail -1
$P --spin 2 --nmaps 8 2>&1 | tail -1
$P --spin 2 --nmaps 4 2>&1 | tail -1
unset HCU_LEGENDRE_GEN
echo "== auto"
$P --spin 0 --nmaps 10 2>&1 | tail -1
$P --spin 2 --nmaps 4 2>&1 | tail -1
$P --spin 2 --nmaps 20 2>&1 | tail -1
