#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job10.log 2>&1
echo "== pytest subset"; timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_cl.py tests/test_gpu_sht.py::test_sparse_map_nside_8192 -m gpu -q 2>&1 | tail -40
