#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job33.log 2>&1
echo "== new feature tests"; timeout 900 python -m pytest tests/test_discrete.py tests/test_io.py tests/test_gpu_overlap.py -x -q 2>&1 | tail -15
