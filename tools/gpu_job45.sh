#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job45.log 2>&1
for ns in 4 16 64; do echo "== AB nside $ns"; timeout 200 python tools/fft_ab.py --nside $ns --ncomp 3 2>&1 | tail -3; done
echo "== AB nside 64 lmax 256"; timeout 200 python tools/fft_ab.py --nside 64 --lmax 256 --ncomp 2 2>&1 | tail -3
echo "== 4096 time"; timeout 300 python tools/fft_ab.py --nside 4096 --ncomp 4 --time --reps 3 2>&1 | tail -5
echo "== pytest sht+dist"; timeout 600 python -m pytest tests/test_gpu_sht.py tests/test_gpu_dist.py -x -q 2>&1 | tail -2
