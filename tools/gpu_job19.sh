#!/bin/bash
# ring FFT generation 2: A/B against generation 1, timings, SHT parity tests
mkdir -p gpurun_out
exec > gpurun_out/job19.log 2>&1
for ns in 16 32 64 256; do echo "== AB nside $ns"; timeout 300 python tools/fft_ab.py --nside $ns --ncomp 3 2>&1 | tail -4; done
echo "== AB nside 8 lmax 32 (old path only)"; timeout 300 python tools/fft_ab.py --nside 8 --lmax 32 2>&1 | tail -3
echo "== AB nside 64 lmax 256"; timeout 300 python tools/fft_ab.py --nside 64 --lmax 256 --ncomp 2 2>&1 | tail -3
echo "== AB nside 1024 time"; timeout 300 python tools/fft_ab.py --nside 1024 --ncomp 4 --time 2>&1 | tail -8
echo "== AB nside 4096 time"; timeout 600 python tools/fft_ab.py --nside 4096 --ncomp 4 --time --reps 2 2>&1 | tail -8
echo "== pytest sht"; timeout 1200 python -m pytest tests/test_gpu_sht.py -x -q 2>&1 | tail -8
