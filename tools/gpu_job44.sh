#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job44.log 2>&1
NCU=/usr/local/cuda/bin/ncu
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:ring2_kernel --launch-skip 2 -c 1 -o gpurun_out/r02_fft2_belt_fwd_final -f python tools/fft_ab.py --nside 4096 --ncomp 4 > /dev/null 2>&1
ls -la gpurun_out/r02_fft2_belt_fwd_final.ncu-rep
