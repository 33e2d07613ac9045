#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job12.log 2>&1
echo "== pytest all"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30
echo "== C5 bench (niter 0)"; timeout 1200 python bench.py --config C5 --niter 0 --steps 1 --warmup 0 > gpurun_out/job12_c5.json 2> gpurun_out/job12_c5.err; tail -c 600 gpurun_out/job12_c5.err; cut -c1-1500 gpurun_out/job12_c5.json
