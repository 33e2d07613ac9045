#!/bin/bash
# C5 with the final code + final checks
mkdir -p gpurun_out
exec > gpurun_out/job41.log 2>&1
echo "== C5 bench (niter 0)"; timeout 900 python bench.py --config C5 --niter 0 --steps 1 --warmup 0 > gpurun_out/r02_bench_c5.json 2> gpurun_out/job41_c5.err; tail -c 400 gpurun_out/job41_c5.err; cut -c1-1200 gpurun_out/r02_bench_c5.json
echo; echo "== pytest all"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "== bench C1 full line"; timeout 600 python bench.py --config C1 --steps 2 --warmup 3 2>/dev/null | tail -1 | cut -c1-600
