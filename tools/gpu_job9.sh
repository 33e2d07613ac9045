#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job9.log 2>&1
echo "== pytest all"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== bench C4"; timeout 900 python bench.py --config C4 --steps 1 --warmup 1 --no-cpu > gpurun_out/job9_c4.json 2> gpurun_out/job9_c4.err; tail -c 1500 gpurun_out/job9_c4.err; cut -c1-600 gpurun_out/job9_c4.json
