#!/bin/bash
# N = 8: multi-GPU parity record + default bench + lanes A/B
mkdir -p gpurun_out
exec > gpurun_out/job18.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518"
echo "== dist_check n8"; timeout 600 $TR tools/dist_check.py --nside 1024 --niter 3 2>&1 | grep -E "dist_check|Error|error" | tail -5
echo "== bench default n8"
timeout 900 $TR bench.py --gpus 8 2> gpurun_out/job18_n8.err | tail -1 > gpurun_out/r02_bench_c4_n8.json
tail -2 gpurun_out/job18_n8.err | cut -c1-500
echo "== bench n8 lanes 2 (device arm only)"
HCU_BENCH_LANES=2 timeout 600 $TR bench.py --gpus 8 --steps 2 --warmup 2 --no-cpu --no-e2e 2> gpurun_out/job18_n8_l2.err | tail -1 > gpurun_out/r02_bench_c4_n8_lanes2.json
tail -1 gpurun_out/job18_n8_l2.err | cut -c1-500
echo "== reference arm under torchrun (rank 0 only)"
timeout 600 $TR bench.py --gpus 8 --impl reference --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
