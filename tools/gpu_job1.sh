#!/bin/bash
# first GPU session of round 2: parity of the second-generation Legendre kernels + A/B timing
mkdir -p gpurun_out
O=gpurun_out/job1
exec > $O.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
echo "== pytest gen2 NW16"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "== pytest sht gen2 NW12"; HCU_LEGENDRE_NW=12 timeout 600 python -m pytest tests/test_gpu_sht.py tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -8
for cfg in "GEN=1" "GEN=2 NW=12" "GEN=2 NW=16"; do
  eval $cfg
  export HCU_LEGENDRE_GEN=$GEN HCU_LEGENDRE_NW=$NW
  echo "== timing gen $GEN nw $NW"
  timeout 200 python tools/prof_sht.py --nside 2048 --nmaps 8 --spin 0 --niter 1 --reps 2 2>&1 | tail -1
  timeout 200 python tools/prof_sht.py --nside 2048 --nmaps 8 --spin 2 --niter 1 --reps 2 2>&1 | tail -1
  timeout 200 python tools/prof_sht.py --nside 2048 --nmaps 4 --spin 2 --niter 1 --reps 2 2>&1 | tail -1
done
unset HCU_LEGENDRE_GEN HCU_LEGENDRE_NW
echo "== bench C4 gen2"; timeout 600 python bench.py --config C4 --steps 1 --warmup 1 --no-cpu --no-e2e 2>&1 | tail -3
echo "== ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:legendre -c 2 -o gpurun_out/job1_leg python tools/prof_sht.py --nside 1024 --nmaps 8 --spin 2 --niter 1 --reps 1 2>&1 | tail -3
ls -la gpurun_out
