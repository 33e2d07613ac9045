#!/bin/bash
# final single-GPU measurements of round 2: default bench (C4), C2 pair, launch list, ncu full captures
mkdir -p gpurun_out
exec > gpurun_out/job15.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
echo "== bench default (C4)"; timeout 1500 python bench.py > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/job15_c4.err; tail -c 300 gpurun_out/job15_c4.err; cut -c1-300 gpurun_out/r02_bench_c4_n1.json
echo "== bench reference arm default"; timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_c4_reference.json 2>> gpurun_out/job15_c4.err; cut -c1-400 gpurun_out/r02_bench_c4_reference.json
echo "== bench C2"; timeout 600 python bench.py --config C2 --steps 3 --warmup 3 > gpurun_out/r02_bench_c2_n1.json 2>> gpurun_out/job15_c4.err; cut -c1-300 gpurun_out/r02_bench_c2_n1.json
echo "== reference arm C2 in full"; timeout 1200 python bench.py --impl reference --config C2 --cpu-full --steps 1 --warmup 0 > gpurun_out/r02_bench_c2_reference_full.json 2>> gpurun_out/job15_c4.err; cut -c1-600 gpurun_out/r02_bench_c2_reference_full.json
echo "== ncu launch list C2"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --config C2 --steps 1 --warmup 1 --no-cpu --no-e2e > /dev/null 2>&1; wc -l gpurun_out/r02_launches_c2.csv
echo "== ncu full: C4-shape legendre analysis + synthesis (4 spin-2 fields)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:legendre -c 2 -o gpurun_out/r02_ncu_legendre_c4shape python tools/prof_sht.py --nside 4096 --nmaps 8 --spin 2 --niter 1 --reps 1 2>&1 | tail -1
echo "== ncu full: map_page"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:map_page -c 1 -s 20 -o gpurun_out/r02_ncu_map_page python bench.py --config C3 --steps 1 --warmup 0 --no-cpu --no-e2e 2>&1 | tail -1
ls -la gpurun_out | tail -12
