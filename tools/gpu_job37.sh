#!/bin/bash
# final single-GPU verification of the round
mkdir -p gpurun_out
exec > gpurun_out/job37.log 2>&1
NCU=/usr/local/cuda/bin/ncu
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest all"; timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== default bench"; timeout 1500 python bench.py 2> gpurun_out/job37_bench.err | tail -1 > gpurun_out/r02_bench_c4_n1_final.json; tail -3 gpurun_out/job37_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_c4_n1_final.json"))
print(d["value"], d["e2e"]["value"], d["checksum"], d["e2e"]["checksum"], d["roofline"]["frac"], d["roofline_synthesis"]["frac"], d["stage_ms_per_step"], d["cpu_baseline"]["value"], d["gpu_launches"], d["clocks"])
PY
echo "== launch list C2"
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_c2_final.csv python bench.py --config C2 --steps 1 --warmup 1 --no-cpu --no-e2e > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_c2_final.csv 2>&1 | head -24
echo "== full: ring2 cap fwd final"
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:ring2_kernel -c 1 -o gpurun_out/r02_fft2_cap_fwd_final -f python tools/fft_ab.py --nside 4096 --ncomp 4 > /dev/null 2>&1
ls -la gpurun_out/r02_fft2_cap_fwd_final.ncu-rep
