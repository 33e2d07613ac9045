#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job25.log 2>&1
echo "== pytest all"; timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
P="timeout 300 python tools/prof_sht.py --niter 1 --reps 2 --nside 2048"
echo "== sht nside 2048 spin2 8 maps"; $P --spin 2 --nmaps 8 2>&1 | tail -1
echo "== sht gen1 fft"; HCU_RINGFFT_GEN=1 $P --spin 2 --nmaps 8 2>&1 | tail -1
echo "== bench C4 quick"; timeout 900 python bench.py --config C4 --steps 1 --warmup 1 --no-cpu 2> gpurun_out/job25.err | tail -1 > gpurun_out/job25_c4.json; tail -2 gpurun_out/job25.err | cut -c1-400
python - <<'PY'
import json
d=json.load(open("gpurun_out/job25_c4.json"))
print(d["value"], d["e2e"]["value"], d["checksum"], d["e2e"]["checksum"], d["roofline"]["frac"], d["stage_ms_per_step"])
PY
