#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job17.log 2>&1
echo "== pytest all"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
P="timeout 300 python tools/prof_sht.py --niter 1 --reps 2 --nside 2048"
echo "== default"; $P --spin 2 --nmaps 8 2>&1 | tail -1
echo "== burst variant (gen1 analysis)"; HERACLES_CUDA_LIB=$PWD/heracles_b200/lib/exp/lib_burst.so $P --spin 2 --nmaps 8 2>&1 | tail -1
echo "== bench C4 quick"; timeout 900 python bench.py --config C4 --steps 1 --warmup 1 --no-cpu 2> gpurun_out/job17.err | tail -1 > gpurun_out/job17_c4.json; tail -2 gpurun_out/job17.err | cut -c1-400
python - <<'PY'
import json
d=json.load(open("gpurun_out/job17_c4.json"))
print(d["value"], d["e2e"]["value"], d["checksum"], d["e2e"]["checksum"], d["roofline"]["frac"], d["roofline_synthesis"]["frac"], d["roofline_map_values"]["achieved"], d["roofline_map_values"]["tile_sorted"])
PY
