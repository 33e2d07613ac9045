#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job39.log 2>&1
echo "== overlap test"; timeout 600 python -m pytest tests/test_gpu_overlap.py -x -q 2>&1 | tail -3
echo "== bench C4 e2e, transform priority"; HCU_BENCH_VERBOSE=1 timeout 1200 python bench.py --config C4 --steps 1 --warmup 1 --no-cpu 2> gpurun_out/job39.err | tail -1 > gpurun_out/job39_c4.json; grep -E "e2e|device-resident" gpurun_out/job39.err | cut -c1-200
python - <<'PY'
import json
d=json.load(open("gpurun_out/job39_c4.json"))
print(d["value"], d["e2e"]["value"], d["checksum"], d["e2e"]["checksum"])
PY
