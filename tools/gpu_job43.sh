#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job43.log 2>&1
echo "== pytest map/overlap/replay/dist"; timeout 900 python -m pytest tests/test_gpu_map.py tests/test_gpu_overlap.py tests/test_fields_replay.py tests/test_gpu_dist.py -q 2>&1 | tail -3
echo "== bench C4"; HCU_BENCH_VERBOSE=1 timeout 1200 python bench.py --config C4 --steps 1 --warmup 1 --no-cpu 2> gpurun_out/job43.err | tail -1 > gpurun_out/job43_c4.json; grep -E "e2e|device-resident" gpurun_out/job43.err | cut -c1-200
python - <<'PY'
import json
d=json.load(open("gpurun_out/job43_c4.json"))
print(d["value"], d["e2e"]["value"], d["checksum"], d["e2e"]["checksum"], d["roofline_ringfft"])
PY
