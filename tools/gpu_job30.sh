#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job30.log 2>&1
echo "== bench C4 e2e overlap"; HCU_BENCH_VERBOSE=1 timeout 1200 python bench.py --config C4 --steps 1 --warmup 1 --no-cpu 2> gpurun_out/job30.err | tail -1 > gpurun_out/job30_c4.json; grep -E "e2e|device-resident" gpurun_out/job30.err | cut -c1-300
python - <<'PY'
import json
for f in ("job30_c4",):
    d=json.load(open(f"gpurun_out/{f}.json"))
    print(f, d["value"], d["e2e"]["value"], d["checksum"], d["e2e"]["checksum"], d["e2e"]["api"][:80])
PY
