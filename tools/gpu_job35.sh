#!/bin/bash
# N = 4 sanity of the final code
mkdir -p gpurun_out
exec > gpurun_out/job35.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514"
echo "== dist_check n4"; timeout 600 $TR tools/dist_check.py --nside 1024 --niter 3 2>&1 | grep -E "dist_check|Error|error" | tail -5
echo "== bench default n4"
timeout 900 $TR bench.py --gpus 4 --no-cpu 2> gpurun_out/job35_n4.err | tail -1 > gpurun_out/r02_bench_c4_n4_peer.json
grep -E "device-resident" gpurun_out/job35_n4.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_c4_n4_peer.json"))
print(d["value"], d.get("e2e",{}).get("value"), d["checksum"], d.get("e2e",{}).get("checksum"), d.get("dist_parity",{}).get("max_norm_err"), d["dist_stage_ms_per_rank"], d.get("dist_exchange","")[:20])
PY
