#!/bin/bash
# N = 8: peer-memory exchange, refit ring blocks, fused ring FFT
mkdir -p gpurun_out
exec > gpurun_out/job31.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518"
echo "== dist_check n8"; timeout 600 $TR tools/dist_check.py --nside 1024 --niter 3 2>&1 | grep -E "dist_check|Error|error" | tail -5
echo "== bench default n8"
HCU_BENCH_VERBOSE=1 timeout 900 $TR bench.py --gpus 8 2> gpurun_out/job31_n8.err | tail -1 > gpurun_out/r02_bench_c4_n8_peer.json
grep -E "e2e rank|device-resident" gpurun_out/job31_n8.err | cut -c1-500
echo "== bench n8 lanes 2 (device arm only)"
HCU_BENCH_LANES=2 timeout 600 $TR bench.py --gpus 8 --steps 2 --warmup 2 --no-cpu --no-e2e 2> gpurun_out/job31_n8_l2.err | tail -1 > gpurun_out/r02_bench_c4_n8_peer_lanes2.json
tail -1 gpurun_out/job31_n8_l2.err | cut -c1-500
python - <<'PY'
import json
for n in ("r02_bench_c4_n8_peer","r02_bench_c4_n8_peer_lanes2"):
    try:
        d=json.load(open(f"gpurun_out/{n}.json"))
        print(n, d["value"], d.get("e2e",{}).get("value"), d["checksum"], d.get("dist_parity",{}).get("max_norm_err"), d["dist_stage_ms_per_rank"], d.get("dist_exchange","")[:30])
    except Exception as e: print(n, "failed", e)
PY
