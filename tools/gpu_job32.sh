#!/bin/bash
# N = 8: cached peer exchange; lanes 1 vs 2 through the whole bench line
mkdir -p gpurun_out
exec > gpurun_out/job32.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518"
echo "== bench default n8 (lanes 1)"
HCU_BENCH_VERBOSE=1 timeout 900 $TR bench.py --gpus 8 2> gpurun_out/job32_n8.err | tail -1 > gpurun_out/r02_bench_c4_n8_peer.json
grep -E "e2e rank|device-resident" gpurun_out/job32_n8.err | cut -c1-300
echo "== bench n8 lanes 2"
HCU_BENCH_LANES=2 HCU_DIST_LANES=2 HCU_BENCH_VERBOSE=1 timeout 600 $TR bench.py --gpus 8 --no-cpu 2> gpurun_out/job32_n8_l2.err | tail -1 > gpurun_out/r02_bench_c4_n8_peer_lanes2.json
grep -E "e2e rank|device-resident" gpurun_out/job32_n8_l2.err | cut -c1-300
python - <<'PY'
import json
for n in ("r02_bench_c4_n8_peer","r02_bench_c4_n8_peer_lanes2"):
    try:
        d=json.load(open(f"gpurun_out/{n}.json"))
        print(n, d["value"], d.get("e2e",{}).get("value"), d["checksum"], d.get("e2e",{}).get("checksum"), d.get("dist_parity",{}).get("max_norm_err"), d["dist_stage_ms_per_rank"]["a2a"])
    except Exception as e: print(n, "failed", e)
PY
