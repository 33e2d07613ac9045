#!/bin/bash
# N = 2: peer-memory exchange vs NCCL all-to-all
mkdir -p gpurun_out
exec > gpurun_out/job27.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527"
echo "== dist_check n2 peer"; timeout 600 $TR tools/dist_check.py --nside 1024 --niter 3 2>&1 | grep -E "dist_check|Error|error|Warn" | tail -5
echo "== bench C4 n2 peer"
timeout 900 $TR bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu --no-e2e 2> gpurun_out/job27_peer.err | tail -1 > gpurun_out/r02_bench_c4_n2_peer.json
tail -1 gpurun_out/job27_peer.err | cut -c1-600
echo "== bench C4 n2 nccl"
HCU_DIST_EXCHANGE=nccl timeout 900 $TR bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu --no-e2e 2> gpurun_out/job27_nccl.err | tail -1 > gpurun_out/r02_bench_c4_n2_nccl.json
tail -1 gpurun_out/job27_nccl.err | cut -c1-600
python - <<'PY'
import json
for n in ("peer","nccl"):
    try:
        d=json.load(open(f"gpurun_out/r02_bench_c4_n2_{n}.json"))
        print(n, d["value"], d["checksum"], d.get("dist_parity"), d["dist_stage_ms_per_rank"], d.get("dist_exchange","")[:40])
    except Exception as e: print(n, "failed", e)
PY
