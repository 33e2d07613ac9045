#!/bin/bash
# N = 2: multi-GPU parity record + lanes A/B
mkdir -p gpurun_out
exec > gpurun_out/job11.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "== dist_check n2"; timeout 600 $TR tools/dist_check.py --nside 1024 --niter 3 2>&1 | grep -E "dist_check|Error|error" | tail -5
for L in 1 2; do
  echo "== bench C3 n2 lanes $L"
  HCU_BENCH_LANES=$L timeout 900 $TR bench.py --gpus 2 --config C3 --steps 2 --warmup 2 --no-cpu 2> gpurun_out/job11_l$L.err | tail -1 > gpurun_out/job11_c3_l$L.json
  tail -3 gpurun_out/job11_l$L.err | cut -c1-400
  python - <<PY
import json
d=json.load(open("gpurun_out/job11_c3_l$L.json"))
print({k:d[k] for k in ("value","n_gpus","stage_ms_per_step","dist_parity")}, d["e2e"]["value"], d["checksum"], d["e2e"]["checksum"])
print(d.get("dist_stage_ms_per_rank"))
PY
done
