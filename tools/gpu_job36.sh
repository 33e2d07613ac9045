#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job36.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527"
for ns in 8 4 2; do
echo "== bench C4 n2 peer, push streams $ns"
HCU_PUSH_STREAMS=$ns timeout 600 $TR bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu --no-e2e 2> gpurun_out/job36.err | tail -1 > gpurun_out/job36_$ns.json
python - <<PY
import json
d=json.load(open("gpurun_out/job36_$ns.json"))
print($ns, d["value"], d["checksum"], d["dist_stage_ms_per_rank"]["a2a"], d["dist_stage_ms_per_rank"]["fft"])
PY
done
