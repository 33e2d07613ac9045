#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job38.log 2>&1
for ns in 2 4 8 16 64 256; do echo "== AB nside $ns"; timeout 300 python tools/fft_ab.py --nside $ns --ncomp 3 2>&1 | tail -3; done
echo "== AB nside 64 lmax 256"; timeout 300 python tools/fft_ab.py --nside 64 --lmax 256 --ncomp 2 2>&1 | tail -3
echo "== AB nside 4 lmax 16"; timeout 300 python tools/fft_ab.py --nside 4 --lmax 16 --ncomp 2 2>&1 | tail -3
echo "== pytest sht + dist + map"; timeout 1200 python -m pytest tests/test_gpu_sht.py tests/test_gpu_dist.py tests/test_fields_replay.py -x -q 2>&1 | tail -3
echo "== C2 quick"; timeout 600 python bench.py --config C2 --steps 2 --warmup 1 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['checksum'], d['stage_ms_per_step'])"
