#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job42.log 2>&1
for s in 262144 524288 1048576 2097152; do HCU_SLOT_ROWS=$s timeout 300 python tools/map_e2e.py 2>&1 | tail -1; done
