#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job20.log 2>&1
NCU=/usr/local/cuda/bin/ncu
echo "== launch list"
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -k regex:ring2_kernel --csv --log-file gpurun_out/r02_fft2_launches.csv python tools/fft_ab.py --nside 4096 --ncomp 4 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02_fft2_launches.csv")) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value"); ig=hdr.index("Grid Size"); ib=hdr.index("Block Size")
for r in rows[1:]:
    print(r[ik][:40], r[ig], r[ib], r[iv])
PY
echo "== full: cap fwd M=8192"
timeout 900 $NCU --set full --clock-control none --import-source on -k regex:ring2_kernel -c 1 -o gpurun_out/r02_fft2_cap_fwd -f python tools/fft_ab.py --nside 4096 --ncomp 4 > /dev/null 2>&1
echo "== full: belt fwd"
timeout 900 $NCU --set full --clock-control none --import-source on -k regex:ring2_kernel --launch-skip 10 -c 1 -o gpurun_out/r02_fft2_belt_fwd -f python tools/fft_ab.py --nside 4096 --ncomp 4 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
