#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job21.log 2>&1
for ns in 16 64 256; do echo "== AB nside $ns"; timeout 300 python tools/fft_ab.py --nside $ns --ncomp 3 2>&1 | tail -3; done
echo "== AB nside 64 lmax 256"; timeout 300 python tools/fft_ab.py --nside 64 --lmax 256 --ncomp 2 2>&1 | tail -3
echo "== AB nside 128 lmax 100"; timeout 300 python tools/fft_ab.py --nside 128 --lmax 100 --ncomp 2 2>&1 | tail -3
echo "== AB nside 1024 time"; timeout 300 python tools/fft_ab.py --nside 1024 --ncomp 4 --time 2>&1 | tail -7
echo "== AB nside 4096 time"; timeout 600 python tools/fft_ab.py --nside 4096 --ncomp 4 --time --reps 2 2>&1 | tail -7
echo "== AB nside 4096 12 comps time"; timeout 600 python tools/fft_ab.py --nside 4096 --ncomp 12 --time --reps 2 2>&1 | tail -7
NCU=/usr/local/cuda/bin/ncu
echo "== launch list"
timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ring2_kernel --csv --log-file gpurun_out/r02_fft2_launches.csv python tools/fft_ab.py --nside 4096 --ncomp 4 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02_fft2_launches.csv")) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value"); ig=hdr.index("Grid Size"); ib=hdr.index("Block Size"); im=hdr.index("Metric Name"); ii=hdr.index("ID")
for r in rows[1:]:
    print(r[ii], r[ik][22:42], r[ig], r[ib], r[im], r[iv])
PY
