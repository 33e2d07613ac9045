#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job24.log 2>&1
python - <<'PY'
import torch
p=torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size, "persist max", getattr(p,"persisting_l2_cache_max_size",None), "window max", getattr(p,"access_policy_max_window_size",None))
PY
echo "== default"; timeout 600 python tools/fft_ab.py --nside 4096 --ncomp 4 --time --reps 2 2>&1 | tail -2
echo "== persist"; HCU_R2_PERSIST=1 timeout 600 python tools/fft_ab.py --nside 4096 --ncomp 4 --time --reps 2 2>&1 | tail -3
echo "== persist occ1 nt512"; HCU_R2_PERSIST=1 HCU_R2_BELT_NT=512 HCU_R2_BELT_OCC=1 timeout 600 python tools/fft_ab.py --nside 4096 --ncomp 4 --time --reps 2 2>&1 | tail -3
NCU=/usr/local/cuda/bin/ncu
echo "== launch list persist"
HCU_R2_PERSIST=1 timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ring2_kernel --csv --log-file gpurun_out/r02_fft2_launches_p.csv python tools/fft_ab.py --nside 4096 --ncomp 4 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02_fft2_launches_p.csv")) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value"); ig=hdr.index("Grid Size"); ib=hdr.index("Block Size"); im=hdr.index("Metric Name"); ii=hdr.index("ID")
for r in rows[1:]:
    print(r[ii], r[ik][22:42], r[ig], r[ib], r[im], r[iv])
PY
