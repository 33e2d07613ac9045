#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/job16.log 2>&1
echo "== pytest sht+dist+fields (auto, start table on)"; timeout 900 python -m pytest tests/test_gpu_sht.py tests/test_gpu_dist.py tests/test_fields_replay.py tests/test_gpu_dices.py -m gpu -q 2>&1 | tail -4
echo "== pytest sht gen1 forced"; HCU_LEGENDRE_GEN=1 timeout 900 python -m pytest tests/test_gpu_sht.py -m gpu -q --deselect tests/test_gpu_sht.py::test_sparse_map_nside_8192 2>&1 | tail -3
python - <<'PY'
# bit-identity with and without the table (same process: two contexts)
import numpy as np, heracles_b200 as hb
from heracles_b200 import _lib
nside, lmax = 256, 512
rng = np.random.default_rng(1)
m = rng.standard_normal((4, 12*nside*nside))
res = []
for on in (1, 0):
    ctx = hb.get_context(0)
    _lib.check(ctx.lib.hcu_set_start_table(ctx.handle, on))
    mp = hb.CudaHealpixMapper(nside, lmax, deconvolve=False, niter=2, pixel_weights=None)
    a0 = np.asarray(mp.transform(m, spin=0)); a2 = np.asarray(mp.transform(m.reshape(2,2,-1), spin=2))
    res.append((a0.copy(), a2.copy()))
print("start table on/off: spin0 max|d| %.3e spin2 max|d| %.3e (relative to max %.3e)" % (abs(res[0][0]-res[1][0]).max(), abs(res[0][1]-res[1][1]).max(), abs(res[0][0]).max()))
PY
P="timeout 300 python tools/prof_sht.py --niter 1 --reps 2"
for T in 1 0; do
  export HCU_START_TABLE=$T
  echo "== start table $T, nside 2048"
  $P --nside 2048 --spin 0 --nmaps 10 2>&1 | tail -1
  $P --nside 2048 --spin 2 --nmaps 8 2>&1 | tail -1
  echo "== start table $T, nside 4096 spin 2"
  $P --nside 4096 --spin 2 --nmaps 8 --reps 1 2>&1 | tail -1
done
unset HCU_START_TABLE
echo "== burst variant (gen1 analysis)"; HERACLES_CUDA_LIB=$PWD/heracles_b200/lib/exp/lib_burst.so $P --nside 2048 --spin 2 --nmaps 8 2>&1 | tail -1
