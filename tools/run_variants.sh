# usage: bash tools/run_variants.sh [lib ...]  -- time the SHT stages with alternative builds of the library
for v in default "$@"; do
  echo "== $v"
  if [ "$v" != default ]; then export HERACLES_CUDA_LIB=$PWD/heracles_b200/lib/exp/lib_$v.so; fi
  timeout 120 python tools/prof_sht.py --nside 2048 --nmaps 12 --spin 0 --niter 1 --reps 2 2>&1 | tail -1
  timeout 120 python tools/prof_sht.py --nside 2048 --nmaps 8 --spin 2 --niter 1 --reps 2 2>&1 | tail -1
done
