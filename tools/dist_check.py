"""
torchrun --nproc-per-node N tools/dist_check.py [--nside 256] -- multi-GPU parity check:
the ring-block / m-distributed transform and Cl over N ranks (NCCL) against the single-GPU
hcu_map2alm / hcu_alm2cl of rank 0 on the same maps.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import heracles_b200 as hb
from heracles_b200.dist import DistributedPipeline

ap = argparse.ArgumentParser()
ap.add_argument("--nside", type=int, default=256)
ap.add_argument("--niter", type=int, default=3)
ap.add_argument("--backend", default="nccl", help="gloo: host collectives (with --same-device the ranks can share one GPU)")
ap.add_argument("--same-device", action="store_true", help="every rank on cuda:0 (peer-memory exchange between processes of one GPU)")
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
if args.same_device:
    local = 0
torch.cuda.set_device(local)
if args.backend == "nccl":
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
else:
    dist.init_process_group(args.backend)
nside, lmax = args.nside, 2 * args.nside
mapper = hb.CudaHealpixMapper(nside, lmax, deconvolve=False, niter=args.niter, device=local)
npix = 12 * nside * nside
rng = np.random.default_rng(7)
npos, nshe = 3, 5
full_pos = rng.standard_normal((npos, npix))
full_she = rng.standard_normal((nshe, 2, npix))
# every rank holds a share of the maps; the shares sum to the full maps
share = (rank + 1) / (world * (world + 1) / 2)
dp = DistributedPipeline(mapper, npos, nshe)
for i in range(npos):
    m = mapper.create(spin=0)
    m[:] = full_pos[i] * share
    dp.put(0, i, m)
for i in range(nshe):
    m = mapper.create(2, spin=2)
    m[:] = full_she[i] * share
    dp.put(2, i, m)
del m
cl = dp.spectra().cpu().numpy()
torch.cuda.synchronize()
ok = True
if rank == 0:
    a0 = np.asarray(mapper.transform(full_pos, spin=0))
    a2 = np.asarray(mapper.transform(full_she, spin=2)).reshape(2 * nshe, -1)
    alm = np.concatenate([a0, a2])
    ref = np.asarray(hb.alm2cl(alm, alm))
    n = alm.shape[0]
    auto = np.array([ref[i, i] for i in range(n)])
    err = 0.0
    for i in range(n):
        for j in range(i, n):
            err = max(err, float(np.max(np.abs(cl[i, j] - ref[i, j]) / np.sqrt(np.abs(auto[i] * auto[j]) + 1e-300))))
    ok = err < 1e-10
    print(f"dist_check world={world} nside={nside} niter={args.niter}: max |dCl| / sqrt(Cl_ii Cl_jj) = {err:.3e} -> {'OK' if ok else 'FAIL'}; "
          f"exchanged {dp.transform.exchanged_bytes / 1e6:.1f} MB per rank through "
          f"{'peer memory' if dp.transform.lanes[0].peers is not None else 'nccl all-to-all'}")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
