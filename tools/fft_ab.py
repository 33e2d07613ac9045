"""A/B of the two ring-FFT generations (k_ringfft.cu vs k_ringfft2.cu) through hcu_map2phase / hcu_phase2map:
per-ring-pair agreement of the phase rows and of the synthesised pixels, and the stage times of both.

  python tools/fft_ab.py --nside 256 [--lmax 512] [--ncomp 3] [--time]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from heracles_b200 import _lib
from heracles_b200.dist import StagedKernels

ap = argparse.ArgumentParser()
ap.add_argument("--nside", type=int, default=256)
ap.add_argument("--lmax", type=int, default=0)
ap.add_argument("--ncomp", type=int, default=3)
ap.add_argument("--time", action="store_true")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
nside, lmax, nc = a.nside, a.lmax or 2 * a.nside, a.ncomp
npix, nrp = 12 * nside * nside, 2 * nside

os.environ["HCU_RINGFFT_GEN"] = "1"
ctx1 = _lib.Context(0)
k1 = StagedKernels(ctx1, nside, lmax)
gen = torch.Generator(device="cuda").manual_seed(5)
maps = torch.randn(nc, npix, device="cuda", dtype=torch.float64, generator=gen)
mlist = torch.arange(lmax + 1, device="cuda", dtype=torch.int32)
ph1 = torch.zeros(lmax + 1, nrp, nc, 4, device="cuda", dtype=torch.float64)
k1.sync_streams()
k1.map2phase(maps, 0, nrp, mlist, ph1)       # builds the geometry of ctx1 under GEN=1
torch.cuda.synchronize()
os.environ["HCU_RINGFFT_GEN"] = "2"
ctx2 = _lib.Context(0)
k2 = StagedKernels(ctx2, nside, lmax)
k2.sync_streams()
ph2 = torch.full_like(ph1, float("nan"))
k2.map2phase(maps, 0, nrp, mlist, ph2)
torch.cuda.synchronize()


def ring_report(name, d, ref):
    # d, ref: [..., nrp, ...] reduced to per-ring-pair max error relative to the global scale
    scale = float(ref.abs().max())
    bad = torch.nonzero(~(d <= 1e-11 * scale)).flatten().tolist()
    print(f"{name}: max rel err {float(torch.nan_to_num(d, nan=float('inf')).max()) / scale:.3e}; ring pairs off: {len(bad)}"
          + (f" first {bad[:12]} last {bad[-4:]}" if bad else ""))
    return not bad


err = (ph2 - ph1).abs().amax(dim=(0, 2, 3))
err = torch.where(torch.isnan(err), torch.full_like(err, float("inf")), err)
ok = ring_report("forward phase", err, ph1)

# inverse: synthesis-direction rows (reN, imN, reS, imS); m = 0 real
syn = torch.randn(lmax + 1, nrp, nc, 4, device="cuda", dtype=torch.float64, generator=gen)
syn[0, :, :, 1] = 0
syn[0, :, :, 3] = 0
o1 = torch.zeros(nc, npix, device="cuda", dtype=torch.float64)
o2 = torch.full_like(o1, float("nan"))
k1.phase2map(syn, nc, None, 0, nrp, o1)
k2.phase2map(syn, nc, None, 0, nrp, o2)
torch.cuda.synchronize()
d = (o2 - o1).abs().amax(dim=0)
d = torch.where(torch.isnan(d), torch.full_like(d, float("inf")), d).cpu().numpy()
# pixel -> ring pair
ring_start = np.zeros(4 * nside, dtype=np.int64)
ring_len = np.zeros(4 * nside, dtype=np.int64)
s = 0
for r in range(1, 4 * nside):
    ln = 4 * min(r, nside, 4 * nside - r)
    ring_start[r], ring_len[r] = s, ln
    s += ln
per_rp = np.zeros(nrp)
for rp in range(nrp):
    r = rp + 1
    e = d[ring_start[r]:ring_start[r] + ring_len[r]].max()
    rs = 4 * nside - r
    e = max(e, d[ring_start[rs]:ring_start[rs] + ring_len[rs]].max())
    per_rp[rp] = e
ok = ring_report("inverse pixels", torch.from_numpy(per_rp), o1) and ok
print("AB", "OK" if ok else "MISMATCH")

if a.time:
    for name, k in (("gen1", k1), ("gen2", k2)):
        ph = torch.empty_like(ph1)
        out = torch.empty_like(o1)
        for fn, label in ((lambda: k.map2phase(maps, 0, nrp, mlist, ph), "forward"),
                          (lambda: k.phase2map(syn, nc, None, 0, nrp, out), "inverse")):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            gb = (8 * npix + 32 * nrp * (lmax + 1)) * nc / 1e9
            print(f"{name} {label}: {ms:.3f} ms for {nc} components = {gb / ms * 1e3:.0f} GB/s algorithmic")
