"""micro-driver for profiling: one map2alm / alm2map call on random maps (device resident)"""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import heracles_b200 as hb
from heracles_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--nside", type=int, default=1024)
ap.add_argument("--lmax", type=int, default=0)
ap.add_argument("--nmaps", type=int, default=10)
ap.add_argument("--spin", type=int, default=0)
ap.add_argument("--niter", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
lmax = a.lmax or 2 * a.nside
ctx = hb.get_context(0)
ctx.set_timing(True)
npix = 12 * a.nside ** 2
nalm = (lmax + 1) * (lmax + 2) // 2
maps = torch.randn(a.nmaps, npix, device="cuda", dtype=torch.float64)
alm = torch.zeros(a.nmaps, nalm, device="cuda", dtype=torch.complex128)
torch.cuda.synchronize()
peak = ctx.fp64_peak()
for r in range(a.reps):
    _lib.check(ctx.lib.hcu_map2alm(ctx.handle, a.nside, lmax, a.spin, a.nmaps, maps.data_ptr(), npix, None, None, a.niter, None, alm.data_ptr(), nalm))
    ms = ctx.sht_timing()
    rec, acc = ctx.sht_work()
    nb = min(a.nmaps, int(ctx.lib.hcu_legendre_batch_size(a.spin)))
    fl = rec * 4 + acc * 4 * nb if a.spin == 0 else rec * 12 + acc * 16 * (nb // 2)
    npass = 1 + a.niter
    syn = f"syn {fl / npass * a.niter / ms[2] / 1e9:.2f} TF/s  " if a.niter else ""
    print(f"rep {r}: fft {ms[0]:.2f} ms  leg_ana {ms[1]:.2f} ms  leg_syn {ms[2]:.2f}  ifft {ms[3]:.2f}  "
          f"ana {fl / ms[1] / 1e9:.2f} TF/s of {peak / 1e12:.1f}  {syn}rec {rec:.3g} acc {acc:.3g} live {acc / max(rec, 1):.2f}")
