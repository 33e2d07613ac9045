"""host-to-map rate of CudaHealpixMapper.map_page with pinned catalogue pages (the PCIe-bound part of the end-to-end arm)
  python tools/map_e2e.py [--nside 4096] [--pages 200] [--rows 1000000]     (HCU_SLOT_ROWS: rows per staging slot)"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import heracles_b200 as hb

ap = argparse.ArgumentParser()
ap.add_argument("--nside", type=int, default=4096)
ap.add_argument("--pages", type=int, default=200)
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--pool", type=int, default=32)
a = ap.parse_args()
g = torch.Generator().manual_seed(3)
cols = {}
for k in ("lon", "lat", "w", "g1", "g2"):
    t = torch.empty(a.pool, a.rows, dtype=torch.float64, pin_memory=True)
    if k == "lon":
        t.uniform_(0, 360, generator=g)
    elif k == "lat":
        t.uniform_(-1, 1, generator=g).asin_().mul_(180 / np.pi)
    elif k == "w":
        t.uniform_(0.5, 1.5, generator=g)
    else:
        t.normal_(0, 0.3, generator=g)
    cols[k] = t.numpy()
mapper = hb.CudaHealpixMapper(a.nside, 2 * a.nside, deconvolve=False, sync=False, pixel_weights=None)
pos, she, stats = mapper.create(spin=0), mapper.create(2, spin=2), mapper.new_page_stats()
for rep in range(2):
    mapper.context.synchronize()
    t0 = time.perf_counter()
    for p in range(a.pages):
        j = p % a.pool
        mapper.map_page(cols["lon"][j], cols["lat"][j], cols["w"][j], cols["g1"][j], cols["g2"][j], pos=pos, she=she, stats=stats)
    mapper.context.synchronize()
    dt = time.perf_counter() - t0
    print(f"rep {rep}: {a.pages} pages x {a.rows} rows, 5 columns: {dt:.3f} s = {a.pages * a.rows * 40 / dt / 1e9:.1f} GB/s host to map "
          f"(HCU_SLOT_ROWS={os.environ.get('HCU_SLOT_ROWS', 'default')})")
