"""
Generates tests/golden/sht_c2.npz and sht_c3.npz: the CPU oracle's map2alm of seeded random maps at
BASELINE.json's configurations C2 (nside 1024, lmax 2048; 3 spin-0 maps and 2 spin-2 fields; niter 0
and 3) and C3 (nside 2048, lmax 4096; 2 spin-0 maps and 1 spin-2 field; niter 0).

The full alm (34 / 134 MB per component) are too large for fixtures, so each file keeps, per component,
   * the alm at 6000 seeded (l, m) positions spread over the whole triangle plus the complete
     m = 0, 1, lmax - 1, lmax columns' first / last entries (the corners of the triangle),
   * the L2 norm of the whole alm row,
   * the full auto spectrum C_l (a checksum over every m of every l),
which pins a CUDA transform to the oracle everywhere without shipping the arrays.  The maps are NOT
stored: both sides draw them from numpy.random.default_rng(seed) (same numpy in this image and on the GPU box).

Run here (CPU only, ~20 min on 8 threads):   python tests/golden/make_sht_golden.py
Consumed by tests/test_gpu_sht.py::test_full_map_parity_baseline_configs.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

NSAMPLE = 6000


def maps_for(nside, n, seed):
    """the seeded maps both sides use: standard normal, one row per component"""
    return np.random.default_rng(seed).standard_normal((n, 12 * nside * nside))


def sample_index(lmax, seed):
    rng = np.random.default_rng(seed)
    nalm = (lmax + 1) * (lmax + 2) // 2
    idx = rng.integers(0, nalm, NSAMPLE)
    corners = []
    for m in (0, 1, 2, 3, lmax // 2, lmax - 1, lmax):
        base = m * (2 * lmax + 1 - m) // 2
        for l in (m, m + 1, m + 2, (m + lmax) // 2, lmax - 1, lmax):
            if m <= l <= lmax:
                corners.append(base + l)
    return np.unique(np.concatenate([idx, np.array(corners, dtype=np.int64)]))


def digest(alm, idx):
    return dict(samples=alm[:, idx], norm=np.sqrt((np.abs(alm) ** 2).sum(axis=1)),
                cl=np.stack([oracle.alm2cl(a, a) for a in alm]))


def run(name, nside, lmax, n0, n2, niters, seed):
    out = {"nside": nside, "lmax": lmax, "n0": n0, "n2": n2, "seed": seed, "niters": np.array(niters)}
    idx = sample_index(lmax, seed + 1)
    out["index"] = idx
    m0 = maps_for(nside, n0, seed)
    m2 = maps_for(nside, 2 * n2, seed + 7)
    for niter in niters:
        t = time.perf_counter()
        a0 = oracle.map2alm(nside, lmax, m0, spin=0, niter=niter)
        a2 = oracle.map2alm(nside, lmax, m2, spin=2, niter=niter)
        for spin, a in ((0, a0), (2, a2)):
            for k, v in digest(a, idx).items():
                out[f"s{spin}_n{niter}_{k}"] = v
        print(f"{name}: niter {niter} done in {time.perf_counter() - t:.0f} s", flush=True)
    np.savez_compressed(os.path.join(HERE, name), **out)


if __name__ == "__main__":
    oracle.build()
    which = sys.argv[1:] or ["c2", "c3"]
    if "c2" in which:
        run("sht_c2.npz", 1024, 2048, 3, 2, [0, 3], 1234)
    if "c3" in which:
        run("sht_c3.npz", 2048, 4096, 2, 1, [0], 4321)
