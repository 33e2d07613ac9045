"""
Host-side logic of the multi-GPU path on CPU: the sharding plan, and the ring-block /
m-distributed exchange of ``heracles_b200.dist.DistributedTransform`` run by two gloo
processes with the oracle standing in for the four CUDA stage kernels.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nside,lmax,world", [(8, 16, 1), (8, 16, 2), (16, 40, 3), (64, 128, 8), (4096, 8192, 8)])
def test_shard_plan(nside, lmax, world):
    from heracles_b200.dist import ShardPlan

    plan = ShardPlan(nside, lmax, world, fft_cost=None, align=1)
    assert plan.rp_bounds[0] == 0 and plan.rp_bounds[-1] == 2 * nside
    assert all(b > a for a, b in zip(plan.rp_bounds, plan.rp_bounds[1:]))
    # pixel ranges of the blocks tile the map exactly once
    ranges = sorted(r for g in range(world) for r in plan.pixel_ranges(g))
    assert ranges[0][0] == 0 and ranges[-1][1] == plan.npix
    assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    # blocks are balanced by pixel count
    sizes = [sum(b - a for a, b in plan.pixel_ranges(g)) for g in range(world)]
    assert max(sizes) - min(sizes) <= 16 * nside + 8 * nside
    # the default plan balances the ring-FFT cost (cap rings are slower): polar blocks get fewer pixels, still a tiling
    wplan = ShardPlan(nside, lmax, world)
    wr = sorted(r for g in range(world) for r in wplan.pixel_ranges(g))
    assert wr[0][0] == 0 and wr[-1][1] == wplan.npix and all(a[1] == b[0] for a, b in zip(wr, wr[1:]))
    if world > 1 and nside >= 64:
        assert sum(b - a for a, b in wplan.pixel_ranges(0)) < sizes[0]
    # ... with boundaries on multiples of the 256 ring pairs one Legendre CTA works on, so that no CTA runs partly empty
    if 2 * nside >= 2 * 256 * world:
        assert all(b % 256 == 0 for b in wplan.rp_bounds)
        assert sum(-(-(b - a) // 256) for a, b in zip(wplan.rp_bounds, wplan.rp_bounds[1:])) == 2 * nside // 256
    # every m has exactly one owner and the row order is the concatenation of the owners' lists
    assert sorted(plan.m_all.tolist()) == list(range(lmax + 1))
    assert all(plan.owner_of_m(int(m)) == g for g in range(world) for m in plan.mlists[g])
    assert np.array_equal(plan.m_all[plan.mpos], np.arange(lmax + 1))
    # Legendre work per rank (sum of lmax - m + 1) is balanced
    work = [int(np.sum(lmax + 1 - plan.mlists[g])) for g in range(world)]
    assert (max(work) - min(work)) / max(work) <= 2.0 * world / (lmax + 1) + 1e-12


def test_ring_start_matches_oracle(oracle):
    from heracles_b200.dist import ShardPlan

    nside = 8
    plan = ShardPlan(nside, 16, 2)
    start, npx, *_ = oracle.ring_table(nside)
    for i in range(1, 2 * nside + 1):
        assert plan.ring_start(i) == start[i - 1]


class OracleKernels:
    """the four stage kernels evaluated with the CPU oracle (tests only)"""

    def __init__(self, oracle, nside, lmax):
        self.o, self.nside, self.lmax = oracle, nside, lmax
        self.nr = 4 * nside - 1

    def batch_size(self, spin):
        return 12 if spin == 0 else 8

    def map2phase(self, maps, rp_lo, rp_hi, mlist, phase):
        m = mlist.numpy()
        nb = maps.shape[0]
        with np.errstate(all="ignore"):
            ph = self.o.map2phase(self.nside, self.lmax, np.nan_to_num(maps.numpy()))
        out = phase.numpy().reshape(len(m), rp_hi - rp_lo, nb, 4)
        for r, rp in enumerate(range(rp_lo, rp_hi)):
            n = ph[:, rp, :][:, m]
            s = ph[:, self.nr - 1 - rp, :][:, m] if self.nr - 1 - rp != rp else np.zeros_like(n)
            out[:, r, :, 0], out[:, r, :, 1] = (n + s).real.T, (n + s).imag.T
            out[:, r, :, 2], out[:, r, :, 3] = (n - s).real.T, (n - s).imag.T

    def phase2alm(self, phase, spin, nb, mlist, rp_lo, rp_hi, alm):
        m = mlist.numpy()
        p = phase.numpy().reshape(len(m), rp_hi - rp_lo, nb, 4)
        full = np.zeros((nb, self.nr, self.lmax + 1), dtype=np.complex128)
        for r, rp in enumerate(range(rp_lo, rp_hi)):
            plus = (p[:, r, :, 0] + 1j * p[:, r, :, 1]).T
            minus = (p[:, r, :, 2] + 1j * p[:, r, :, 3]).T
            if self.nr - 1 - rp == rp:
                full[:, rp, m] = plus
            else:
                full[:, rp, m] = 0.5 * (plus + minus)
                full[:, self.nr - 1 - rp, m] = 0.5 * (plus - minus)
        alm.numpy()[...] += self.o.phase2alm(self.nside, self.lmax, full, spin=spin)

    def alm2phase(self, alm, spin, nb, mlist, rp_lo, rp_hi, phase):
        m = mlist.numpy()
        a = np.zeros_like(alm.numpy())
        for mm in m:
            s = self.o.almidx(self.lmax, mm, mm)
            a[:, s : s + self.lmax - mm + 1] = alm.numpy()[:, s : s + self.lmax - mm + 1]
        ph = self.o.alm2phase(self.nside, self.lmax, a, spin=spin)
        out = phase.numpy().reshape(len(m), rp_hi - rp_lo, nb, 4)
        for r, rp in enumerate(range(rp_lo, rp_hi)):
            n, s = ph[:, rp, :][:, m], ph[:, self.nr - 1 - rp, :][:, m]
            out[:, r, :, 0], out[:, r, :, 1] = n.real.T, n.imag.T
            out[:, r, :, 2], out[:, r, :, 3] = s.real.T, s.imag.T

    def phase2map(self, phase, nb, mpos, rp_lo, rp_hi, maps):
        rows = mpos.numpy()
        p = phase.numpy().reshape(self.lmax + 1, rp_hi - rp_lo, nb, 4)
        full = np.zeros((nb, self.nr, self.lmax + 1), dtype=np.complex128)
        for r, rp in enumerate(range(rp_lo, rp_hi)):
            full[:, rp, :] = (p[rows, r, :, 0] + 1j * p[rows, r, :, 1]).T
            if self.nr - 1 - rp != rp:
                full[:, self.nr - 1 - rp, :] = (p[rows, r, :, 2] + 1j * p[rows, r, :, 3]).T
        m = self.o.phase2map(self.nside, self.lmax, full)
        start, npx, *_ = self.o.ring_table(self.nside)
        out = maps.numpy()
        for rp in range(rp_lo, rp_hi):
            for ring in {rp, self.nr - 1 - rp}:
                out[:, start[ring] : start[ring] + npx[ring]] = m[:, start[ring] : start[ring] + npx[ring]]


def _worker(rank, world, port, nside, lmax, spin, niter, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import oracle
    from heracles_b200.dist import DistributedTransform, ShardPlan, allreduce_cl, reduce_maps

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        oracle.set_num_threads(1)
        rng = np.random.default_rng(123)
        k = 3 if spin == 0 else 2
        full = rng.standard_normal((k, 12 * nside * nside))
        # every rank holds a partial map: the sum over ranks is `full`
        part = torch.from_numpy(full * (rank + 1) / (world * (world + 1) / 2))
        reduce_maps(part)
        plan = ShardPlan(nside, lmax, world)
        tr = DistributedTransform(OracleKernels(oracle, nside, lmax), plan, rank, niter=niter)
        alm = torch.zeros(k, plan.nalm, dtype=torch.complex128)
        fl = 1.0 + 0.01 * np.arange(lmax + 1)
        tr.map2alm(part, spin, alm, fl=fl)
        # entries of foreign m stay zero
        own = np.zeros(plan.nalm, bool)
        for m in plan.mlists[rank]:
            s = oracle.almidx(lmax, m, m)
            own[s : s + lmax - m + 1] = True
        assert np.all(alm.numpy()[:, ~own] == 0)
        # Cl from the m-distributed alm: partial sums + all-reduce
        cl = torch.from_numpy(oracle.alm2cl(alm.numpy(), alm.numpy()) * 0)  # shape only
        a = alm.numpy()
        ell = np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)])
        wgt = np.concatenate([np.full(lmax + 1 - m, 1.0 if m == 0 else 2.0) for m in range(lmax + 1)])
        for i in range(k):
            for j in range(k):
                cl[i, j] = torch.from_numpy(np.bincount(ell, weights=wgt * (a[i] * np.conj(a[j])).real, minlength=lmax + 1) / (2 * np.arange(lmax + 1) + 1))
        allreduce_cl(cl)
        dist.all_reduce(alm)  # gather: the other ranks' entries are zero
        if rank == 0:
            ref = oracle.almxfl(oracle.map2alm(nside, lmax, full, spin=spin, niter=niter), fl)
            err = np.linalg.norm(alm.numpy() - ref) / np.linalg.norm(ref)
            cref = oracle.alm2cl(ref, ref)
            cerr = np.abs(cl.numpy() - cref).max() / np.abs(cref).max()
            ret.put((err, cerr, tr.exchanged_bytes))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("spin,niter,world", [(0, 0, 2), (0, 2, 2), (2, 1, 2), (0, 1, 3)])
def test_distributed_transform_gloo(spin, niter, world):
    import torch.multiprocessing as mp

    nside, lmax = 8, 16
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = 29500 + (os.getpid() + 7 * spin + niter + 13 * world) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, nside, lmax, spin, niter, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    err, cerr, nbytes = ret.get()
    assert err < 1e-12, err
    assert cerr < 1e-12, cerr
    assert nbytes > 0


def test_deconvolution_filter_cache_is_keyed_by_content():
    """DistributedPipeline passes a fresh temporary per spin; a cache keyed on id() served the
    spin-0 pixel window to the spin-2 alm once the first temporary was freed and its id recycled"""
    from heracles_b200.dist import DistributedTransform, ShardPlan

    lmax = 6
    dt = DistributedTransform(None, ShardPlan(4, lmax, 1), 0, niter=0)
    nalm = (lmax + 1) * (lmax + 2) // 2
    for rep in range(8):  # allocate / free the same-sized temporary repeatedly: ids do get reused
        for val in (2.0, 3.0):
            fl = np.full(lmax + 1, val)
            full = dt._fl_full(fl)
            assert full.shape == (nalm,)
            assert float(full.real.min()) == val == float(full.real.max())
            del fl
    assert len([k for k in dt._ws if isinstance(k, tuple)]) == 2


def test_angular_power_spectra_staging_keeps_lazy_alms_alive():
    """angular_power_spectra must not key its upload cache on id() of arrays it does not keep alive
    (a lazily loading alm mapping returns a NEW array per access); the functional test is tests/test_gpu_cl.py"""
    import inspect

    from heracles_b200 import twopoint

    src = inspect.getsource(twopoint.angular_power_spectra)
    assert "id(" not in src and "arrays[key] = (a, d)" in src
