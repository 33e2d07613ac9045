#!/usr/bin/env python
"""
bench.py -- catalogue -> HEALPix maps -> alm -> Cl, seconds per run.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C1|C2|C3|C4|auto]
                    [--niter I] [--impl b200|reference]

A "step" is one complete pass of the hot path over the whole synthetic
catalogue of the chosen BASELINE.json configuration: every page of every
tomographic bin is mapped (POS: weights; SHE: w*g1, w*g2), the maps are
normalised as the reference's Field layer does (heracles/fields.py:296-304,446),
all maps are transformed (spin 0 and spin 2) and every component cross spectrum
is computed.  Prints ONE JSON line (rank 0).

  value       device-resident arm: catalogue pages already in HBM, C ABI called
              with device pointers.
  e2e         the same pass through the public plugin API
              (CudaHealpixMapper.map_values / heracles_b200.transform /
              angular_power_spectra) with pinned HOST pages; H2D of every page
              and D2H of the spectra are inside the timed region.
  roofline    the dominant kernel (FP64 Legendre analysis): executed flops /
              CUDA-event time against the DFMA peak measured in the same run.
  cpu_baseline  the oracle (CPU restatement of the reference path) timed on a
              bounded sample and scaled to the full configuration.
"""

from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    # name: nside, lmax, bins, rows per bin, fields
    "C1": dict(nside=256, lmax=512, nbins=1, rows=1_000_000, she=False,
               text="nside=256 lmax=512 spin-0 Positions map from 1e6-galaxy synthetic uniform catalogue"),
    "C2": dict(nside=1024, lmax=2048, nbins=1, rows=100_000_000, she=True,
               text="nside=1024 lmax=2048 POS+SHE from 1e8 galaxies, single field pair"),
    "C3": dict(nside=2048, lmax=4096, nbins=5, rows=100_000_000, she=True,
               text="nside=2048 lmax=4096 5 tomographic bins x POS+SHE, all cross-spectra"),
    "C4": dict(nside=4096, lmax=8192, nbins=10, rows=200_000_000, she=True,
               text="nside=4096 lmax=8192 10 tomographic bins x POS+SHE from 2e9 synthetic galaxies"),
}
C5 = dict(nside=8192, lmax=16384, fields=40,
          text="nside=8192 lmax=16384 spin-2 shear map2alm batch of 40 (Q,U) maps (SHT-only stress)")
# small configuration on which bench.py checks the N-GPU result against a 1-GPU run before timing
PARITY_CFG = dict(nside=256, lmax=512, nbins=2, rows=2_000_000, she=True, text="multi-GPU parity check")
PAGE_ROWS = 1_000_000  # heracles/catalog/base.py:315
POOL_PAGES = 64        # distinct pages per bin, cycled to reach the row count
METRIC = "catalog_to_cl_seconds_per_run"
UNIT = "s/run"


# ---------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(smax)) if smax else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def checksum_matrix(cl):
    """ONE checksum for every arm and every N: the sum over l of all component spectra C_l^{ij} with i <= j
    (cl: [ncomp, ncomp, lmax + 1] host array; the lower triangle is ignored)"""
    return float(np.triu(cl.sum(axis=-1)).sum())


def checksum_results(cls):
    """the same number from the dictionary angular_power_spectra returns: every block holds distinct component
    pairs except the (BE) element of a spin-2 auto block, which repeats (EB)"""
    tot = 0.0
    for (k1, k2, i1, i2), c in cls.items():
        c = np.asarray(c)
        tot += float(c.sum())
        if c.ndim == 3 and k1 == k2 and i1 == i2:
            tot -= float(c[1, 0].sum())
    return tot


def pick_config(name, free_bytes):
    if name != "auto":
        return name
    return "C4" if free_bytes > 150e9 else ("C3" if free_bytes > 40e9 else "C2")


def legendre_flops(spin, ncomp, rec, acc):
    """executed flops from the kernels' work counters (SURVEY 8(d) per-cell figures)"""
    if spin == 0:
        return rec * 4.0 + acc * 4.0 * ncomp
    return rec * 12.0 + acc * 16.0 * (ncomp // 2)


def nominal_sht_flops(cfg, niter):
    nalm = (cfg["lmax"] + 1) * (cfg["lmax"] + 2) // 2
    n0 = cfg["nbins"]
    n2 = cfg["nbins"] if cfg["she"] else 0
    per_pass = nalm * (2 * cfg["nside"]) * ((4 + 4 * n0) + ((12 + 16 * n2) if n2 else 0))
    return per_pass * (1 + 2 * niter)


# ---------------------------------------------------------------------------
# the B200 arm
# ---------------------------------------------------------------------------
class Pipeline:
    def __init__(self, cfg, niter, rank, world, torch, hb):
        self.cfg, self.niter, self.rank, self.world = cfg, niter, rank, world
        self.torch, self.hb = torch, hb
        self.ctx = hb.get_context(torch.cuda.current_device())
        self.ctx.set_timing(True)  # stage_ms_per_step / roofline need the per-stage CUDA events
        self.lib = self.ctx.lib
        self.h = self.ctx.handle
        self.stream = torch.cuda.Stream()
        self.ctx.set_stream(self.stream.cuda_stream)
        nside, lmax = cfg["nside"], cfg["lmax"]
        self.npix = 12 * nside * nside
        self.nalm = (lmax + 1) * (lmax + 2) // 2
        self.nbins = cfg["nbins"]
        self.ncomp = self.nbins * (3 if cfg["she"] else 1)
        self.pages = max(1, cfg["rows"] // PAGE_ROWS)
        self.page_rows = min(PAGE_ROWS, cfg["rows"])
        self.pool = min(POOL_PAGES, self.pages)
        self.dist = None
        if world > 1:
            # ring-block / m-distributed transform over the ranks (heracles_b200/dist.py)
            from heracles_b200.dist import DistributedTransform, ShardPlan, attach_peers, make_lanes

            self.plan = ShardPlan(nside, lmax, world)
            # HCU_BENCH_LANES=2: the FFT / all-to-all stages of one Legendre batch are queued under the Legendre kernels of
            # another (measured: no gain, profiles/r02_lanes_c3_n2.txt; default 1).  The asynchronous SHE map reduction
            # gets a communicator of its own (collectives of ONE communicator run in issue order, so it would
            # otherwise queue in front of the spin-0 exchange)
            self.lanes = make_lanes(self.ctx, nside, lmax, int(os.environ.get("HCU_BENCH_LANES", "1")))
            for lane in self.lanes:
                lane.k.ctx.set_timing(True)
            # exchange through NVLink peer memory (HCU_DIST_EXCHANGE=nccl: all_to_all_single)
            attach_peers(self.lanes, lambda lane: lane.k.ctx, self.plan, rank, torch.device("cuda"))
            self.kernels = self.lanes[0].k
            self.dist = DistributedTransform(self.kernels, self.plan, rank, niter=niter, device=torch.device("cuda"), lanes=self.lanes)
            from heracles_b200.dist import _reduce_group

            self.reduce_group = self.lanes[-1].group if len(self.lanes) > 1 else _reduce_group()
        self.make_catalogue()

    # synthetic catalogue: uniform positions, w ~ U(0.5,1.5), g ~ N(0,0.3)  
    def make_catalogue(self):
        torch = self.torch
        dev = torch.device("cuda")
        n = self.pool * self.page_rows
        self.cat = []
        self.norm = []
        for b in range(self.nbins):
            g = torch.Generator(device=dev)
            g.manual_seed(50 + 1000 * b)
            lon = torch.rand(n, generator=g, device=dev, dtype=torch.float64) * 360.0
            lat = torch.rad2deg(torch.asin(torch.rand(n, generator=g, device=dev, dtype=torch.float64) * 2 - 1))
            w = torch.rand(n, generator=g, device=dev, dtype=torch.float64) + 0.5
            cols = {"lon": lon, "lat": lat, "w": w}
            if self.cfg["she"]:
                # per page a contiguous (2, rows) block like np.r_[[re, im]] (fields.py:428)
                # g1, g2 per page as a contiguous (2, rows) block; hcu_map_page forms w g1, w g2 (fields.py:426)
                gg = torch.randn(self.pool, 2, self.page_rows, generator=g, device=dev, dtype=torch.float64) * 0.3
                cols["g"] = gg.contiguous()
            self.cat.append(cols)
            # the Field layer's running means over the whole (cycled) catalogue
            reps = np.bincount(np.arange(self.pages) % self.pool, minlength=self.pool).astype(np.float64)
            wp = w.view(self.pool, self.page_rows).sum(dim=1).cpu().numpy()
            ngal = self.pages * self.page_rows
            wmean = float((wp * reps).sum() / ngal)
            nbar = ngal * wmean / self.npix                      # fields.py:283 (fsky = 1)
            wbar = ngal / (4 * math.pi) * wmean * (4 * math.pi / self.npix)  # fields.py:440
            self.norm.append((nbar, wbar))
        torch.cuda.synchronize()

    def my_pages(self):
        """catalogue rows are sharded across ranks page by page"""
        return range(self.rank, self.pages, self.world)

    def alloc_outputs(self):
        torch = self.torch
        dev = torch.device("cuda")
        self.maps = torch.empty(self.ncomp, self.npix, device=dev, dtype=torch.float64)
        self.alm = torch.empty(self.ncomp, self.nalm, device=dev, dtype=torch.complex128)
        self.cl = torch.empty(self.ncomp, self.ncomp, self.cfg["lmax"] + 1, device=dev, dtype=torch.float64)

    # ---- stages (device-resident) ----
    def stage_map(self):
        """every page of every bin -> POS and SHE maps, one fused hcu_map_page launch per page (one ang2pix)"""
        lib, h, npix, nside = self.lib, self.h, self.npix, self.cfg["nside"]
        self.maps.zero_()
        rows = self.page_rows
        for b in range(self.nbins):
            c = self.cat[b]
            pos_ptr = self.maps[b].data_ptr()
            she_ptr = self.maps[self.nbins + 2 * b].data_ptr() if self.cfg["she"] else None
            lon0, lat0, w0 = c["lon"].data_ptr(), c["lat"].data_ptr(), c["w"].data_ptr()
            g0 = c["g"].data_ptr() if self.cfg["she"] else 0
            for p in self.my_pages():
                off = (p % self.pool) * rows * 8
                g1 = g0 + 2 * off if self.cfg["she"] else None
                g2 = g0 + 2 * off + rows * 8 if self.cfg["she"] else None
                self.check(lib.hcu_map_page(h, nside, 0, lon0 + off, lat0 + off, w0 + off, g1, g2, rows, pos_ptr, she_ptr, npix, None))

    def tile_sorted_rate(self, reps=3):
        """
        SURVEY 8(d): the scatter rate on a TILE-SORTED catalogue -- the rows of every page of bin 0 ordered by their
        nside = 64 NEST parent pixel, the locality a survey catalogue has in file order -- next to the uniform one the
        timed step uses.  Returns GB/s of algorithmic bytes (88 B / row POS + SHE, 40 B / row POS only).
        """
        torch = self.torch
        c = self.cat[0]
        rows, n = self.page_rows, self.pool * self.page_rows
        ipix = torch.empty(n, dtype=torch.int64, device="cuda")
        self.check(self.lib.hcu_ang2pix(self.h, 64, 1, c["lon"].data_ptr(), c["lat"].data_ptr(), n, ipix.data_ptr()))
        order = torch.argsort(ipix.view(self.pool, rows), dim=1)  # sorted inside every page
        lon = torch.gather(c["lon"].view(self.pool, rows), 1, order).contiguous()
        lat = torch.gather(c["lat"].view(self.pool, rows), 1, order).contiguous()
        w = torch.gather(c["w"].view(self.pool, rows), 1, order).contiguous()
        g = torch.gather(c["g"], 2, order[:, None, :].expand(-1, 2, -1)).contiguous() if self.cfg["she"] else None
        del ipix, order
        npix, nside = self.npix, self.cfg["nside"]
        pos, she = self.maps[0], (self.maps[self.nbins:self.nbins + 2] if self.cfg["she"] else None)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = None
        with torch.cuda.stream(self.stream):
            for _ in range(reps):
                ev0.record()
                for p in range(self.pool):
                    off = p * rows * 8
                    self.check(self.lib.hcu_map_page(self.h, nside, 0, lon.data_ptr() + off, lat.data_ptr() + off, w.data_ptr() + off,
                                                     g.data_ptr() + 2 * off if g is not None else None,
                                                     g.data_ptr() + 2 * off + rows * 8 if g is not None else None, rows,
                                                     pos.data_ptr(), she.data_ptr() if she is not None else None, npix, None))
                ev1.record()
                ev1.synchronize()
                ms = ev0.elapsed_time(ev1)
                best = ms if best is None else min(best, ms)
        return n * (88 if self.cfg["she"] else 40) / (best * 1e-3) / 1e9

    def stage_normalise(self, scale=True, shift=True):
        """scale: pos /= nbar, she /= wbar (linear: may precede the sum over ranks); shift: pos -= vis (once, after it)"""
        lib, h, npix = self.lib, self.h, self.npix
        for b in range(self.nbins):
            nbar, wbar = self.norm[b]
            p = self.maps[b].data_ptr()
            if scale:
                self.check(lib.hcu_divide(h, p, npix, nbar))          # pos /= nbar
                if self.cfg["she"]:
                    self.check(lib.hcu_divide(h, self.maps[self.nbins + 2 * b].data_ptr(), 2 * npix, wbar))
            if shift:
                self.check(lib.hcu_add_scalar(h, p, npix, -1.0))      # pos -= vis (full sky)

    def stage_transform(self, stats, spins=(0, 2)):
        lib, h, cfg = self.lib, self.h, self.cfg
        nb = self.nbins
        calls = [(0, 0, nb)] if 0 in spins else []
        if cfg["she"] and 2 in spins:
            calls.append((2, nb, 2 * nb))
        if self.dist is not None:
            # every rank transforms its ring block / its m; alm rows are zero for foreign m
            for spin, row0, n in calls:
                self.dist.map2alm(self.maps[row0:row0 + n], spin, self.alm[row0:row0 + n])
            return  # Legendre time and work are collected once per timed loop, see dist_stats()
        for spin, row0, n in calls:
            self.check(lib.hcu_map2alm(h, cfg["nside"], cfg["lmax"], spin, n, self.maps[row0].data_ptr(), self.npix,
                                       None, None, self.niter, None, self.alm[row0].data_ptr(), self.nalm))
            ms = self.ctx.sht_timing()
            rec, acc = self.ctx.sht_work()
            stats["fft_ms"] += ms[0] + ms[3]
            stats["leg_ana_ms"] += ms[1]
            stats["leg_syn_ms"] += ms[2]
            # counters cover every analysis pass of the call; the recursion is shared by the
            # components of a batch (12 spin-0 maps / 4 spin-2 fields), so weight the batches
            cap = int(self.lib.hcu_legendre_batch_size(spin))
            nbat = -(-n // cap)
            for i in range(nbat):
                nc = min(cap, n - i * cap)
                stats["leg_ana_flops"] += legendre_flops(spin, nc, rec / nbat, acc / nbat)

    def dist_flops_per_cell(self):
        """executed flops per accumulated cell-batch of one analysis pass, summed over the Legendre batches"""
        nb, tot, nbat = self.nbins, 0.0, 0
        for spin, n in ((0, nb), (2, 2 * nb if self.cfg["she"] else 0)):
            cap = int(self.lib.hcu_legendre_batch_size(spin))
            for i in range(-(-n // cap) if n else 0):
                nc = min(cap, n - i * cap)
                tot += legendre_flops(spin, nc, 1.0, 1.0)
                nbat += 1
        return tot, nbat

    def stage_cl(self):
        cfg = self.cfg
        # at N > 1 the alm are m-distributed: every rank sums its own m (m = rank mod world) only
        self.check(self.lib.hcu_alm2cl_mslice(self.h, self.ncomp, self.alm.data_ptr(), self.nalm, cfg["lmax"], self.ncomp,
                                              self.alm.data_ptr(), self.nalm, cfg["lmax"], cfg["lmax"], self.world, self.rank,
                                              self.cl.data_ptr()))
        if self.world > 1:  # partial sums over this rank's m
            import torch.distributed as dist

            dist.all_reduce(self.cl)

    def check(self, status):
        if status != 0:
            from heracles_b200 import _lib
            _lib.check(status)

    def step(self, stats):
        torch = self.torch
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        with torch.cuda.stream(self.stream):
            ev[0].record()
            self.stage_map()
            ev[1].record()
            if self.world > 1:
                # sum the partial maps over the ranks: POS now, SHE asynchronously (NCCL's own stream) while
                # the spin-0 transform runs; the linear normalisations come first, "- vis" after the sum
                import torch.distributed as dist

                from heracles_b200.dist import ReadyOnce

                self.stage_normalise(scale=True, shift=False)
                dist.all_reduce(self.maps[:self.nbins])
                work = dist.all_reduce(self.maps[self.nbins:], group=self.reduce_group, async_op=True) if self.cfg["she"] else None
                self.stage_normalise(scale=False, shift=True)
                ev[2].record()
                nb = self.nbins
                jobs = [(self.maps[:nb], 0, self.alm[:nb], None)]
                if self.cfg["she"]:
                    jobs.append((self.maps[nb:], 2, self.alm[nb:], None))
                self.dist.map2alm_jobs(jobs, {2: ReadyOnce(work)})
            else:
                self.stage_normalise()
                ev[2].record()
                self.stage_transform(stats)
            ev[3].record()
            self.stage_cl()
            ev[4].record()
        ev[4].synchronize()
        stats["map_ms"] += ev[0].elapsed_time(ev[1])
        stats["norm_ms"] += ev[1].elapsed_time(ev[2])
        stats["sht_ms"] += ev[2].elapsed_time(ev[3])
        stats["cl_ms"] += ev[3].elapsed_time(ev[4])
        return ev[0].elapsed_time(ev[4])

    # ---- end-to-end through the public API with pinned host pages ----
    def make_host_pool(self):
        """pinned host copies of the catalogue pages THIS rank maps (pool index -> row of the host arrays)"""
        torch = self.torch
        needed = sorted({p % self.pool for p in self.my_pages()})
        self.hidx = {pi: j for j, pi in enumerate(needed)}
        sel = torch.tensor(needed, device="cuda")
        self.hcat = []
        for b in range(self.nbins):
            c = self.cat[b]
            hc = {}
            for k, v in c.items():
                v = v.view(self.pool, -1, self.page_rows) if k == "g" else v.view(self.pool, self.page_rows)
                v = v.index_select(0, sel)
                t = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                t.copy_(v)
                hc[k] = t.numpy()
            self.hcat.append(hc)
        torch.cuda.synchronize()
        self.cat = None  # the device-resident catalogue is not needed by the end-to-end arm
        if self.dist is not None:
            self.dist.release()

    def e2e_step(self):
        hb = self.hb
        cfg, rows = self.cfg, self.page_rows

        class F:  # the two attributes heracles.mapping.transform reads from a Field
            def __init__(self, mapper, spin):
                self.mapper_or_error, self.spin = mapper, spin

        mapper = hb.CudaHealpixMapper(cfg["nside"], cfg["lmax"], deconvolve=False, niter=self.niter, sync=False, pixel_weights=None)
        fields = {"POS": F(mapper, 0), "SHE": F(mapper, 2)}
        t0 = time.perf_counter()
        maps = {}
        h2d = 0
        dist_mode = self.world > 1
        dp = None
        if dist_mode:
            # maps go bin by bin into device-resident stacks; the managed maps are freed right away
            from heracles_b200.dist import DistributedPipeline

            dp = DistributedPipeline(mapper, self.nbins, self.nbins if cfg["she"] else 0)
        else:
            vis = mapper.create()
            vis += 1.0
        # finished maps are transformed on a second library context while the next bins are still being mapped (the
        # mapping is PCIe-bound, the transform FP64-bound); HCU_BENCH_OVERLAP=0: map everything, then transform
        ov = None
        if not dist_mode and os.environ.get("HCU_BENCH_OVERLAP", "1") != "0":
            ov = hb.OverlappedTransform(mapper)
        pos = she = None
        stats = mapper.new_page_stats()  # one accumulator, cleared per bin: a cudaFree per bin would wait for the whole device
        for b in range(self.nbins):
            hc = self.hcat[b]
            if dist_mode and pos is not None:
                # the maps of the previous bin were copied into the device stacks: reuse the managed
                # buffers (allocating and freeing 4.8 GB of managed memory per bin is slow)
                pos *= 0.0
                if she is not None:
                    she *= 0.0
            else:
                pos = mapper.create(spin=0)
                she = mapper.create(2, spin=2) if cfg["she"] else None
            stats *= 0.0
            for p in self.my_pages():
                j = self.hidx[p % self.pool]
                lon, lat, w = hc["lon"][j], hc["lat"][j], hc["w"][j]
                if she is not None:
                    mapper.map_page(lon, lat, w, hc["g"][j][0], hc["g"][j][1], pos=pos, she=she, stats=stats)
                    h2d += 5 * rows * 8
                else:
                    mapper.map_page(lon, lat, w, pos=pos, stats=stats)
                    h2d += 3 * rows * 8
            if dist_mode:
                nbar, wbar = self.norm[b]  # the normalisation needs the sums over ALL ranks' pages
            else:
                # the Field layer's running means (fields.py:269-271, 283, 430-440), reduced on the device
                (ngal, wmean, _), _ = mapper.page_means(stats)
                nbar = ngal * wmean / self.npix
                wbar = ngal / (4 * math.pi) * wmean * (4 * math.pi / self.npix)
            pos /= nbar
            if she is not None:
                she /= wbar
            if dist_mode:
                dp.put(0, b, pos)
                if she is not None:
                    dp.put(2, b, she)
                continue
            pos -= vis
            if ov is not None:
                ov.submit(("POS", b), pos, spin=0)
                if she is not None:
                    ov.submit(("SHE", b), she, spin=2)
                continue
            maps["POS", b] = pos
            if she is not None:
                maps["SHE", b] = she
        bad = mapper.context.bad_rows()
        assert bad == 0
        t_map = time.perf_counter() - t0
        if dist_mode:
            # partial maps of this rank's pages -> sum over ranks -> (once) the visibility subtraction
            # -> ring-block / m-distributed transform -> Cl block on every rank
            def finish(stack, spin):
                if spin == 0:
                    stack.sub_(1.0)

            cl = dp.spectra(finish=finish)
            host = cl.cpu().numpy()
            d2h = host.nbytes
            checksum = checksum_matrix(host)
            ncl = self.ncomp * (self.ncomp + 1) // 2
            dt = time.perf_counter() - t0
            if self.rank == 0 and os.environ.get("HCU_BENCH_VERBOSE"):
                print(f"e2e rank 0: mapping {t_map:.3f} s, reduce + transform + Cl {dt - t_map:.3f} s", file=sys.stderr, flush=True)
            return dt, h2d, d2h, checksum, ncl
        alms = ov.finish() if ov is not None else hb.transform(fields, maps)
        for (k, i), a in alms.items():
            hb.update_metadata(a, spin=0 if k == "POS" else 2)
        cls = hb.angular_power_spectra(alms, debias=False)
        d2h = sum(np.asarray(c).nbytes for c in cls.values())
        checksum = checksum_results(cls)
        dt = time.perf_counter() - t0
        if os.environ.get("HCU_BENCH_VERBOSE"):
            print(f"e2e: mapping {t_map:.3f} s, transform + Cl {dt - t_map:.3f} s", file=sys.stderr, flush=True)
        return dt, h2d, d2h, checksum, len(cls)


def host_threads():
    """the cores this process may use -- NOT OMP_NUM_THREADS, which torchrun sets to 1"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def cpu_sample(cfg, niter, threads=None, sample_nside=512, sample_rows=2_000_000):
    """
    Time the oracle (CPU restatement of the reference path) on a BOUNDED sample of the configuration and scale
    to one full run with the stages' cost laws (rows; nside * lmax^2 per map and pass; nalm per spectrum).
    Returns (seconds_per_run, info).  The number is an extrapolation and says so (info["extrapolated"]).
    """
    import oracle

    oracle.build()
    oracle.set_num_threads(threads or host_threads())
    cores = oracle.num_threads()
    rng = np.random.default_rng(50)
    # 1. catalogue -> map at the real nside, 1 thread like the reference (healpy.py:157-160)
    n = sample_rows
    lon = rng.uniform(0, 360, n)
    lat = np.degrees(np.arcsin(rng.uniform(-1, 1, n)))
    w = rng.uniform(0.5, 1.5, n)
    g = rng.normal(0, 0.3, (2, n)) * w
    nside = cfg["nside"]
    pos = np.zeros(12 * nside * nside)
    t = time.perf_counter()
    oracle.map_values(nside, lon, lat, pos, w)
    t_pos = (time.perf_counter() - t) / n
    t_she = 0.0
    if cfg["she"]:
        she = np.zeros((2, 12 * nside * nside))
        t = time.perf_counter()
        oracle.map_values(nside, lon, lat, she, g)
        t_she = (time.perf_counter() - t) / n
        del she
    del pos
    rows_total = cfg["rows"] * cfg["nbins"]
    t_map = (t_pos + t_she) * rows_total
    # 2. transforms at reduced resolution, all threads, same niter; cost ~ nside * lmax^2 per map
    ns = min(nside, sample_nside)
    ls = 2 * ns
    m = rng.standard_normal((2, 12 * ns * ns))
    t = time.perf_counter()
    oracle.map2alm(ns, ls, m[:1], spin=0, niter=niter)
    t0 = time.perf_counter() - t
    scale = (nside / ns) * (cfg["lmax"] / ls) ** 2
    t_sht = t0 * scale * cfg["nbins"]
    if cfg["she"]:
        t = time.perf_counter()
        oracle.map2alm(ns, ls, m, spin=2, niter=niter)
        t_sht += (time.perf_counter() - t) * scale * cfg["nbins"]
    # 3. alm2cl, the reference's running-mean loop, one thread
    na = (ls + 1) * (ls + 2) // 2
    a = rng.standard_normal(na) + 1j * rng.standard_normal(na)
    t = time.perf_counter()
    for _ in range(3):
        oracle.alm2cl(a, a)
    t_cl1 = (time.perf_counter() - t) / 3
    ncomp = cfg["nbins"] * (3 if cfg["she"] else 1)
    nalm = (cfg["lmax"] + 1) * (cfg["lmax"] + 2) // 2
    t_cl = t_cl1 * (nalm / na) * ncomp * (ncomp + 1) / 2
    total = t_map + t_sht + t_cl
    info = {
        "value": total, "unit": UNIT, "cores": cores, "kind": "port",
        "extrapolated": ns != nside or n != rows_total,
        "scaling_law": "map: rows; map2alm: nside * lmax^2 per map and pass; alm2cl: nalm per spectrum",
        "sampled": {"rows": n, "nside": ns, "lmax": ls, "niter": niter},
        "stage_seconds_extrapolated": {"map": t_map, "sht": t_sht, "cl": t_cl},
        "sample": (f"oracle (CPU restatement of the healpy/ducc path): ang2pix+scatter on {n} rows at nside={nside} "
                   f"(1 thread, {t_pos * 1e9:.0f}+{t_she * 1e9:.0f} ns/row) x {rows_total:.3g} rows; map2alm spin0+spin2 "
                   f"niter={niter} at nside={ns} lmax={ls} ({cores} threads) scaled by nside*lmax^2 x {cfg['nbins']} bins; "
                   f"alm2cl at lmax={ls} scaled by nalm x {ncomp * (ncomp + 1) // 2} spectra"),
    }
    return total, info


def cpu_full(cfg, niter, threads=None):
    """
    The WHOLE configuration on the host cores, nothing extrapolated (affordable for C1 and C2 only): every page
    of every bin through the oracle's ang2pix + scatter (1 thread, like the reference), the Field normalisation,
    map2alm of every map (all threads) and every component spectrum.  Returns (seconds, info).
    """
    import oracle

    oracle.build()
    oracle.set_num_threads(threads or host_threads())
    cores = oracle.num_threads()
    nside, lmax, nb = cfg["nside"], cfg["lmax"], cfg["nbins"]
    npix = 12 * nside * nside
    pages = max(1, cfg["rows"] // PAGE_ROWS)
    rows = min(PAGE_ROWS, cfg["rows"])
    t_start = time.perf_counter()
    pos = np.zeros((nb, npix))
    she = np.zeros((nb, 2, npix)) if cfg["she"] else None
    t_gen = 0.0
    for b in range(nb):
        wsum = 0.0
        for p in range(pages):
            tg = time.perf_counter()
            rng = np.random.default_rng(50 + 1000 * b + p)
            lon = rng.uniform(0, 360, rows)
            lat = np.degrees(np.arcsin(rng.uniform(-1, 1, rows)))
            w = rng.uniform(0.5, 1.5, rows)
            g = rng.normal(0, 0.3, (2, rows)) * w if cfg["she"] else None
            t_gen += time.perf_counter() - tg  # synthetic page generation is not part of the path
            oracle.map_values(nside, lon, lat, pos[b], w)
            if she is not None:
                oracle.map_values(nside, lon, lat, she[b], g)
            wsum += w.sum()
        ngal = pages * rows
        wmean = wsum / ngal
        pos[b] /= ngal * wmean / npix
        pos[b] -= 1.0
        if she is not None:
            she[b] /= ngal / (4 * math.pi) * wmean * (4 * math.pi / npix)
    t_map = time.perf_counter() - t_start - t_gen
    t = time.perf_counter()
    alm = [oracle.map2alm(nside, lmax, pos, spin=0, niter=niter)]
    if she is not None:
        alm.append(oracle.map2alm(nside, lmax, she.reshape(2 * nb, npix), spin=2, niter=niter))
    alm = np.concatenate(alm)
    t_sht = time.perf_counter() - t
    t = time.perf_counter()
    ncomp = alm.shape[0]
    cl = np.zeros((ncomp, ncomp, lmax + 1))
    for i in range(ncomp):
        for j in range(i, ncomp):
            cl[i, j] = oracle.alm2cl(alm[i], alm[j])
    t_cl = time.perf_counter() - t
    total = t_map + t_sht + t_cl
    info = {"value": total, "unit": UNIT, "cores": cores, "kind": "port", "extrapolated": False,
            "stage_seconds": {"map": t_map, "sht": t_sht, "cl": t_cl}, "checksum": checksum_matrix(cl),
            "sample": f"the whole configuration ({cfg['rows'] * nb:.3g} rows, {ncomp} components, niter={niter}), measured"}
    return total, info


def run_reference(args, cfg_name):
    """`--impl reference`: the CPU arm.  healpy/ducc cannot be installed here, so this is the
    oracle port on all host cores.  Rank 0 only; the other ranks exit without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[cfg_name]
    threads = host_threads()
    times, info = [], None
    full = args.cpu_full and cfg_name in ("C1", "C2")
    for i in range(args.warmup + args.steps):
        if full:
            if i < args.warmup:
                continue  # a full CPU pass needs no warm-up repeats (minutes each)
            t, info = cpu_full(cfg, args.niter, threads)
        else:
            # warm-up steps on a cheap sample (page cache, thread pool, tables), timed steps on the real one
            warm = i < args.warmup
            t, info = cpu_sample(cfg, args.niter, threads, sample_nside=128 if warm else 512,
                                 sample_rows=200_000 if warm else 2_000_000)
        if i >= args.warmup:
            times.append(t)
    val = float(np.mean(times))
    info["value"] = val
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": val * 1e3, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(cfg_name, cfg, args.niter),
        "note": "healpy/ducc are not installable here: the CPU arm is the oracle port (kind 'port') on all host cores",
        "cpu_baseline": info,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def bench_config(cfg_name, cfg, niter):
    """the `config` object both arms print (same keys, so the driver can match them)"""
    ncomp = cfg["nbins"] * (3 if cfg["she"] else 1)
    pages = max(1, cfg["rows"] // PAGE_ROWS)
    return {
        "workload": f"{cfg_name}: {cfg['text']}", "nside": cfg["nside"], "lmax": cfg["lmax"],
        "fields": cfg["nbins"] * (2 if cfg["she"] else 1), "rows": cfg["rows"] * cfg["nbins"], "niter": niter,
        "spectra": ncomp * (ncomp + 1) // 2,
        "pages": f"{pages} pages of {min(PAGE_ROWS, cfg['rows'])} rows per bin cycling a pool of {min(POOL_PAGES, pages)} distinct pages",
        "l2": "inputs larger than L2 (catalogue and maps are GBs); no flush needed",
    }


def dist_parity_check(args, rank, world, torch, hb):
    """
    N > 1: before anything is timed, the ring-block / m-distributed pipeline over all ranks is compared with the
    single-GPU pipeline (run redundantly on every rank) on a small configuration.  Returns the largest
    |dC_l^{ij}| / sqrt(C_l^{ii} C_l^{jj}) over all component pairs and l; bench.py refuses to time a run whose
    ranks disagree with one GPU by more than 1e-10 (north_star's tolerance for Cl).
    """
    import torch.distributed as dist

    cfg = PARITY_CFG
    niter = min(args.niter, 1)
    out = []
    for r, w in ((rank, world), (0, 1)):
        pipe = Pipeline(cfg, niter, r, w, torch, hb)
        pipe.alloc_outputs()
        pipe.step(dict.fromkeys(["map_ms", "norm_ms", "sht_ms", "cl_ms", "fft_ms", "leg_ana_ms", "leg_syn_ms", "leg_ana_flops"], 0.0))
        torch.cuda.synchronize()
        out.append(pipe.cl.cpu().numpy())
        if pipe.dist is not None:
            pipe.dist.release()
        del pipe
    torch.cuda.empty_cache()
    cl_n, cl_1 = out
    auto = np.sqrt(np.abs(np.einsum("iil->il", cl_1)))
    norm = auto[:, None, :] * auto[None, :, :]
    iu = np.triu_indices(cl_1.shape[0])
    lmin = 2
    err = float((np.abs(cl_n - cl_1)[iu][:, lmin:] / norm[iu][:, lmin:]).max())
    t = torch.tensor([err], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    err = float(t.item())
    res = {"max_norm_err": err, "tolerance": 1e-10, "nside": cfg["nside"], "lmax": cfg["lmax"], "components": int(cl_1.shape[0]),
           "niter": niter, "checksum_n_gpu": checksum_matrix(cl_n), "checksum_1_gpu": checksum_matrix(cl_1)}
    if not err <= 1e-10:
        raise SystemExit(f"multi-GPU parity check failed: {json.dumps(res)}")
    return res


def run_c5(args, rank, world, local, torch, hb):
    """
    BASELINE.json config 5: nside 8192, lmax 16384, spin-2 map2alm of 40 (Q, U) maps -- SHT only.  The fields are
    independent, so they are sharded over the ranks (no collective on the data path); each rank transforms its
    fields in Legendre batches of `--c5-batch` fields drawn on the device.  One step = all 40 fields once.
    """
    cfg = C5
    nside, lmax = cfg["nside"], cfg["lmax"]
    npix, nalm = 12 * nside * nside, (lmax + 1) * (lmax + 2) // 2
    mine = list(range(rank, cfg["fields"], world))
    nb = max(1, min(args.c5_batch, 4))
    ctx = hb.get_context(local)
    ctx.set_timing(True)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    fp64_peak = ctx.fp64_peak()
    dev = torch.device("cuda")
    maps = torch.empty(2 * nb, npix, device=dev, dtype=torch.float64)
    alm = torch.empty(2 * nb, nalm, device=dev, dtype=torch.complex128)
    from heracles_b200 import _lib

    def one_pass(timed):
        tot = dict(ms=0.0, leg=0.0, fft=0.0, flops=0.0, check=0.0)
        for b0 in range(0, len(mine), nb):
            fields = mine[b0:b0 + nb]
            n = 2 * len(fields)
            for i, f in enumerate(fields):  # seeds 0..39; generation is not part of the transform
                g = torch.Generator(device=dev)
                g.manual_seed(f)
                maps[2 * i:2 * i + 2].normal_(generator=g)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record()
                _lib.check(ctx.lib.hcu_map2alm(ctx.handle, nside, lmax, 2, n, maps.data_ptr(), npix, None, None, args.niter,
                                               None, alm.data_ptr(), nalm))
                e1.record()
            e1.synchronize()
            ms = ctx.sht_timing()
            rec, acc = ctx.sht_work()
            tot["ms"] += e0.elapsed_time(e1)
            tot["leg"] += ms[1]
            tot["fft"] += ms[0] + ms[3]
            tot["flops"] += legendre_flops(2, n, rec, acc)
            # sum |a_lm|^2 without a temporary (the transform's workspaces leave little room next to 8 maps of 6.4 GB)
            tot["check"] += float(torch.linalg.vector_norm(torch.view_as_real(alm[:n])).item() ** 2)
        return tot

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_pass(False)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launch_count()
    barrier()
    acc = dict(ms=0.0, leg=0.0, fft=0.0, flops=0.0, check=0.0)
    for _ in range(args.steps):
        t = one_pass(True)
        for k in acc:
            acc[k] += t[k]
    barrier()
    clocks = sampler.stop()
    l1 = ctx.launch_count()
    ms = acc["ms"]
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        t = torch.tensor([acc["check"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        acc["check"] = float(t.item())
    ms_per_step = ms / args.steps
    nominal = nalm * (2 * nside) * (12 + 16 * nb) * (1 + 2 * args.niter) * (cfg["fields"] / nb)
    leg_tf = acc["flops"] / max(acc["leg"], 1e-9) * 1e3 / 1e12
    line = {
        "metric": "sht_seconds_per_run", "value": ms_per_step / 1e3, "unit": "s/run", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C5: {cfg['text']}", "nside": nside, "lmax": lmax, "fields": cfg["fields"], "niter": args.niter,
                   "fields_per_batch": nb, "sharding": "fields over ranks, no data-path collective",
                   "l2": "inputs larger than L2 (6.4 GB per map); no flush needed"},
        "sht_fp64_tflops": nominal / (ms_per_step * 1e-3) / 1e12,
        "sht_fp64_tflops_note": "nominal Legendre flops of SURVEY 8(d) (full triangle x all ring pairs) / wall time of the whole transform, all ranks",
        "stage_ms_per_step_rank0": {"legendre_ms": acc["leg"] / args.steps, "fft_ms": acc["fft"] / args.steps},
        "roofline": {"kernel": "legendre_analysis2_kernel", "bound": "tensor", "achieved": leg_tf, "peak": fp64_peak / 1e12,
                     "unit": "TFLOP/s", "frac": leg_tf / (fp64_peak / 1e12), "traffic": None,
                     "peak_source": "DFMA microkernel measured in this run (MEASURED_PEAKS.json has no FP64 figure)"},
        "gpu_launches": int((l1[0] - l0[0]) + (l1[1] - l0[1])), "clocks": clocks, "checksum": acc["check"],
        "e2e": None,
    }
    if rank == 0:
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="auto")
    ap.add_argument("--niter", type=int, default=3, help="map2alm Jacobi iterations (healpy default 3)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-full", action="store_true",
                    help="--impl reference: run the WHOLE configuration on the CPU (C1 / C2 only), nothing extrapolated")
    ap.add_argument("--c5-batch", type=int, default=4, help="--config C5: spin-2 fields per Legendre batch")
    args = ap.parse_args()

    if args.impl == "reference":
        name = args.config if args.config != "auto" else "C4"
        run_reference(args, name)
        return

    import torch

    import heracles_b200 as hb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.config == "C5":
        run_c5(args, rank, world, local, torch, hb)
        if world > 1:
            dist.destroy_process_group()
        return
    free, total_mem = torch.cuda.mem_get_info()
    cfg_name = pick_config(args.config, free)
    cfg = CONFIGS[cfg_name]
    dist_parity = dist_parity_check(args, rank, world, torch, hb) if world > 1 else None

    pipe = Pipeline(cfg, args.niter, rank, world, torch, hb)
    pipe.alloc_outputs()
    ctx = pipe.ctx
    fp64_peak = ctx.fp64_peak()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    stats = dict.fromkeys(["map_ms", "norm_ms", "sht_ms", "cl_ms", "fft_ms", "leg_ana_ms", "leg_syn_ms", "leg_ana_flops"], 0.0)
    for _ in range(args.warmup):
        pipe.step(dict(stats))
    barrier()
    if pipe.dist is not None:
        for lane in pipe.lanes:
            lane.k.timing = True
        pipe.dist.timing = True
        work0 = [sum(x) for x in zip(*(lane.k.ctx.sht_work() for lane in pipe.lanes))]
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launch_count()
    for k in stats:
        stats[k] = 0.0
    total_ms = 0.0
    barrier()
    for _ in range(args.steps):
        total_ms += pipe.step(stats)
    barrier()
    clocks = sampler.stop()
    l1 = ctx.launch_count()
    if pipe.dist is not None:
        # the staged path: Legendre analysis time from CUDA events around hcu_phase2alm, executed cells
        # from the kernels' work counters (equal shares for the batches of a pass)
        # with two lanes the Legendre kernels of one batch share the SMs with the FFT / exchange of another, so these
        # event-bracketed times overlap and their sum can exceed the step
        stats["leg_ana_ms"] = sum(lane.k.analysis_ms() for lane in pipe.lanes)
        for lane in pipe.lanes:
            lane.k.timing = False
        # per-rank stage times of the distributed transform (a2a includes waiting for the slowest rank)
        st = pipe.dist.stage_ms()
        pipe.dist.timing = False
        names = ["fft", "a2a", "leg_ana", "leg_syn", "ifft"]
        t = torch.tensor([st.get(k, 0.0) for k in names], device="cuda", dtype=torch.float64)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        dist_stage = {k: [round(float(a[i]) / args.steps, 1) for a in allt] for i, k in enumerate(names)}
        stats["fft_ms"] = st.get("fft", 0.0) + st.get("ifft", 0.0)
        stats["leg_syn_ms"] = st.get("leg_syn", 0.0)
        work1 = [sum(x) for x in zip(*(lane.k.ctx.sht_work() for lane in pipe.lanes))]
        per_cell, nbat = pipe.dist_flops_per_cell()
        rec, acc = (work1[0] - work0[0]) / nbat, (work1[1] - work0[1]) / nbat
        nb = cfg["nbins"]
        for spin, n in ((0, nb), (2, 2 * nb if cfg["she"] else 0)):
            cap = int(pipe.lib.hcu_legendre_batch_size(spin))
            for i in range(-(-n // cap) if n else 0):
                stats["leg_ana_flops"] += legendre_flops(spin, min(cap, n - i * cap), rec, acc)
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    checksum = checksum_matrix(pipe.cl.cpu().numpy())

    # roofline of the dominant kernel and of the HBM-bound scatter
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    leg_tflops = stats["leg_ana_flops"] / max(stats["leg_ana_ms"], 1e-9) * 1e3 / 1e12
    rows_step = len(pipe.my_pages()) * pipe.page_rows * cfg["nbins"]
    # fused page kernel: lon, lat, w (+ g1, g2) read once; one 16-byte read-modify-write per map row touched
    map_bytes = rows_step * ((24 + 16) + ((16 + 32) if cfg["she"] else 0))
    map_gbs = map_bytes * args.steps / max(stats["map_ms"], 1e-9) * 1e3 / 1e9
    # DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture) is read from
    # the committed summary profiles/r02_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep files
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        pass
    syn_flops = stats["leg_ana_flops"] * args.niter / (1 + args.niter) if args.niter else 0.0
    syn_tflops = syn_flops / max(stats["leg_syn_ms"], 1e-9) * 1e3 / 1e12 if args.niter else None
    roofline = {
        "kernel": "legendre_analysis_kernel (+ legendre_analysis2_kernel for 5..8 spin-0 maps)", "bound": "tensor",
        "pipe": "FP64 (DMMA mma.sync.m8n8k4.f64 + DFMA share it)",
        "achieved": leg_tflops, "peak": fp64_peak / 1e12,
        "unit": "TFLOP/s", "frac": leg_tflops / (fp64_peak / 1e12),
        "traffic": traffic.get(f"legendre_analysis_{cfg_name.lower()}"), "traffic_unit": "bytes per launch (4 spin-2 fields)",
        "traffic_source": "profiles/r02_traffic.json" if traffic else None,
        "peak_source": "DFMA microkernel measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
        "share_of_step": stats["leg_ana_ms"] / total_ms,
    }
    roofline_syn = None
    if syn_tflops is not None:
        roofline_syn = {
            "kernel": "legendre_synthesis_kernel", "bound": "tensor", "achieved": syn_tflops, "peak": fp64_peak / 1e12,
            "unit": "TFLOP/s", "frac": syn_tflops / (fp64_peak / 1e12),
            "flops": "the executed cells of the analysis passes (same geometry, same skipping) x SURVEY 8(d) per-cell figures",
            "traffic": traffic.get(f"legendre_synthesis_{cfg_name.lower()}"), "share_of_step": stats["leg_syn_ms"] / total_ms,
        }
    sorted_gbs = pipe.tile_sorted_rate() if pipe.cat is not None else None
    roofline_map = {
        "kernel": "map_page_kernel", "bound": "hbm", "achieved": map_gbs, "peak": hbm_peak, "unit": "GB/s",
        "frac": map_gbs / hbm_peak, "input": "uniform random positions (no locality inside a page)",
        "tile_sorted": {"achieved": sorted_gbs, "frac": sorted_gbs / hbm_peak if sorted_gbs else None,
                        "input": "the same rows ordered by their nside-64 NEST parent pixel inside every page (SURVEY 8(d))"},
        "traffic": traffic.get("map_page_1e6_rows"), "traffic_unit": "bytes per 1e6-row POS+SHE launch",
        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
        "share_of_step": stats["map_ms"] / total_ms,
    }

    # ring FFT stage (HBM-bound): per component and pass 8 npix bytes of map + 32 nrp (lmax + 1) bytes of phase rows;
    # (1 + niter) analysis and niter synthesis passes; at N > 1 every rank does its ring-pair block (1 / N of it)
    ncomp_maps = cfg["nbins"] * (3 if cfg["she"] else 1)
    fft_bytes = (8.0 * 12 * cfg["nside"] ** 2 + 32.0 * 2 * cfg["nside"] * (cfg["lmax"] + 1)) * ncomp_maps * (1 + 2 * args.niter) / world
    fft_gbs = fft_bytes / (stats["fft_ms"] / args.steps * 1e-3) / 1e9 if stats["fft_ms"] else None
    roofline_fft = {
        "kernel": "ring2_kernel (fused ring FFT, k_ringfft2.cu)", "bound": "hbm", "achieved": fft_gbs, "peak": hbm_peak, "unit": "GB/s",
        "frac": fft_gbs / hbm_peak if fft_gbs else None, "bytes": "algorithmic: 8 npix + 32 nrp (lmax + 1) per component and pass",
        "share_of_step": stats["fft_ms"] / total_ms,
    }

    line = {
        "metric": METRIC, "value": ms_per_step / 1e3, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(cfg_name, cfg, args.niter),
        "stage_ms_per_step": {k: stats[k] / args.steps for k in ("map_ms", "norm_ms", "sht_ms", "cl_ms", "fft_ms", "leg_ana_ms", "leg_syn_ms")},
        "sht_fp64_tflops_nominal": nominal_sht_flops(cfg, args.niter) / (stats["sht_ms"] / args.steps * 1e-3) / 1e12 if stats["sht_ms"] else None,
        "roofline": roofline, "roofline_synthesis": roofline_syn, "roofline_map_values": roofline_map,
        "roofline_ringfft": roofline_fft,
        "gpu_launches": int((l1[0] - l0[0]) + (l1[1] - l0[1])),
        "gpu_launches_detail": {"own_kernels": int(l1[0] - l0[0]), "cufft_execs": int(l1[1] - l0[1])},
        "clocks": clocks, "checksum": checksum,
    }
    if pipe.dist is not None:
        line["dist_stage_ms_per_rank"] = dist_stage
        line["dist_exchange"] = ("peer: ring-FFT / Legendre-synthesis kernels write the phase rows into the consuming rank's buffer over "
                                 "NVLink peer memory (CUDA IPC), one 1-element all-reduce per stage as the barrier"
                                 if pipe.lanes[0].peers is not None else "nccl: all_to_all_single")
        line["dist_parity"] = dist_parity

    # end to end through the plugin API, host pages
    if rank == 0:
        print("device-resident arm: " + json.dumps({k: line[k] for k in ("value", "n_gpus", "stage_ms_per_step")}), file=sys.stderr, flush=True)
    if not args.no_e2e:
        pipe.make_host_pool()
        del pipe.maps, pipe.alm, pipe.cl
        torch.cuda.empty_cache()
        ctx.trim()
        res = None
        pipe.e2e_step() if args.e2e_steps > 0 else None  # warm-up
        barrier()
        tt = 0.0
        for _ in range(max(1, args.e2e_steps)):
            res = pipe.e2e_step()
            tt += res[0]
        barrier()
        if world > 1:
            t = torch.tensor([tt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tt = float(t.item())
        api = ("CudaHealpixMapper.map_page(sync=False) + heracles_b200.dist.DistributedPipeline.spectra, pinned host pages"
               if world > 1 else
               "CudaHealpixMapper.map_page(sync=False) + heracles_b200." + ("transform" if os.environ.get("HCU_BENCH_OVERLAP", "1") == "0"
                                                                            else "OverlappedTransform (transforms of finished bins beside the mapping of the next)")
               + " + angular_power_spectra, pinned host pages")
        line["e2e"] = {"value": tt / max(1, args.e2e_steps), "unit": UNIT, "h2d_bytes_per_step": int(res[1]),
                       "d2h_bytes_per_step": int(res[2]), "spectra": res[4], "checksum": res[3], "api": api}

    if rank == 0 and not args.no_cpu:
        _, info = cpu_sample(cfg, args.niter)
        line["cpu_baseline"] = info
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
