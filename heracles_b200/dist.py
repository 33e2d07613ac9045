"""
Multi-GPU catalogue -> Cl path: one process per GPU, ``torch.distributed`` for the plumbing.

The reference is a single process (``heracles/mapping.py:130-174`` transforms the maps one
after the other on the host), so there is no reference interface to mirror here; this module
shards the SAME computation:

1. catalogue pages are split over the ranks (``pages[rank::world]``); every rank builds full
   partial maps with ``CudaHealpixMapper.map_values`` / ``hcu_map_values``;
2. ``reduce_maps`` sums the partial maps over the ranks; afterwards rank g works on ITS block
   of ring pairs only (a northern ring together with its southern mirror; blocks are balanced
   by pixel count);
3. ``DistributedTransform.map2alm``: local ring FFTs (``hcu_map2phase``) write the ring Fourier
   coefficients ordered by destination rank, one ``all_to_all_single`` makes them
   m-distributed (rank g owns m = g, g + world, ...: the Legendre work per m falls linearly in m,
   so this interleaving balances it to within 2 world / lmax), the Legendre analysis
   (``hcu_phase2alm``) runs per source block.  The Jacobi iterations of ``hp.map2alm``
   (healpy's default ``iter=3``) run the same way backwards: ``hcu_alm2phase`` per destination
   block, all-to-all, ``hcu_phase2map`` on the local rings, residual on the local pixels;
4. every rank reduces the spectra of its own m (``hcu_alm2cl`` on alm that are zero for the
   other m) and one ``all_reduce`` of ``nspec x (lmax + 1)`` doubles finishes the Cl.

The exchange logic is backend agnostic: ``StagedKernels`` binds the four C-ABI stage functions
for CUDA tensors; the CPU tests drive the same ``DistributedTransform`` over gloo with a
stand-in that evaluates the stages with the oracle.
"""

from __future__ import annotations

import ctypes
import hashlib
import os

import numpy as np

__all__ = ["ShardPlan", "StagedKernels", "Lane", "make_lanes", "DistributedTransform", "reduce_maps", "allreduce_cl", "as_torch", "DistributedPipeline"]


# ---------------------------------------------------------------------------------------
# the sharding plan (pure host logic)
# ---------------------------------------------------------------------------------------
class ShardPlan:
    """which ring pairs and which m every rank owns"""

    def __init__(self, nside: int, lmax: int, world: int, fft_cost=(1.52e-6, 0.041, 3.39e-6), align: int = 256):
        """fft_cost = (a, b, c): relative ring-FFT cost of a ring pair, a * M log2 M + b * nside / 4096 for a polar-cap
        pair (M = power-of-two Bluestein length of its 4 sub-FFTs; the second term is the emission of its lmax + 1 phase
        rows, the same for every ring) and c * pixels for a belt pair.  Fitted to the launch times of the fused ring-FFT
        kernels at nside 4096 (k_ringfft2.cu; profiles/r02_fft2_launches.txt): Bluestein rings of length 8192 cost 1.8 x a
        belt pair, the rings below it 0.8 x on average.  The ring-pair blocks are balanced by this cost;
        fft_cost=None balances by pixel count.  The boundaries are then snapped to multiples of `align` ring pairs, the
        ring-pair group one Legendre CTA works on: a block of 1454 ring pairs costs six CTAs per m, the last one two thirds
        full -- measured at C4 on 8 GPUs, unaligned blocks made 34 instead of 32 groups and the Legendre stage 5.5 % slower,
        far more than the ring-FFT imbalance the snapping introduces (the FFT stage is 6 % of the Legendre time)."""
        if world < 1:
            raise ValueError("world must be >= 1")
        nrp = 2 * nside
        if world > nrp:
            raise ValueError("more ranks than ring pairs")
        self.nside, self.lmax, self.world = int(nside), int(lmax), int(world)
        self.npix = 12 * nside * nside
        self.nalm = (lmax + 1) * (lmax + 2) // 2
        self.nrp = nrp
        # pixels per ring pair: caps 2 x 4i, belt 2 x 4 nside, the equator ring counts once
        i = np.arange(1, nrp + 1, dtype=np.int64)
        if fft_cost is None:
            npair = np.where(i < nside, 8.0 * i, 8.0 * nside)
            npair[-1] = 4 * nside
        else:
            a, b, c = fft_cost
            m = 2.0 ** np.ceil(np.log2(np.maximum(2 * i - 1, 2)))
            npair = np.where(i < nside, a * m * np.log2(m) + b * nside / 4096.0, c * 8.0 * nside)
            npair[-1] = c * 4.0 * nside
        cum = np.concatenate([[0], np.cumsum(npair)])
        bounds = [0]
        for g in range(1, world):
            target = cum[-1] * g / world
            b = int(np.searchsorted(cum, target, side="left"))
            b = max(b, bounds[-1] + 1)           # every block has at least one ring pair
            b = min(b, nrp - (world - g))        # ... also the remaining ones
            bounds.append(b)
        bounds.append(nrp)
        if align and align > 1 and nrp >= 2 * align * world:
            snapped = [0]
            for g in range(1, world):
                b = int(round(bounds[g] / align)) * align
                b = max(b, snapped[-1] + align)
                b = min(b, nrp - (world - g) * align)
                snapped.append(b)
            snapped.append(nrp)
            bounds = snapped
        self.rp_bounds = bounds
        # m owned by rank g: g, g + world, ...; the concatenation ordered by owner is the row order
        # of every exchanged phase array
        self.mlists = [np.arange(g, lmax + 1, world, dtype=np.int32) for g in range(world)]
        self.m_all = np.concatenate(self.mlists).astype(np.int32)
        self.m_off = np.concatenate([[0], np.cumsum([len(m) for m in self.mlists])]).astype(np.int64)
        self.mpos = np.empty(lmax + 1, dtype=np.int32)
        self.mpos[self.m_all] = np.arange(lmax + 1, dtype=np.int32)

    # -- geometry -------------------------------------------------------------------------
    def ring_start(self, i: int) -> int:
        """first pixel of northern-hemisphere / belt ring i (1 <= i <= 2 nside + 1 as an end marker)"""
        ns = self.nside
        if i <= ns:
            return 2 * i * (i - 1)
        return 2 * ns * (ns - 1) + (i - ns) * 4 * ns

    def rp_range(self, g: int) -> tuple[int, int]:
        return self.rp_bounds[g], self.rp_bounds[g + 1]

    def nrp_of(self, g: int) -> int:
        return self.rp_bounds[g + 1] - self.rp_bounds[g]

    def pixel_ranges(self, g: int) -> list[tuple[int, int]]:
        """pixel index ranges (RING order) of the ring pairs of rank g: north run, south run"""
        lo, hi = self.rp_range(g)
        a, b = lo + 1, hi  # northern ring numbers a..b inclusive
        n0, n1 = self.ring_start(a), self.ring_start(b + 1)
        if b == 2 * self.nside:  # the block ends with the equator ring, which is its own mirror
            return [(n0, self.npix - n0)]
        return [(n0, n1), (self.npix - n1, self.npix - n0)]

    def owner_of_m(self, m: int) -> int:
        return m % self.world


# ---------------------------------------------------------------------------------------
# CUDA stage kernels through the C ABI
# ---------------------------------------------------------------------------------------
class StagedKernels:
    """hcu_map2phase / hcu_phase2alm / hcu_alm2phase / hcu_phase2map on CUDA torch tensors"""

    def __init__(self, ctx, nside: int, lmax: int):
        self.ctx, self.nside, self.lmax = ctx, int(nside), int(lmax)
        self.npix = 12 * nside * nside
        self.nalm = (lmax + 1) * (lmax + 2) // 2
        self.timing = False      # record CUDA events around the Legendre analysis launches
        self._events = []

    def analysis_ms(self) -> float:
        """device milliseconds of the hcu_phase2alm launches since the last call (synchronises)"""
        import torch

        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in self._events)
        self._events = []
        return ms

    def batch_size(self, spin: int) -> int:
        return int(self.ctx.lib.hcu_legendre_batch_size(int(spin)))

    def _check(self, status):
        if status != 0:
            from . import _lib

            _lib.check(status)

    @staticmethod
    def _p(t):
        return None if t is None else t.data_ptr()

    def map2phase(self, maps, rp_lo, rp_hi, mlist, phase):
        nb = maps.shape[0]
        self._check(self.ctx.lib.hcu_map2phase(self.ctx.handle, self.nside, self.lmax, nb, maps.data_ptr(), maps.stride(0),
                                               None, rp_lo, rp_hi, self._p(mlist), mlist.numel(), phase.data_ptr()))

    def phase2alm(self, phase, spin, nb, mlist, rp_lo, rp_hi, alm):
        if self.timing:
            import torch

            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self._check(self.ctx.lib.hcu_phase2alm(self.ctx.handle, self.nside, self.lmax, spin, nb, phase.data_ptr(),
                                               self._p(mlist), mlist.numel(), rp_lo, rp_hi, None, alm.data_ptr(), alm.stride(0)))
        if self.timing:
            e1.record()
            self._events.append((e0, e1))

    def phase2alm_blocks(self, phase, spin, nb, mlist, rp_bounds, alm):
        """all source blocks of an exchanged phase array in one launch"""
        import ctypes

        if self.timing:
            import torch

            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        rb = (ctypes.c_int64 * len(rp_bounds))(*rp_bounds)
        self._check(self.ctx.lib.hcu_phase2alm_blocks(self.ctx.handle, self.nside, self.lmax, spin, nb, phase.data_ptr(),
                                                      self._p(mlist), mlist.numel(), len(rp_bounds) - 1, rb, None,
                                                      alm.data_ptr(), alm.stride(0)))
        if self.timing:
            e1.record()
            self._events.append((e0, e1))

    def alm2phase_blocks(self, alm, spin, nb, mlist, rp_bounds, phase):
        import ctypes

        rb = (ctypes.c_int64 * len(rp_bounds))(*rp_bounds)
        self._check(self.ctx.lib.hcu_alm2phase_blocks(self.ctx.handle, self.nside, self.lmax, spin, nb, alm.data_ptr(),
                                                      alm.stride(0), self._p(mlist), mlist.numel(), len(rp_bounds) - 1, rb,
                                                      phase.data_ptr()))

    def alm2phase(self, alm, spin, nb, mlist, rp_lo, rp_hi, phase):
        self._check(self.ctx.lib.hcu_alm2phase(self.ctx.handle, self.nside, self.lmax, spin, nb, alm.data_ptr(), alm.stride(0),
                                               self._p(mlist), mlist.numel(), rp_lo, rp_hi, phase.data_ptr()))

    def phase2map(self, phase, nb, mpos, rp_lo, rp_hi, maps):
        self._check(self.ctx.lib.hcu_phase2map(self.ctx.handle, self.nside, self.lmax, nb, phase.data_ptr(), self._p(mpos),
                                               rp_lo, rp_hi, maps.data_ptr(), maps.stride(0)))

    def map2phase_peers(self, maps, rp_lo, rp_hi, mlist, row_start, dest_ptrs):
        """hcu_map2phase whose rows [row_start[d], row_start[d+1]) are written to the (peer) address dest_ptrs[d]"""
        import ctypes

        nb, nd = maps.shape[0], len(dest_ptrs)
        rs = (ctypes.c_int32 * (nd + 1))(*row_start)
        bp = (ctypes.c_void_p * nd)(*dest_ptrs)
        self._check(self.ctx.lib.hcu_map2phase_peers(self.ctx.handle, self.nside, self.lmax, nb, maps.data_ptr(), maps.stride(0),
                                                     None, rp_lo, rp_hi, self._p(mlist), mlist.numel(), nd, rs, bp))

    def alm2phase_peers(self, alm, spin, nb, mlist, rp_bounds, block_ptrs):
        """hcu_alm2phase_blocks whose block b is written to the (peer) address block_ptrs[b]"""
        import ctypes

        rb = (ctypes.c_int64 * len(rp_bounds))(*rp_bounds)
        bp = (ctypes.c_void_p * len(block_ptrs))(*block_ptrs)
        self._check(self.ctx.lib.hcu_alm2phase_peers(self.ctx.handle, self.nside, self.lmax, spin, nb, alm.data_ptr(),
                                                     alm.stride(0), self._p(mlist), mlist.numel(), len(rp_bounds) - 1, rb, bp))

    def sync_streams(self):
        """make the library's stream the caller's current torch stream"""
        import torch

        # torch's default stream has handle 0, which hcu_set_stream reads as "the context's own
        # stream": name the legacy default stream explicitly (cudaStreamLegacy = 1)
        self.ctx.set_stream(torch.cuda.current_stream().cuda_stream or 1)


# ---------------------------------------------------------------------------------------
# exchange through peer memory
# ---------------------------------------------------------------------------------------
class PeerExchange:
    """
    The ring-block <-> m-distributed exchange WITHOUT a collective: every rank owns a double-buffered phase array in
    device memory that all other ranks of the node have opened over CUDA IPC, and the PRODUCING kernels (the ring FFT's
    row emission, the Legendre synthesis' flush) write every row straight into the buffer of the rank that consumes it
    -- NVLink peer stores instead of a send buffer + ``all_to_all_single`` + a receive buffer.

    (The analysis direction stages the rows of the OTHER ranks locally and pushes them as one contiguous copy-engine
    copy per destination, ``push``: scattered 32-byte stores over NVLink reach a fraction of the link rate; the
    synthesis direction's Legendre kernel is compute-bound and stores remotely for free.)

    Ordering: stage n uses half n % 2 of every buffer and ends with one (stream-ordered, 1-element) all-reduce.  A rank
    leaves that barrier only after every rank's producer kernel of stage n has completed, so the consumer may read; and
    a producer of stage n + 2 starts only after the barrier of stage n + 1, which every rank enters after ITS consumer
    of stage n -- so no half is overwritten while it is still being read.
    """

    PER_MAX = 48  # doubles per (m, ring pair): 12 components x 4

    def __init__(self, ctx, plan, rank, group=None, device=None):
        """Collective: every rank of `group` must call it.  No rank is left waiting in a collective the others never
        reach when the allocation or cudaIpcOpenMemHandle fails somewhere: every step that can fail locally is followed
        by an exchange in which all ranks take part, and a failure anywhere raises on every rank."""
        import torch
        import torch.distributed as dist

        from . import _lib

        self.ctx, self.plan, self.rank, self.group = ctx, plan, int(rank), group
        self.device = device
        self.base, self.ptr, self._opened, self.stage = 0, [], [], 0
        W = plan.world
        lo, hi = plan.rp_range(rank)
        n_ana = len(plan.mlists[rank]) * plan.nrp * self.PER_MAX      # what the Legendre analysis of this rank reads
        n_syn = (plan.lmax + 1) * (hi - lo) * self.PER_MAX             # what the inverse ring FFTs of this rank read
        self.half_elems = max(n_ana, n_syn, 1)
        mine, err = None, None
        try:
            self.base = ctx.malloc_device(2 * self.half_elems * 8)
            handle = (ctypes.c_ubyte * 64)()
            _lib.check(ctx.lib.hcu_ipc_export(ctx.handle, ctypes.c_void_p(self.base), handle))
            mine = (bytes(handle), self.half_elems)
        except Exception as e:  # noqa: BLE001 - reported after the exchange
            err = e
        everyone = [None] * W
        dist.all_gather_object(everyone, mine, group=group)
        if err is None and any(x is None for x in everyone):
            err = RuntimeError("another rank could not export its exchange buffer")
        if err is None:
            try:
                for d, (h, n) in enumerate(everyone):
                    if d == rank:
                        b = self.base
                    else:
                        p = ctypes.c_void_p()
                        buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                        _lib.check(ctx.lib.hcu_ipc_open(ctx.handle, buf, ctypes.byref(p)))
                        b = p.value
                        self._opened.append(b)
                    self.ptr.append((b, b + 8 * n))
            except Exception as e:  # noqa: BLE001
                err = e
        # every rank has opened every buffer before anybody writes -- or nobody uses the exchange
        self.flag = torch.tensor([0.0 if err is None else 1.0], dtype=torch.float32, device=device)
        dist.all_reduce(self.flag, group=group)
        failed = float(self.flag.item()) != 0.0
        self.flag.zero_()
        if failed:
            self.close()
            raise RuntimeError(f"peer-memory exchange unavailable: {err or 'a peer could not open the buffers'}")

    def next_half(self) -> int:
        h = self.stage & 1
        self.stage += 1
        return h

    def barrier(self):
        """stream-ordered: everything queued on the current stream so far, on every rank, precedes what follows"""
        import torch.distributed as dist

        dist.all_reduce(self.flag, group=self.group)

    def push(self, src, chunks):
        """copy-engine pushes: chunks = [(rank, offset in src, elements, destination address)]; the copies run on a few
        side streams (ordered after what the current stream has queued) and the current stream waits for them"""
        import torch

        if not chunks:
            return
        cur = torch.cuda.current_stream()
        nstreams = int(os.environ.get("HCU_PUSH_STREAMS", "4"))
        if not hasattr(self, "_copy_streams") or len(self._copy_streams) != nstreams:
            self._copy_streams = [torch.cuda.Stream(device=self.device) for _ in range(nstreams)]
        ev = torch.cuda.Event()
        ev.record(cur)
        # one copy engine does not fill the NVLink ports: every destination's chunk is cut so that about as many copies
        # as there are side streams are in flight
        parts = max(1, nstreams // len(chunks))
        used, i = [], 0
        for d, off, n, addr in chunks:
            step = -(-n // parts)
            for o in range(0, n, step):
                m = min(step, n - o)
                st = self._copy_streams[i % nstreams]
                i += 1
                if st not in used:
                    st.wait_event(ev)
                    used.append(st)
                dst = torch.as_tensor(_CudaView(addr + 8 * o, (m,), "<f8"), device=self.device)
                with torch.cuda.stream(st):
                    dst.copy_(src[off + o:off + o + m], non_blocking=True)
        for st in used:
            done = torch.cuda.Event()
            done.record(st)
            cur.wait_event(done)

    def local(self, half: int, n: int):
        """torch view of the first n doubles of one half of this rank's buffer"""
        import torch

        return torch.as_tensor(_CudaView(self.ptr[self.rank][half], (n,), "<f8"), device=self.device)

    def close(self):
        if self.base:
            self.ctx.synchronize()
            for p in self._opened:
                self.ctx.lib.hcu_ipc_close(self.ctx.handle, ctypes.c_void_p(p))
            self._opened = []
            self.ctx.free(self.base)
            self.base = 0


# ---------------------------------------------------------------------------------------
# the distributed transform
# ---------------------------------------------------------------------------------------
class Lane:
    """
    One independent instruction stream for whole Legendre batches: its own stage kernels (i.e. its own library context
    with its workspaces and cuFFT plans), process group (collectives of ONE communicator run in issue order), CUDA
    stream and exchange buffers.  With two lanes the ring FFTs and the all-to-all of one batch run while the other
    batch is in its (FP64-bound) Legendre kernels.
    """

    def __init__(self, kernels, group=None, stream=None, peers=None):
        self.k, self.group, self.stream = kernels, group, stream
        self.peers = peers  # PeerExchange: the producing kernels write into the consumers' buffers (no all-to-all)
        self.ws = {}

    def context(self):
        import contextlib

        if self.stream is None:
            return contextlib.nullcontext()
        import torch

        return torch.cuda.stream(self.stream)


class DistributedTransform:
    """
    ``hp.map2alm`` (``heracles/healpy.py:183-189``) over the ranks of a process group.

    maps : tensor ``[ncomp, npix]`` float64 in which at least the pixels of this rank's ring
        block hold the (rank-summed) map; spin 2: rows are (Q, U) pairs.
    alm  : tensor ``[ncomp, nalm]`` complex128; on return the entries whose m this rank owns
        hold the result, every other entry is zero (so a SUM all-reduce gathers them).
    lanes : optional list of :class:`Lane`; the Legendre batches of a call are dealt out to the lanes and run
        concurrently (default: one lane on the caller's stream).
    """

    def __init__(self, kernels, plan: ShardPlan, rank: int, group=None, niter: int = 3, device=None, lanes=None):
        import torch

        self.k, self.plan, self.rank, self.group, self.niter = kernels, plan, int(rank), group, int(niter)
        self.device = device if device is not None else torch.device("cpu")
        dev = self.device
        self.mlist_me = torch.from_numpy(plan.mlists[rank].copy()).to(dev)
        self.m_all = torch.from_numpy(plan.m_all.copy()).to(dev)
        self.mpos = torch.from_numpy(plan.mpos.copy()).to(dev)
        self.lanes = list(lanes) if lanes else [Lane(kernels, group, None)]
        self._ws = {}
        self.exchanged_bytes = 0
        self.timing = False   # CUDA events around the stages (device tensors only)
        self._ev = {}

    def _mark(self, name):
        """context manager: accumulate device time of a stage under `name` when self.timing"""
        import contextlib

        if not self.timing:
            return contextlib.nullcontext()
        import torch

        outer = self

        class _T:
            def __enter__(self):
                self.a = torch.cuda.Event(enable_timing=True)
                self.b = torch.cuda.Event(enable_timing=True)
                self.a.record()

            def __exit__(self, *exc):
                self.b.record()
                outer._ev.setdefault(name, []).append((self.a, self.b))

        return _T()

    def stage_ms(self):
        """device milliseconds per stage since the last call (synchronises); with two lanes the stages of
        different batches overlap, so the sum exceeds the elapsed time"""
        import torch

        torch.cuda.synchronize()
        out = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in self._ev.items()}
        self._ev = {}
        return out

    # -- helpers ------------------------------------------------------------------------------
    def _buf(self, lane, name, n):
        import torch

        t = lane.ws.get(name)
        if t is None or t.numel() < n:
            t = torch.empty(n, dtype=torch.float64, device=self.device)
            lane.ws[name] = t
        return t[:n]

    def release(self):
        self._ws.clear()
        for lane in self.lanes:
            lane.ws.clear()  # (the peer-exchange buffers are cached per process: close_peer_exchanges())

    def _all_to_all(self, lane, out, inp, out_splits, in_splits):
        import torch.distributed as dist

        if self.plan.world == 1:
            out.copy_(inp)
            return
        dist.all_to_all_single(out, inp, output_split_sizes=out_splits, input_split_sizes=in_splits, group=lane.group)
        self.exchanged_bytes += 8 * (sum(in_splits) - in_splits[self.rank])

    # -- one analysis pass over a batch: alm += A(maps); yields between its stages ---------------------
    def _analysis(self, lane, maps, spin, alm):
        plan, g, W = self.plan, self.rank, self.plan.world
        k = lane.k
        nb = maps.shape[0]
        lo, hi = plan.rp_range(g)
        nrp_me, nm_me = hi - lo, len(plan.mlists[g])
        per = nb * 4
        in_splits = [len(plan.mlists[d]) * nrp_me * per for d in range(W)]
        out_splits = [nm_me * plan.nrp_of(s) * per for s in range(W)]
        if lane.peers is not None:
            # Rows of destination d belong in MY block of d's blocked phase array (behind the blocks of the ranks before
            # me).  My own rows are written there by the kernel; the others are emitted into a local staging array
            # (scattered 32-byte stores over NVLink run at a fraction of the link rate -- measured, profiles/) and
            # pushed as ONE contiguous copy per destination by the copy engines: no SM, no collective.
            ex = lane.peers
            half = ex.next_half()
            before = sum(plan.nrp_of(s) for s in range(g))
            row_start = [0]
            for d in range(W):
                row_start.append(row_start[-1] + len(plan.mlists[d]))
            send = self._buf(lane, "send", (plan.lmax + 1) * nrp_me * per)
            target = [ex.ptr[d][half] + 8 * len(plan.mlists[d]) * before * per for d in range(W)]
            dest = [target[d] if d == g else send.data_ptr() + 8 * row_start[d] * nrp_me * per for d in range(W)]
            with self._mark("fft"):
                k.map2phase_peers(maps, lo, hi, self.m_all, row_start, dest)
            yield
            with self._mark("a2a"):
                ex.push(send, [(d, row_start[d] * nrp_me * per, in_splits[d], target[d]) for d in range(W) if d != g and in_splits[d]])
                ex.barrier()
            self.exchanged_bytes += 8 * (sum(in_splits) - in_splits[g])
            recv = ex.local(half, sum(out_splits))
            yield
        else:
            send = self._buf(lane, "send", (plan.lmax + 1) * nrp_me * per)
            with self._mark("fft"):
                k.map2phase(maps, lo, hi, self.m_all, send)
            yield
            recv = self._buf(lane, "recv", sum(out_splits))
            with self._mark("a2a"):
                self._all_to_all(lane, recv, send, out_splits, in_splits)
            yield
        off = 0
        with self._mark("leg_ana"):
            if nm_me and hasattr(k, "phase2alm_blocks") and W <= 16:
                k.phase2alm_blocks(recv, spin, nb, self.mlist_me, list(plan.rp_bounds), alm)
            else:
                for s in range(W):
                    slo, shi = plan.rp_range(s)
                    if nm_me and out_splits[s]:
                        k.phase2alm(recv[off:off + out_splits[s]], spin, nb, self.mlist_me, slo, shi, alm)
                    off += out_splits[s]
        yield

    # -- one synthesis pass over a batch: maps (local rings) = S(alm) ------------------------------
    def _synthesis(self, lane, alm, spin, maps):
        plan, g, W = self.plan, self.rank, self.plan.world
        k = lane.k
        nb = alm.shape[0]
        lo, hi = plan.rp_range(g)
        nrp_me, nm_me = hi - lo, len(plan.mlists[g])
        per = nb * 4
        in_splits = [nm_me * plan.nrp_of(d) * per for d in range(W)]
        out_splits = [len(plan.mlists[s]) * nrp_me * per for s in range(W)]
        if lane.peers is not None:
            # my rows of destination d's phase array start behind the rows of the ranks before me
            ex = lane.peers
            half = ex.next_half()
            rows_before = sum(len(plan.mlists[s]) for s in range(g))
            blocks = [ex.ptr[d][half] + 8 * rows_before * plan.nrp_of(d) * per for d in range(W)]
            with self._mark("leg_syn"):
                if nm_me:
                    k.alm2phase_peers(alm, spin, nb, self.mlist_me, list(plan.rp_bounds), blocks)
            yield
            with self._mark("a2a"):
                ex.barrier()
            self.exchanged_bytes += 8 * (sum(in_splits) - in_splits[g])
            recv = ex.local(half, sum(out_splits))
            yield
            with self._mark("ifft"):
                k.phase2map(recv, nb, self.mpos, lo, hi, maps)
            yield
            return
        send = self._buf(lane, "send", sum(in_splits))
        off = 0
        with self._mark("leg_syn"):
            if nm_me and hasattr(k, "alm2phase_blocks") and W <= 16:
                k.alm2phase_blocks(alm, spin, nb, self.mlist_me, list(plan.rp_bounds), send)
            else:
                for d in range(W):
                    dlo, dhi = plan.rp_range(d)
                    if nm_me and in_splits[d]:
                        k.alm2phase(alm, spin, nb, self.mlist_me, dlo, dhi, send[off:off + in_splits[d]])
                    off += in_splits[d]
        yield
        recv = self._buf(lane, "recv", sum(out_splits))
        with self._mark("a2a"):
            self._all_to_all(lane, recv, send, out_splits, in_splits)
        yield
        # rows of recv are ordered by source rank = the order of plan.m_all
        with self._mark("ifft"):
            k.phase2map(recv, nb, self.mpos, lo, hi, maps)
        yield

    def _batch(self, lane, mb, spin, ab, ready=None):
        """all passes of one Legendre batch; a generator that yields after every stage it has queued"""
        import torch

        plan = self.plan
        ranges = plan.pixel_ranges(self.rank)
        if ready is not None:
            ready()  # e.g. the (asynchronous) reduction of these maps over the ranks: this lane's stream waits for it
        ab.zero_()
        yield from self._analysis(lane, mb, spin, ab)
        if self.niter > 0:
            n = mb.shape[0]
            resid = self._buf(lane, "resid", n * plan.npix).view(n, plan.npix)
            for _ in range(self.niter):
                yield from self._synthesis(lane, ab, spin, resid)
                for a, b in ranges:  # residual on the pixels of this rank's rings
                    torch.sub(mb[:, a:b], resid[:, a:b], out=resid[:, a:b])
                yield from self._analysis(lane, resid, spin, ab)

    # -- public -----------------------------------------------------------------------------------
    def map2alm(self, maps, spin: int, alm, fl=None):
        return self.map2alm_jobs([(maps, spin, alm, fl)])[0]

    def map2alm_jobs(self, jobs, ready=None):
        """
        Several transforms at once: ``jobs`` is a list of ``(maps, spin, alm, fl)``.  Their Legendre batches are dealt
        out to the lanes round robin and the lanes' stage sequences are queued interleaved, so that with two lanes the
        FFT / exchange stages of one batch overlap the Legendre kernels of another.  ``ready``: optional dict
        ``spin -> callable`` that makes the CURRENT stream wait until the maps of that spin may be read.
        """
        import torch

        plan = self.plan
        per_lane = [[] for _ in self.lanes]
        nbatch = 0
        for maps, spin, alm, fl in jobs:
            if spin not in (0, 2):
                msg = f"spin-{spin} maps not yet supported"
                raise NotImplementedError(msg)
            ncomp = maps.shape[0]
            if maps.shape[-1] != plan.npix or alm.shape != (ncomp, plan.nalm):
                raise ValueError("maps / alm do not match the plan")
            if spin == 2 and ncomp % 2:
                raise ValueError("spin-2 input needs (Q, U) pairs")
            cap = self.lanes[0].k.batch_size(spin)
            for c0 in range(0, ncomp, cap):
                c1 = min(ncomp, c0 + cap)
                per_lane[nbatch % len(self.lanes)].append((maps[c0:c1], spin, alm[c0:c1]))
                nbatch += 1
        use_streams = any(lane.stream is not None for lane in self.lanes)
        start = None
        if use_streams:
            start = torch.cuda.Event()
            start.record()

        def lane_steps(lane, batches):
            if start is not None and lane.stream is not None:
                lane.stream.wait_event(start)  # everything queued so far (maps, normalisation) is visible to the lane
            for mb, spin, ab in batches:
                yield from self._batch(lane, mb, spin, ab, None if ready is None else ready.get(spin))

        active = [(lane, lane_steps(lane, b)) for lane, b in zip(self.lanes, per_lane) if b]
        while active:
            for item in list(active):
                lane, it = item
                with lane.context():
                    try:
                        next(it)
                    except StopIteration:
                        active.remove(item)
        if use_streams:
            cur = torch.cuda.current_stream()
            for lane in self.lanes:
                if lane.stream is not None:
                    ev = torch.cuda.Event()
                    ev.record(lane.stream)
                    cur.wait_event(ev)
        for maps, spin, alm, fl in jobs:
            if fl is not None:
                alm.mul_(self._fl_full(fl))
        return [alm for _, _, alm, _ in jobs]

    def _fl_full(self, fl):
        """fl[l] expanded to the alm layout (complex128 [nalm])"""
        import torch

        lmax = self.plan.lmax
        # keyed by CONTENT: callers pass temporaries (mapper._fl(spin)), whose id() is recycled
        # as soon as the previous filter is freed -- the spin-0 window must never serve spin 2
        fl = np.ascontiguousarray(fl, dtype=np.float64)
        key = ("fl", hashlib.sha1(fl.tobytes()).hexdigest())
        t = self._ws.get(key)
        if t is None:
            full = np.concatenate([fl[m:lmax + 1] for m in range(lmax + 1)])
            t = torch.from_numpy(full).to(self.device).to(torch.complex128)
            self._ws[key] = t
        return t


def exchange_mode() -> str:
    """HCU_DIST_EXCHANGE = peer (default: producer kernels write into the consumers' buffers over NVLink peer memory)
    or nccl (send buffer + all_to_all_single + receive buffer)"""
    mode = os.environ.get("HCU_DIST_EXCHANGE", "peer").lower()
    if mode not in ("peer", "nccl"):
        msg = f"HCU_DIST_EXCHANGE must be 'peer' or 'nccl', not {mode!r}"
        raise ValueError(msg)
    return mode


def attach_peers(lanes, ctx_of, plan, rank, device):
    """give every lane a PeerExchange (all ranks call this in the same order); on failure (no IPC between the ranks'
    devices) every rank falls back to the NCCL exchange together"""
    import warnings

    if plan.world <= 1 or exchange_mode() != "peer":
        return lanes
    for lane in lanes:
        # Opening the peers' buffers costs tens of milliseconds per rank (cudaIpcOpenMemHandle maps gigabytes): the
        # exchange of a (context, geometry, world, communicator) is created once per process and reused by every
        # pipeline built later -- like the library's cached plans and workspaces.  close_peer_exchanges() frees them.
        ctx = ctx_of(lane)
        key = (id(ctx), plan.nside, plan.lmax, plan.world, tuple(plan.rp_bounds), int(rank), id(lane.group))
        if key in _PEER_CACHE:
            lane.peers = _PEER_CACHE[key]
            continue
        try:  # (collective; fails on every rank together)
            lane.peers = _PEER_CACHE[key] = PeerExchange(ctx, plan, rank, group=lane.group, device=device)
        except RuntimeError as e:
            _PEER_CACHE[key] = None
            warnings.warn(f"{e}; using the NCCL all-to-all", stacklevel=2)
    return lanes


_PEER_CACHE: dict = {}


def close_peer_exchanges():
    """free the cached peer-memory exchange buffers of this process (every rank must call it)"""
    for ex in _PEER_CACHE.values():
        if ex is not None:
            ex.close()
    _PEER_CACHE.clear()


def make_lanes(ctx, nside, lmax, n=2, group=None):
    """
    ``n`` lanes for a CUDA DistributedTransform: lane 0 is the given context and process group on the CALLER's current
    stream (the one ``ctx`` is set to), every further lane has a library context, stream and world communicator of
    its own (creating a communicator costs about a second; they are cached per process).
    """
    import torch
    import torch.distributed as dist

    from . import _lib

    lanes = []
    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    for i in range(n):
        if i == 0:
            lanes.append(Lane(StagedKernels(ctx, nside, lmax), group, None))
            continue
        c = _lib.extra_context(ctx.device, i)
        st = torch.cuda.Stream(device=ctx.device)
        c.set_stream(st.cuda_stream)
        lanes.append(Lane(StagedKernels(c, nside, lmax), _extra_group(i) if multi else group, st))
    return lanes


class ReadyOnce:
    """
    "the maps of this spin may be read": waits for an asynchronous reduction, runs ``finish`` ONCE (on the first lane
    that asks) and makes every other stream that asks later wait for that, too.
    """

    def __init__(self, work=None, finish=None):
        self.work, self.finish, self.event = work, finish, None

    def __call__(self):
        import torch

        if self.work is not None:
            self.work.wait()  # the current stream waits for the collective
        if self.event is None:
            if self.finish is not None:
                self.finish()
            self.event = torch.cuda.Event()
            self.event.record()
        else:
            torch.cuda.current_stream().wait_event(self.event)


def reduce_maps(maps, group=None):
    """sum the partial maps of all ranks (every rank gets the full sum)"""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(maps, group=group)
    return maps


def allreduce_cl(cl, group=None):
    """Cl computed from m-distributed alm are partial sums over this rank's m"""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(cl, group=group)
    return cl


# ---------------------------------------------------------------------------------------
# public multi-GPU entry point on top of CudaHealpixMapper
# ---------------------------------------------------------------------------------------
class _CudaView:
    """exposes a raw device pointer through __cuda_array_interface__ (zero-copy torch view)"""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr, "version": 3, "strides": None}


def as_torch(arr, device):
    """torch view of a contiguous ``DeviceArray`` (the managed memory ``Mapper.create()`` returns)"""
    import torch

    ptr = getattr(arr, "device_ptr", None)
    if ptr is None:
        return torch.as_tensor(np.ascontiguousarray(arr)).to(device)
    return torch.as_tensor(_CudaView(ptr, arr.shape, arr.dtype.str), device=device)


_EXTRA_GROUPS: dict = {}


def _extra_group(i: int = 1):
    """extra world communicators, one per index, created on first use (creating one costs about a second);
    every rank must ask for them in the same order"""
    import torch.distributed as dist

    if i not in _EXTRA_GROUPS:
        _EXTRA_GROUPS[i] = dist.new_group()
    return _EXTRA_GROUPS[i]


def _reduce_group():
    """the communicator of lane 1, which also carries the asynchronous spin-2 map reduction"""
    return _extra_group(1)


class DistributedPipeline:
    """
    Maps of one ``CudaHealpixMapper`` (every rank mapped ITS pages into them) -> angular power
    spectra over all ranks.

        dp = DistributedPipeline(mapper, npos, nshe)
        dp.put(0, i, pos_map)          # (npix,)   partial spin-0 map i of this rank
        dp.put(2, i, she_map)          # (2, npix) partial spin-2 map i of this rank
        cl = dp.spectra(finish=...)    # [ncomp, ncomp, lmax + 1] on every rank

    ``put`` copies the (managed-memory) map into a device-resident stack, so the caller can free
    it right away; ``spectra`` sums the stacks over the ranks, applies ``finish(stack, spin)``
    (e.g. the visibility subtraction, which must happen once, after the sum), transforms them
    ring-block / m-distributed and reduces the component spectra (rows: spin-0 maps first, then
    (E, B) per spin-2 field; only the upper triangle j >= i is filled).
    """

    def __init__(self, mapper, npos: int = 0, nshe: int = 0, group=None, lanes: int | None = None):
        import torch
        import torch.distributed as dist

        if lanes is None:
            lanes = int(os.environ.get("HCU_DIST_LANES", "1"))

        self.mapper, self.group = mapper, group
        self.ctx = mapper.context
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device("cuda", self.ctx.device)
        self.plan = ShardPlan(mapper.nside, mapper.lmax, self.world)
        # lanes = 2: the ring FFTs and the all-to-all of one Legendre batch are queued under the Legendre kernels of
        # another.  Measured (profiles/r02_lanes_c3_n2.txt): no gain -- a Legendre CTA takes a whole SM (all registers,
        # 205 KB of shared memory), so the second lane's kernels time-slice the SMs instead of sharing them; default 1.
        lanes = make_lanes(self.ctx, mapper.nside, mapper.lmax, lanes if self.world > 1 else 1, group=group)
        # the exchange goes through NVLink peer memory unless HCU_DIST_EXCHANGE=nccl (or IPC is unavailable)
        attach_peers(lanes, lambda lane: lane.k.ctx, self.plan, self.rank, self.device)
        self.kernels = lanes[0].k
        self.transform = DistributedTransform(self.kernels, self.plan, self.rank, group=group, niter=mapper.niter,
                                              device=self.device, lanes=lanes)
        # a communicator of its own for the asynchronous spin-2 map reduction (collectives of ONE communicator run in
        # issue order, so it would otherwise queue in front of the spin-0 exchange)
        self.reduce_group = (lanes[-1].group if len(lanes) > 1 else _reduce_group()) if self.world > 1 and group is None else group
        # from here on the library works on torch's current stream, so that the mapper's kernels, the copies of `put`
        # and the transforms are ordered without host synchronisation
        self.ctx.synchronize()
        self.kernels.sync_streams()
        self.stacks = {}
        if npos:
            self.stacks[0] = torch.zeros(npos, self.plan.npix, dtype=torch.float64, device=self.device)
        if nshe:
            self.stacks[2] = torch.zeros(2 * nshe, self.plan.npix, dtype=torch.float64, device=self.device)

    def put(self, spin: int, index: int, m) -> None:
        """copy partial map ``index`` of the given spin into the device stack (asynchronous on the library's stream:
        the caller may reuse or free ``m`` after the next ``synchronize`` of the context, which ``spectra`` does)"""
        self.kernels.sync_streams()  # the copy is ordered after the kernels that mapped into m
        if hasattr(m, "to_device"):
            m.to_device()
        src = as_torch(m, self.device).reshape(-1, self.plan.npix)
        n = src.shape[0]
        self.stacks[spin][index * n:(index + 1) * n].copy_(src)
        self._pending = True

    def flush(self) -> None:
        """wait until every ``put`` has landed (the sources may be reused afterwards)"""
        import torch

        if getattr(self, "_pending", False):
            torch.cuda.current_stream().synchronize()
            self._pending = False

    def alms(self, spin, finish=None, reduced=False):
        """stack of one spin -> m-distributed alm tensor [rows, nalm] (reduced: already summed over the ranks)"""
        import torch

        self.kernels.sync_streams()
        stack = self.stacks[spin]
        if not reduced:
            reduce_maps(stack, self.group)
        if finish is not None:
            finish(stack, spin)
        alm = torch.zeros(stack.shape[0], self.plan.nalm, dtype=torch.complex128, device=self.device)
        self.transform.map2alm(stack, spin, alm, fl=self.mapper._fl(spin))
        return alm

    def spectra(self, finish=None):
        import torch

        import torch.distributed as dist

        self.kernels.sync_streams()
        jobs, ready, alms = [], {}, []
        for spin in (0, 2):
            if spin not in self.stacks:
                continue
            stack = self.stacks[spin]
            fin = (lambda st=stack, sp=spin: finish(st, sp)) if finish is not None else None
            work = None
            if self.world > 1:
                if spin == 0:
                    dist.all_reduce(stack, group=self.group)
                else:  # summed over the ranks on lane 1's communicator while the spin-0 batch is transformed
                    work = dist.all_reduce(stack, group=self.reduce_group, async_op=True)
            ready[spin] = ReadyOnce(work, fin)
            alm = torch.zeros(stack.shape[0], self.plan.nalm, dtype=torch.complex128, device=self.device)
            alms.append(alm)
            jobs.append((stack, spin, alm, self.mapper._fl(spin)))
        self.transform.map2alm_jobs(jobs, ready)
        alm = torch.cat(alms) if len(alms) > 1 else alms[0]
        n, lmax = alm.shape[0], self.plan.lmax
        cl = torch.zeros(n, n, lmax + 1, dtype=torch.float64, device=self.device)
        # only the m this rank owns contribute (the other entries are zero and need not be read)
        self.kernels._check(self.ctx.lib.hcu_alm2cl_mslice(self.ctx.handle, n, alm.data_ptr(), alm.stride(0), lmax, n,
                                                           alm.data_ptr(), alm.stride(0), lmax, lmax, self.world, self.rank,
                                                           cl.data_ptr()))
        allreduce_cl(cl, self.group)
        return cl
