"""
Transforms that run BESIDE the catalogue mapping.

The reference maps every catalogue first (``heracles.mapping.map_catalogs``, ``heracles/mapping.py:61-127``) and
transforms afterwards (``transform``, ``:130-174``).  On a B200 the two phases use different parts of the machine: the
mapping is bound by the PCIe link (40 bytes per catalogue row, 80 GB for 2e9 rows) and needs a few percent of the SMs,
the transform is FP64-bound and needs no host data.  :class:`OverlappedTransform` hands every finished map to a
worker thread that transforms Legendre-batch-sized groups on a SECOND library context (its own CUDA stream and
workspaces) while the caller keeps mapping the next tomographic bin on the first; the mapping stream has the higher
priority, so its short scatter kernels slip in between the Legendre CTAs (a Legendre CTA owns a whole SM, so the two
do not run side by side on one SM: the gain is bounded, see ``priority``).

    ov = OverlappedTransform(mapper)
    for bin in bins:
        ... mapper.map_page(...) for all pages of the bin, normalise ...
        ov.submit(("POS", bin), pos, spin=0)
        ov.submit(("SHE", bin), she, spin=2)
    alms = ov.finish()            # dict in submission order, like heracles_b200.transform

The results are the ones ``heracles_b200.transform`` returns for the same maps (same kernels, same batching rules).
"""

from __future__ import annotations

import queue
import threading

from . import _lib
from .arrays import update_metadata
from .mapper import CudaHealpixMapper

__all__ = ["OverlappedTransform"]


class OverlappedTransform:
    def __init__(self, mapper: CudaHealpixMapper, batch=None, priority: str | None = None):
        """priority: which side gets the high-priority CUDA stream while both run -- "mapping" (default: its short
        scatter kernels go first whenever an SM frees up) or "transform" (the mapping only gets what the Legendre
        kernels leave free).  Measured at C4 (2e9 rows, 20 fields, one B200): 16.2-16.4 s end to end with "mapping",
        16.9 s with "transform" (a scatter launch then waits for the end of a 0.7 s Legendre launch, the mapping
        stretches to 10 s and holds the later batches up) and 16.9 s without any overlap.  HCU_OVERLAP_PRIORITY overrides."""
        import os

        priority = os.environ.get("HCU_OVERLAP_PRIORITY", priority or "mapping")
        if priority not in ("transform", "mapping"):
            raise ValueError("priority must be 'transform' or 'mapping'")
        self.mapper = mapper
        ctx = mapper.context
        # a mapper of the same geometry and options on a library context of its own
        self.worker = CudaHealpixMapper(
            mapper.nside, mapper.lmax, deconvolve=mapper.deconvolve, niter=mapper.niter, pixwin=mapper._pixwin,
            pixel_weights=mapper._pixel_weights_arg, weights_mode=mapper.weights_mode, scheme=mapper.scheme,
            device=ctx.device, sync=mapper.sync, context=_lib.extra_context(ctx.device, 101),
        )
        # Legendre batch sizes in maps: 12 spin-0 maps, 4 spin-2 fields (8 components)
        self.batch = {0: int(ctx.lib.hcu_legendre_batch_size(0)), 2: int(ctx.lib.hcu_legendre_batch_size(2)) // 2}
        if batch:
            self.batch.update(batch)
        import torch

        # one of the two contexts moves to a high-priority stream for the lifetime of this object
        with torch.cuda.device(ctx.device):
            self._stream = torch.cuda.Stream(priority=-1)
        self._hi = ctx if priority == "mapping" else self.worker.context
        self._hi.synchronize()
        self._hi.set_stream(self._stream.cuda_stream)
        self._q: queue.Queue = queue.Queue()
        self._order: list = []
        self._alms: dict = {}
        self._error: BaseException | None = None
        self._thread = threading.Thread(target=self._run, name="heracles-b200-transform", daemon=True)
        self._thread.start()

    # -- caller side --------------------------------------------------------------------------------------------
    def submit(self, key, m, spin: int = 0) -> None:
        """map `m` is final (all pages mapped, normalised): transform it as soon as its batch is complete"""
        if spin not in (0, 2):
            msg = f"spin-{spin} maps not yet supported"
            raise NotImplementedError(msg)
        if self._error is not None:
            raise self._error
        m = getattr(m, "array", m)
        if (m.dtype.metadata or {}).get("spin") is None:
            update_metadata(m, spin=spin)
        if hasattr(m, "to_device"):
            m.to_device()
        self.mapper.context.synchronize()  # every kernel that wrote the map has finished
        self._order.append(key)
        self._q.put((key, m, spin))

    def finish(self) -> dict:
        """transform what is left (incomplete batches) and return ``{key: alm}`` in submission order"""
        self._q.put(None)
        self._thread.join()
        if self._stream is not None:
            self._hi.synchronize()
            self._hi.set_stream(None)  # back to the context's own stream
            self._stream = None
        if self._error is not None:
            raise self._error
        return {k: self._alms[k] for k in self._order}

    # -- worker thread ------------------------------------------------------------------------------------------
    def _flush(self, pending, spin):
        from .mapping import transform_maps

        entries = pending[spin]
        if not entries:
            return
        alms = transform_maps(self.worker, [m for _, m in entries], spin=spin)
        self.worker.context.synchronize()
        for (key, _), alm in zip(entries, alms):
            self._alms[key] = alm
        pending[spin] = []

    def _run(self):
        pending = {0: [], 2: []}
        try:
            while True:
                item = self._q.get()
                if item is None:
                    break
                key, m, spin = item
                pending[spin].append((key, m))
                if len(pending[spin]) >= self.batch[spin]:
                    self._flush(pending, spin)
            for spin in (2, 0):
                self._flush(pending, spin)
        except BaseException as e:  # noqa: BLE001 - re-raised in the caller's thread
            self._error = e
            while True:  # drain, so that submit() never blocks
                try:
                    if self._q.get_nowait() is None:
                        break
                except queue.Empty:
                    break
