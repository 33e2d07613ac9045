"""
Batched drop-in for ``heracles.mapping.transform`` (``heracles/mapping.py:130-174``).

The reference transforms the maps one at a time (``mapper.transform(m, spin=s)``
per dictionary entry).  The Legendre recursion on the GPU costs the same for
one map as for ten, so this version groups the entries that share a
``CudaHealpixMapper`` configuration and spin and transforms each group with one
``hcu_map2alm_many`` call.  Keys, spin checks, error messages, metadata and the
output dictionary are the reference's; entries whose field uses another mapper
are transformed through that mapper's own ``transform``.
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .arrays import DeviceArray, update_metadata
from .mapper import CudaHealpixMapper, _native, _ptr


def transform_maps(mapper: CudaHealpixMapper, maps, spin: int = 0):
    """
    Transform a list of maps of ONE mapper and spin in as few device passes as
    possible.  ``maps`` is a sequence of arrays of shape ``(npix,)`` (spin 0) or
    ``(2, npix)`` (spin 2); returns the list of alm arrays (``(nalm,)`` or
    ``(2, nalm)``), each carrying the map's metadata plus ``deconv``.
    """
    if spin not in (0, 2):
        msg = f"spin-{spin} maps not yet supported"
        raise NotImplementedError(msg)
    ctx = mapper.context
    npix = mapper.npix
    lmax = mapper.lmax
    nalm = (lmax + 1) * (lmax + 2) // 2
    fl = mapper._fl(spin)
    pw = mapper._pixel_weights
    mapper._apply_modes()
    keep = []
    rows_in, rows_out, alms = [], [], []
    for m in maps:
        if m.shape[-1] != npix:
            raise ValueError("data is not a map of this mapper")
        lead = m.shape[:-1]
        if spin == 2 and (len(lead) == 0 or lead[-1] != 2):
            raise ValueError("spin-2 data must have shape (..., 2, npix)")
        m = mapper._ring_view(m)
        if isinstance(m, DeviceArray) and m.device_ptr is not None:
            m.to_device()
            src = m
        else:
            src = _native(m)
        keep.append(src)
        alm = DeviceArray.zeros(ctx, (*lead, nalm), dtype=np.complex128)
        alms.append(alm)
        base_in = src.device_ptr if isinstance(src, DeviceArray) else _ptr(src)
        nrow = int(np.prod(lead, dtype=np.int64)) if lead else 1
        for r in range(nrow):
            rows_in.append(base_in + r * npix * 8)
            rows_out.append(alm.device_ptr + r * nalm * 16)
    n = len(rows_in)
    if n:
        pin = (ctypes.c_void_p * n)(*rows_in)
        pout = (ctypes.c_void_p * n)(*rows_out)
        _lib.check(
            ctx.lib.hcu_map2alm_many(
                ctx.handle, mapper.nside, lmax, spin, n, pin, None,
                ctypes.c_void_p(_ptr(pw) if pw is not None else 0), mapper.niter,
                ctypes.c_void_p(_ptr(fl) if fl is not None else 0), pout,
            )
        )
    for m, alm in zip(maps, alms):
        update_metadata(alm, **{**(m.dtype.metadata or {}), "deconv": mapper.deconvolve})
    del keep
    return alms


def _group_key(mapper):
    return (
        id(mapper.context), mapper.nside, mapper.lmax, mapper.deconvolve, mapper.niter,
        id(mapper._pixwin), id(mapper._pixel_weights), mapper.weights_mode, mapper.scheme,
    )


def transform(fields, data, *, out=None, progress=None):
    """transform data to alms -- signature and behaviour of heracles.mapping.transform"""
    if out is None:
        try:
            from heracles.core import TocDict

            out = TocDict()
        except Exception:  # heracles itself not importable next to this backend
            out = {}

    current, total = 0, len(data)
    groups: dict = {}
    order = []
    for (k, i), m in data.items():
        current += 1
        if progress is not None:
            progress.update(current, total)
        m = getattr(m, "array", m)
        try:
            field = fields[k]
        except KeyError:
            msg = f"unknown field name: {k}"
            raise ValueError(msg) from None
        s = field.spin
        m_spin = (m.dtype.metadata or {}).get("spin")
        if m_spin is None:
            update_metadata(m, spin=s)
        elif m_spin != s:
            msg = f"spin mismatch for field {k!r}: map has spin {m_spin}, field has spin {s}"
            raise ValueError(msg)
        mapper = field.mapper_or_error
        order.append((k, i))
        if isinstance(mapper, CudaHealpixMapper) and s in (0, 2):
            groups.setdefault((_group_key(mapper), s), (mapper, s, []))[2].append(((k, i), m))
        else:
            out[k, i] = mapper.transform(m, spin=s)

    results = {}
    for mapper, s, entries in groups.values():
        alms = transform_maps(mapper, [m for _, m in entries], spin=s)
        for (key, _), alm in zip(entries, alms):
            results[key] = alm
    for key in order:  # reference insertion order
        if key in results:
            out[key] = results[key]
    return out
