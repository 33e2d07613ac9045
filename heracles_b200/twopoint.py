"""
alm -> Cl on the device: ``alm2cl`` with the reference's signature and block
output (``heracles/twopoint.py:63-101``) and an ``angular_power_spectra`` that
keeps the reference's pair selection, key order, metadata and bias rules
(``heracles/twopoint.py:173-299``) while every spectrum comes from the batched
``hcu_alm2cl`` kernel.
"""

from __future__ import annotations

from itertools import combinations_with_replacement, product

import numpy as np

from . import _lib
from .arrays import DeviceArray, update_metadata

c_vp = _lib.c_vp


def alm2lmax(alm, mmax=None):
    """lmax of an alm array (twopoint.py:55-60)"""
    return (int((8 * np.shape(alm)[-1] + 1) ** 0.5 + 0.01) - 3) // 2


def _rows(alm, ctx):
    """(pointer, nrows, stride, keepalive) for a (..., nalm) complex128 array"""
    n = alm.shape[-1]
    if isinstance(alm, DeviceArray) and alm.device_ptr is not None and alm.dtype == np.complex128:
        alm.to_device()
        return alm.device_ptr, alm.size // n, n, alm
    host = np.ascontiguousarray(alm, dtype=np.complex128)
    dev = DeviceArray.zeros(ctx, host.shape, dtype=np.complex128)
    ctx.memcpy(dev.device_ptr, host.__array_interface__["data"][0], host.nbytes)
    ctx.synchronize()
    return dev.device_ptr, host.size // n, n, dev


def alm2cl(alm, alm2=None, *, lmax=None, context=None):
    """
    Angular (cross-)power spectrum block of *alm* and *alm2*; output shape
    ``(*alm.shape[:-1], *alm2.shape[:-1], lmax_out + 1)`` like the reference.
    """
    ctx = context or _lib.get_context()
    same = alm2 is None or alm2 is alm
    alm = np.asanyarray(alm)
    alm2 = alm if same else np.asanyarray(alm2)
    l1, l2 = alm2lmax(alm), alm2lmax(alm2)
    if lmax is None:
        lmax = min(l1, l2)
    lout = min(lmax, l1, l2)
    pa, na, sa, keep_a = _rows(alm, ctx)
    if same:
        pb, nb, sb, keep_b = pa, na, sa, keep_a
    else:
        pb, nb, sb, keep_b = _rows(alm2, ctx)
    cl = DeviceArray.zeros(ctx, (na, nb, lout + 1), dtype=np.float64)
    _lib.check(ctx.lib.hcu_alm2cl(ctx.handle, na, c_vp(pa), sa, l1, nb, c_vp(pb), sb, l2, lmax, c_vp(cl.device_ptr)))
    ctx.synchronize()
    del keep_a, keep_b
    out = np.array(cl._host(), copy=True)
    return out.reshape(*alm.shape[:-1], *alm2.shape[:-1], lout + 1)


def _toc_match(key, include, exclude):
    """heracles.core.toc_match (core.py:63-88)"""

    def match(pattern):
        return all(p is Ellipsis or p == k for p, k in zip(pattern, key))

    if include is not None and not any(match(p) for p in include):
        return False
    if exclude is not None and any(match(p) for p in exclude):
        return False
    return True


def _debias(cl, bias, md):
    """twopoint._debias_cl for non-deconvolved HEALPix kernels or explicit pixel windows"""
    spin1, spin2 = md.get("spin_1", 0), md.get("spin_2", 0)
    lmin = max(abs(spin1), abs(spin2))
    lmax = cl.shape[-1] - 1
    bl = np.zeros(cl.shape)
    if spin1 != 0 and spin2 != 0:
        bl[[0, 1], [0, 1], ..., lmin:] = bias
    else:
        bl[..., lmin:] = bias
    for i, s in (1, spin1), (2, spin2):
        if md.get(f"kernel_{i}") == "healpix":
            nside = md.get(f"nside_{i}")
            deconv = md.get(f"deconv_{i}", True)
            if nside is not None and deconv:
                from .mapper import CudaHealpixMapper

                pw = CudaHealpixMapper(nside, lmax, deconvolve=True)._get_pixwin()
                pw = pw[0] if s == 0 else pw[1] if s == 2 else None
                if pw is not None:
                    bl[..., lmin:] /= np.asarray(pw)[lmin : lmax + 1]
    cl[:] -= bl
    return cl


def angular_power_spectra(
    alms,
    alms2=None,
    *,
    lmax=None,
    debias=True,
    bins=None,
    weights=None,
    include=None,
    exclude=None,
    out=None,
    context=None,
):
    """
    Drop-in for ``heracles.twopoint.angular_power_spectra``.  Keys, ordering,
    metadata (``*_1`` / ``*_2``, ``bias``) and the debiasing follow the
    reference; the spectra are computed on the device, one ``hcu_alm2cl`` block
    per pair of alm arrays, with the alm resident in managed memory.
    When the ``heracles`` package is importable the results are wrapped in its
    ``Result`` type and optionally binned, exactly as upstream does.
    """
    ctx = context or _lib.get_context()
    if alms2 is None:
        pairs = combinations_with_replacement(alms, 2)
        alms2 = alms
    else:
        pairs = product(alms, alms2)

    try:  # optional: upstream result container
        from heracles.result import Result, binned
    except Exception:  # pragma: no cover - heracles not installed next to this backend
        Result = binned = None
    if bins is not None and binned is None:
        raise RuntimeError("binning needs the heracles package (heracles.result.binned)")

    cls = {} if out is None else out
    names = set()
    staged: dict = {}

    def dev(a):
        # the entry keeps `a` itself alive: a lazily loading alm mapping hands out a NEW array per
        # access, and a freed array's id() would otherwise be recycled for a different (k, i)
        key = id(a)
        if key not in staged:
            if isinstance(a, DeviceArray) and a.device_ptr is not None:
                staged[key] = (a, a)
            else:  # host alm: upload once, reuse for every pair it takes part in
                h = np.ascontiguousarray(a, dtype=np.complex128)
                d = DeviceArray.zeros(ctx, h.shape, dtype=np.complex128)
                ctx.memcpy(d.device_ptr, h.__array_interface__["data"][0], h.nbytes)
                ctx.synchronize()
                staged[key] = (a, d)
        return staged[key][1]

    for (k1, i1), (k2, i2) in pairs:
        if (k1, k2, i1, i2) in cls or (k2, k1, i2, i1) in cls:
            continue
        swapped = (k1, k2) not in names and (k2, k1) in names
        if swapped:
            i1, i2 = i2, i1
            k1, k2 = k2, k1
        if not _toc_match((k1, k2, i1, i2), include, exclude):
            continue
        if swapped:
            alm1, alm2 = alms2[k1, i1], alms[k2, i2]
        else:
            alm1, alm2 = alms[k1, i1], alms2[k2, i2]

        cl = alm2cl(dev(alm1), dev(alm2), lmax=lmax, context=ctx)

        md1 = alm1.dtype.metadata or {}
        md2 = alm2.dtype.metadata or {}
        s1, s2 = md1.get("spin", None), md2.get("spin", None)
        if s1 is None or s2 is None:
            raise ValueError(f"missing spin metadata for {k1} or {k2}")
        md = {}
        for key, value in md1.items():
            md[f"{key}_1"] = value
        for key, value in md2.items():
            md[f"{key}_2"] = value
        bias = None
        if k1 == k2 and i1 == i2:
            fsky, musq, dens = md1.get("fsky"), md1.get("musq"), md1.get("dens")
            if fsky is not None and musq is not None and dens is not None:
                bias = (0.5 if s1 == s2 == 2 else 1.0) * fsky * musq / dens
        if bias is not None:
            md["bias"] = bias
        if debias and bias is not None:
            _debias(cl, bias, md)
        update_metadata(cl, **md)
        if Result is not None:
            cl = Result(cl, spin=(s1, s2), axis=-1)
            if bins is not None:
                cl = binned(cl, bins, weights)
        cls[k1, k2, i1, i2] = cl
        names.add((k1, k2))
    return cls
