"""
alm -> Cl on the device: ``alm2cl`` with the reference's signature and block
output (``heracles/twopoint.py:63-101``) and an ``angular_power_spectra`` that
keeps the reference's pair selection, key order, metadata and bias rules
(``heracles/twopoint.py:173-299``) while every spectrum comes from ONE batched
``hcu_alm2cl_rows`` launch over all alm.
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


def _pixwin_for(nside, lmax, pixwin):
    """(pw_T, pw_P) for the shot-noise debiasing: the caller's arrays, else HEALPix' table under DATAPATH / healpy"""
    if pixwin is not None:
        if isinstance(pixwin, np.ndarray) and pixwin.ndim == 1:
            return pixwin, pixwin
        return pixwin
    from .mapper import CudaHealpixMapper, read_pixwin_fits
    import os

    try:
        import healpy

        return healpy.pixwin(nside, lmax=lmax, pol=True)
    except Exception:
        pass
    if CudaHealpixMapper.DATAPATH:
        path = os.path.join(CudaHealpixMapper.DATAPATH, "pixel_window_n%04d.fits" % nside)
        if os.path.exists(path):
            return read_pixwin_fits(path)
    raise RuntimeError(
        "debiasing deconvolved HEALPix spectra needs the pixel window: pass pixwin=(pw_T, pw_P) (the arrays the mapper "
        "deconvolved with) to angular_power_spectra, set CudaHealpixMapper.DATAPATH, or use debias=False"
    )


def _debias(cl, bias, md, pixwin=None):
    """twopoint._debias_cl (heracles/twopoint.py:104-170); no device context is touched here"""
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
                pw = _pixwin_for(nside, lmax, pixwin)
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
    pixwin=None,
):
    """
    Drop-in for ``heracles.twopoint.angular_power_spectra``.  Keys, ordering,
    metadata (``*_1`` / ``*_2``, ``bias``) and the debiasing follow the
    reference (``heracles/twopoint.py:198-290``); the reference calls ``alm2cl`` once per pair of alm arrays
    (``:243``), here ALL pairs come from ONE ``hcu_alm2cl_rows`` launch over the device-resident alm (every alm is
    read once per 4 x 4 tile of spectra) and each pair's block is sliced out of it.  Mixed ``lmax`` or more than 64
    alm rows fall back to one ``hcu_alm2cl`` per pair.
    ``pixwin``: the ``(pw_T, pw_P)`` the mapper deconvolved with -- used for the bias of deconvolved spectra
    (the reference asks healpy for it, ``twopoint.py:152-161``).
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

    # ---- pass 1: the reference's pair selection (twopoint.py:204-236) ----
    names = set()
    todo = []
    seen = set(cls)
    for (k1, i1), (k2, i2) in pairs:
        if (k1, k2, i1, i2) in seen or (k2, k1, i2, i1) in seen:
            continue
        swapped = (k1, k2) not in names and (k2, k1) in names
        if swapped:
            i1, i2 = i2, i1
            k1, k2 = k2, k1
        if not _toc_match((k1, k2, i1, i2), include, exclude):
            continue
        # which mapping each side comes from (twopoint.py:232-236)
        src1, src2 = (1, 0) if swapped else (0, 1)
        todo.append((k1, k2, i1, i2, src1, src2))
        seen.add((k1, k2, i1, i2))
        names.add((k1, k2))

    # ---- the alm arrays taking part: fetched ONCE per (mapping, key) -- a lazily loading mapping returns a new
    # array per access -- and kept alive together with their device copy until the end of the call ----
    maps_ = (alms, alms2)
    same_mapping = alms2 is alms
    arrays = {}

    def entry(src, k, i):
        key = (0 if same_mapping else src, k, i)
        if key not in arrays:
            a = maps_[src][k, i]
            if isinstance(a, DeviceArray) and a.device_ptr is not None and a.dtype == np.complex128:
                a.to_device()
                d = a
            else:  # host alm: upload once, reuse for every pair it takes part in
                h = np.ascontiguousarray(a, dtype=np.complex128)
                d = DeviceArray.zeros(ctx, h.shape, dtype=np.complex128)
                ctx.memcpy(d.device_ptr, h.__array_interface__["data"][0], h.nbytes)
            arrays[key] = (a, d)
        return arrays[key]

    for k1, k2, i1, i2, s1_, s2_ in todo:
        entry(s1_, k1, i1)
        entry(s2_, k2, i2)
    ctx.synchronize()

    # ---- one Gram block over all rows when the arrays share lmax ----
    gram = None
    row0 = {}
    lmaxes = {alm2lmax(d) for _, d in arrays.values()}
    nrows = sum(d.size // d.shape[-1] for _, d in arrays.values())
    if arrays and len(lmaxes) == 1 and nrows <= 64:
        import ctypes

        lin = lmaxes.pop()
        lout = lin if lmax is None else min(lmax, lin)
        ptrs = []
        for key, (_, d) in arrays.items():
            row0[key] = len(ptrs)
            n = d.shape[-1]
            ptrs.extend(d.device_ptr + 16 * n * r for r in range(d.size // n))
        tab = (ctypes.c_void_p * len(ptrs))(*ptrs)
        dev = DeviceArray.zeros(ctx, (len(ptrs), len(ptrs), lout + 1), dtype=np.float64)
        _lib.check(ctx.lib.hcu_alm2cl_rows(ctx.handle, len(ptrs), tab, lin, lout, c_vp(dev.device_ptr)))
        ctx.synchronize()
        gram = np.array(dev._host(), copy=True)
        del dev

    # ---- pass 2: metadata, bias, Result (twopoint.py:238-290) ----
    for k1, k2, i1, i2, s1_, s2_ in todo:
        key1, key2 = (0 if same_mapping else s1_, k1, i1), (0 if same_mapping else s2_, k2, i2)
        (alm1, d1), (alm2, d2) = arrays[key1], arrays[key2]
        if gram is not None:
            n1, n2 = d1.size // d1.shape[-1], d2.size // d2.shape[-1]
            r1, r2 = row0[key1], row0[key2]
            cl = np.array(gram[r1:r1 + n1, r2:r2 + n2], copy=True).reshape(*d1.shape[:-1], *d2.shape[:-1], gram.shape[-1])
        else:
            cl = alm2cl(d1, d2, lmax=lmax, context=ctx)

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
            _debias(cl, bias, md, pixwin)
        update_metadata(cl, **md)
        if Result is not None:
            cl = Result(cl, spin=(s1, s2), axis=-1)
            if bins is not None:
                cl = binned(cl, bins, weights)
        cls[k1, k2, i1, i2] = cl
    return cls
