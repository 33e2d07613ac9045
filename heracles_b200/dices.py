"""
Batched DICES jackknife on the device -- the heaviest repeat caller of the hot path.

The reference (``heracles/dices/jackknife.py``) computes, for every jackknife region ``k = 1..njk`` and for the full
footprint ``k = 0``, ``transform(fields, region-masked maps)`` for the data AND the visibility maps -- a deep copy and
a host-side mask multiply per region (``_get_region_maps``), one ``hp.map2alm`` per map, every alm set written to FITS
-- and then, for every tuple of ``nd`` deleted regions, ``angular_power_spectra`` of ``alm_full - sum alm_regions``
(``_compute_single_jk_cls``, ``_subtract_alms``, ``_accumulate_alms``).

Here the maps stay on the device: the region masks are applied by a kernel straight into the transform's input batch,
the region maps of ONE map share the Legendre recursion (up to 12 spin-0 / 4 spin-2 region maps per launch instead of
one map per ``map2alm``), the alm stay in device memory, the delete-``nd`` alm are formed with ``hcu_axpy`` and every
tuple's spectra come from one ``hcu_alm2cl_rows`` Gram launch.  Keys, metadata and the order of the results follow the
reference; its bias and mask corrections (``correct_bias``, ``correct_footprint_*``) are host-side arithmetic on
(lmax + 1)-length arrays and are applied through the ``correct`` hook when the ``heracles`` package is importable.
"""

from __future__ import annotations

from itertools import combinations

import numpy as np

from . import _lib
from .arrays import DeviceArray, update_metadata
from .mapper import CudaHealpixMapper, _native, _ptr
from .mapping import transform_maps
from .twopoint import angular_power_spectra

c_vp = _lib.c_vp

__all__ = ["region_ids", "region_alms", "delete_alms", "jackknife_cls"]


def region_ids(jk_map):
    """the jackknife regions of a region map: its distinct non-zero values (jackknife.py: ``njk``)"""
    u = np.unique(np.asarray(jk_map))
    return [float(v) for v in u if v != 0]


def _device_map(ctx, m):
    if isinstance(m, DeviceArray) and m.device_ptr is not None:
        m.to_device()
        return m
    d = DeviceArray.zeros(ctx, np.shape(m))
    h = _native(m)
    ctx.memcpy(d.device_ptr, _ptr(h), h.nbytes)
    update_metadata(d, **(getattr(m, "dtype", np.dtype(float)).metadata or {}))
    return d


def region_alms(fields, maps, jk_map, regions=None, *, progress=None):
    """
    ``{k: transform(fields, _get_region_maps(maps, jk_map, k))}`` for the given regions (default: all), computed in
    Legendre-batch-sized groups of region maps straight from device-resident maps.  ``k = 0`` stands for the full
    footprint (no mask), as in ``_compute_single_jk_alm``.  Every field must use a ``CudaHealpixMapper``.
    Returns ``{k: {(name, bin): alm}}`` with the alm in device (managed) memory and the maps' metadata plus ``deconv``.
    """
    jk_map = np.asarray(jk_map, dtype=np.float64)
    if regions is None:
        regions = region_ids(jk_map)
    regions = [float(r) for r in regions]
    out = {r if r != int(r) else int(r): {} for r in regions}
    keys = list(out)
    total, current = len(maps), 0
    jk_dev = {}
    for (name, i), m in maps.items():
        current += 1
        if progress is not None:
            progress.update(current, total)
        m = getattr(m, "array", m)
        try:
            field = fields[name]
        except KeyError:
            msg = f"unknown field name: {name}"
            raise ValueError(msg) from None
        spin = field.spin
        mapper = field.mapper_or_error
        if not isinstance(mapper, CudaHealpixMapper):
            raise TypeError("heracles_b200.dices needs fields mapped with CudaHealpixMapper")
        ctx = mapper.context
        npix = mapper.npix
        if jk_map.shape[-1] != npix:
            raise ValueError("jk_map and maps differ in size")
        if id(ctx) not in jk_dev:
            jk_dev[id(ctx)] = _device_map(ctx, jk_map)
        jk = jk_dev[id(ctx)]
        src = _device_map(ctx, m)
        md = {**(m.dtype.metadata or {})}
        md.setdefault("spin", spin)
        nrow = src.size // npix  # 1 (spin 0) or 2 (spin 2)
        cap = int(ctx.lib.hcu_legendre_batch_size(spin)) // nrow
        for g0 in range(0, len(regions), cap):
            group = regions[g0:g0 + cap]
            work = DeviceArray.zeros(ctx, (len(group), *src.shape))
            for j, r in enumerate(group):
                for row in range(nrow):
                    dst = work.device_ptr + 8 * npix * (j * nrow + row)
                    s = src.device_ptr + 8 * npix * row
                    if r == 0:
                        ctx.memcpy(dst, s, 8 * npix)
                    else:
                        _lib.check(ctx.lib.hcu_region_select(ctx.handle, c_vp(dst), c_vp(s), c_vp(jk.device_ptr), r, npix))
            parts = [work[j] for j in range(len(group))]
            for p in parts:
                update_metadata(p, **md)
            alms = transform_maps(mapper, parts, spin=spin)
            for j, alm in enumerate(alms):
                out[keys[g0 + j]][name, i] = alm
            del work, parts
    return out


def delete_alms(alms_full, alms_regions, regions):
    """``alm_full - sum_{r in regions} alm_r`` on the device (``_subtract_alms(_accumulate_alms(...))``); metadata of the full alm"""
    out = {}
    for key, full in alms_full.items():
        ctx = full._hcu.ctx if isinstance(full, DeviceArray) and full._hcu is not None else _lib.get_context()
        d = DeviceArray.zeros(ctx, full.shape, dtype=np.complex128)
        ctx.memcpy(d.device_ptr, _device_alm(ctx, full).device_ptr, full.nbytes)
        for r in regions:
            a = _device_alm(ctx, alms_regions[r][key])
            _lib.check(ctx.lib.hcu_axpy(ctx.handle, c_vp(d.device_ptr), c_vp(a.device_ptr), -1.0, 2 * full.size))
        update_metadata(d, **(full.dtype.metadata or {}))
        out[key] = d
    return out


def _device_alm(ctx, a):
    if isinstance(a, DeviceArray) and a.device_ptr is not None:
        a.to_device()
        return a
    h = np.ascontiguousarray(a, dtype=np.complex128)
    d = DeviceArray.zeros(ctx, h.shape, dtype=np.complex128)
    ctx.memcpy(d.device_ptr, h.__array_interface__["data"][0], h.nbytes)
    ctx.synchronize()
    return d


def jackknife_cls(data_maps, jk_map, fields, *, nd=1, correct=None, progress=None, **spectra_kwargs):
    """
    Delete-``nd`` jackknife spectra ``{regions: angular_power_spectra(alm_full - sum alm_regions)}`` for every tuple of
    ``nd`` regions (``compute_jk_cls_from_alms``), ``nd = 0`` giving ``{(): cls}`` of the full footprint.
    ``correct(cls, regions)`` is applied to every tuple's spectra (the reference's ``correct_bias`` /
    ``correct_footprint_fsky`` chain); ``spectra_kwargs`` go to :func:`angular_power_spectra`.
    """
    if nd < 0 or nd > 2:
        raise ValueError("number of deletions must be 0, 1 or 2")
    ids = region_ids(jk_map)
    alms = region_alms(fields, data_maps, jk_map, [0] + ids, progress=progress)
    full = alms[0]
    cls = {}
    for regions in combinations([k for k in alms if k != 0], nd):
        a = delete_alms(full, alms, regions) if regions else full
        c = angular_power_spectra(a, **spectra_kwargs)
        cls[regions] = correct(c, regions) if correct is not None else c
    return cls
