"""
Visibility / mask maps for this backend: ``heracles.io.read_vmap`` (``heracles/io.py:360-381``) on the device.

The reference reads a HEALPix map with healpy, zeroes the UNSEEN pixels, ``hp.ud_grade``s it to the requested
resolution and, with ``transform=True``, returns ``hp.almxfl(hp.map2alm(vmap, lmax=lmax, use_pixel_weights=True),
1 / hp.pixwin(nside, lmax=lmax))`` -- the mask alm the mixing matrices are computed from, at nside 4096 / 8192 in
``examples/heracles.cfg:54-65``.  Here the file is parsed with the package's own minimal FITS reader (neither healpy
nor fitsio is needed), NEST files are reordered, the resolution change is ``hcu_ud_grade`` and the transform is
``CudaHealpixMapper.transform`` with ``deconvolve=True``: the same kernels as the catalogue maps, up to nside 8192.
"""

from __future__ import annotations

import warnings

import numpy as np

from .mapper import CudaHealpixMapper, _read_fits_table

__all__ = ["UNSEEN", "read_map", "read_vmap", "vmap_alm"]

UNSEEN = -1.6375e30  # healpy.UNSEEN


def read_map(filename, field=0):
    """``hp.read_map(filename, field=field, dtype=float)``: column `field` of the first binary table of a HEALPix map
    file as a float64 RING map (a file with ORDERING = NESTED is reordered on the device)"""
    cols, hdr = _read_fits_table(filename, with_header=True)
    if not 0 <= field < len(cols):
        raise IndexError(f"{filename}: no map field {field}")
    if str(hdr.get("INDXSCHM", "IMPLICIT")).upper() == "EXPLICIT":
        raise NotImplementedError(f"{filename}: partial-sky (EXPLICIT index) map files are not supported")
    m = np.ascontiguousarray(cols[field], dtype=np.float64)
    nside = int(round((m.size / 12) ** 0.5))
    if 12 * nside * nside != m.size:
        raise ValueError(f"{filename}: {m.size} values are not a HEALPix map")
    if "NSIDE" in hdr and int(hdr["NSIDE"]) != nside:
        raise ValueError(f"{filename}: NSIDE = {hdr['NSIDE']} does not match {m.size} pixels")
    if str(hdr.get("ORDERING", "RING")).upper().startswith("NEST"):
        m = np.array(CudaHealpixMapper(nside, 0, deconvolve=False, scheme="nest", pixel_weights=None)._ring_view(m), dtype=np.float64)
    return m


def vmap_alm(vmap, lmax=None, *, pixwin=None, pixel_weights="auto", niter=3, device=None):
    """mask alm as ``read_vmap(transform=True)`` makes them: ``map2alm`` (pixel weights as the mapper has them, healpy's
    default iterations) divided by the spin-0 pixel window.  `pixwin`: ``(pw_T, pw_P)`` or one array; default: HEALPix'
    table under ``CudaHealpixMapper.DATAPATH`` / healpy"""
    vmap = np.asarray(vmap) if not hasattr(vmap, "device_ptr") else vmap
    nside = int(round((vmap.shape[-1] / 12) ** 0.5))
    if lmax is None:
        lmax = 3 * nside - 1  # healpy's map2alm default
    mapper = CudaHealpixMapper(nside, lmax, deconvolve=True, pixwin=pixwin, pixel_weights=pixel_weights, niter=niter, device=device)
    return mapper.transform(vmap, spin=0)


def read_vmap(filename, nside=None, field=0, *, transform=False, lmax=None, pixwin=None, pixel_weights="auto", device=None):
    """read visibility map from a HEALPix map file -- signature and behaviour of ``heracles.io.read_vmap``"""
    vmap = read_map(filename, field=field)
    # set unseen pixels to zero
    vmap[vmap == UNSEEN] = 0.0
    nside_in = int(round((vmap.size / 12) ** 0.5))
    if nside is not None and nside != nside_in:
        # vmap is provided at a different resolution
        warnings.warn(f"{filename}: changing NSIDE to {nside}", stacklevel=2)
        vmap = CudaHealpixMapper(nside, 0, deconvolve=False, pixel_weights=None, device=device).resample(vmap)
    if transform:
        vmap = vmap_alm(vmap, lmax, pixwin=pixwin, pixel_weights=pixel_weights, device=device)
    return vmap
