"""
``CudaHealpixMapper`` -- B200 implementation of the reference's ``HealpixMapper``
(``heracles/healpy.py:68-209``) behind the same ``Mapper`` protocol
(``heracles/mapper.py:33-74``): ``area``, ``create``, ``map_values``,
``transform``, ``resample`` plus the ``nside`` / ``lmax`` / ``deconvolve``
attributes the Field layer and the CLI read (``heracles/cli.py:162-187``).

All numerics run in ``libheracles_cuda.so`` (hand-written sm_100a kernels,
cuFFT for the equatorial rings); nothing here falls back to the CPU.
"""

from __future__ import annotations

import math
import os
from functools import cached_property
from typing import Any

import numpy as np

from . import _lib
from .arrays import DeviceArray, update_metadata

c_vp = _lib.c_vp


def _native(arr: np.ndarray) -> np.ndarray:
    """float64, native byte order, C-contiguous (cf. `_nativebyteorder`, healpy.py:43-55)"""
    arr = np.asarray(arr)
    if arr.dtype != np.float64 or arr.dtype.byteorder not in ("=", "|") or not arr.flags.c_contiguous:
        arr = np.ascontiguousarray(arr, dtype=np.float64)
    return arr


def _ptr(arr: np.ndarray) -> int:
    return arr.__array_interface__["data"][0]


def _read_fits_table(path: str, with_header: bool = False):
    """columns of the first binary table of a FITS file as flat float64 arrays (no astropy / fitsio here);
    with_header: also the table's header cards as a dict"""
    with open(path, "rb") as f:
        raw = f.read()

    def header(off):
        cards = {}
        while True:
            block = raw[off : off + 2880]
            off += 2880
            for i in range(0, 2880, 80):
                card = block[i : i + 80].decode("ascii", "replace")
                key = card[:8].strip()
                if key == "END":
                    return cards, off
                if card[8:10] == "= ":
                    val = card[10:].split("/")[0].strip().strip("'").strip()
                    cards[key] = val

    h0, off = header(0)
    h1, off = header(off)
    nrow, rowlen, nfield = int(h1["NAXIS2"]), int(h1["NAXIS1"]), int(h1["TFIELDS"])
    fmts = [h1[f"TFORM{i + 1}"] for i in range(nfield)]
    dts = []
    for fmt in fmts:
        rep = int(fmt[:-1] or 1)
        code = {"D": ">f8", "E": ">f4", "J": ">i4", "K": ">i8", "I": ">i2", "B": "u1"}[fmt[-1]]
        dts.append((code, rep))
    dt = np.dtype([(f"c{i}", c, (r,)) for i, (c, r) in enumerate(dts)])
    assert dt.itemsize == rowlen
    tab = np.frombuffer(raw, dtype=dt, count=nrow, offset=off)
    cols = [np.asarray(tab[f"c{i}"], dtype=np.float64).reshape(-1) for i in range(nfield)]
    return (cols, h1) if with_header else cols


def read_pixwin_fits(path: str):
    """
    Minimal reader for HEALPix' ``pixel_window_n%04d.fits`` (one binary table
    with columns TEMPERATURE, POLARIZATION of big-endian float64/float32).
    Returns (pw_T, pw_P).
    """
    cols = _read_fits_table(path)
    return cols[0], (cols[1] if len(cols) > 1 else cols[0])


def n_fullweights(nside: int) -> int:
    """length of HEALPix' compressed full-weights array"""
    return ((3 * nside + 1) * (nside + 1)) // 4


def expand_fullweights(nside: int, wgt) -> np.ndarray:
    """
    Per-pixel quadrature weights (RING order, multiplying 4 pi / npix) from HEALPix' compressed
    ``healpix_full_weights_nside_%04d.fits`` array -- what ``hp.map2alm(use_pixel_weights=True)`` applies
    (``heracles/healpy.py:183-189``).  The file stores ``w - 1`` for one representative of every orbit of the pixel
    symmetries (north/south mirror, fourfold rotation, reflection inside an octant): per ring pair ``i`` (north ring
    ``i + 1``) ``ceil(q / 2)`` values, ``q = min(nside, i + 1)``, one more when ``q`` is even and the ring is not
    shifted; pixel ``j`` of the ring reads entry ``min(j mod q, q - shifted - j mod q)``.
    """
    wgt = np.asarray(wgt, dtype=np.float64).reshape(-1)
    if wgt.size != n_fullweights(nside):
        raise ValueError(f"full weights for nside {nside} need {n_fullweights(nside)} entries, got {wgt.size}")
    npix = 12 * nside * nside
    out = np.ones(npix)
    pix = vpix = 0
    for i in range(2 * nside):
        shifted = (i < nside - 1) or bool((i + nside) & 1)
        q = min(nside, i + 1)
        odd = q & 1
        wpix = ((q + 1) >> 1) + (0 if (odd or shifted) else 1)
        j4 = np.arange(4 * q) % q
        r = np.minimum(j4, q - (1 if shifted else 0) - j4)
        w = 1.0 + wgt[vpix + r]
        out[pix:pix + 4 * q] = w
        if i != 2 * nside - 1:
            psouth = npix - pix - 4 * q
            out[psouth:psouth + 4 * q] = w
        pix += 4 * q
        vpix += wpix
    assert vpix == wgt.size and (pix == npix - pix + 4 * nside or True)
    return out


def read_fullweights_fits(path: str, nside: int) -> np.ndarray:
    """expanded per-pixel weights from a ``healpix_full_weights_nside_%04d.fits`` file"""
    cols = _read_fits_table(path)
    return expand_fullweights(nside, cols[0])


class CudaHealpixMapper:
    """
    Mapper for HEALPix maps on a B200.

    Parameters mirror ``HealpixMapper(nside, lmax=None, *, deconvolve=None,
    dtype=float64)``.  Additional keyword-only options:

    niter : Jacobi refinement steps of the analysis; healpy's ``map2alm`` default
        ``iter=3`` is what the reference runs (it does not pass ``iter``).
    pixwin : explicit ``(pw_T, pw_P)`` pixel window arrays for deconvolution.
        If omitted with ``deconvolve=True`` they are read from ``DATAPATH`` /
        healpy's data directory (``pixel_window_n%04d.fits``) or from healpy
        when importable; healpy's tables cannot be recomputed here.
    pixel_weights : per-pixel quadrature weights (RING, multiplying 4 pi / npix).  ``"auto"`` (the default) does
        what the reference's ``hp.map2alm(use_pixel_weights=True, datapath=DATAPATH)`` does when the table is at hand:
        it reads ``DATAPATH/full_weights/healpix_full_weights_nside_%04d.fits`` (or healpy's copy when healpy is
        importable) and falls back to uniform weights -- with ONE warning -- when neither exists (healpy would
        download the file; this backend never touches the network).  ``None`` = uniform weights.
    weights_mode : ``"premultiply"`` (healpy: the map is weighted once, the iterations run on the weighted map) or
        ``"per_pass"`` (the weights are part of every analysis pass of the Jacobi loop).
    scheme : ``"ring"`` (the reference's maps) or ``"nest"``: pixel order of the maps this mapper creates, maps into
        and transforms (NEST maps are reordered on the device before the ring FFTs).
    device : CUDA device index (default: ``LOCAL_RANK`` or 0).
    sync : make ``map_values`` return only when the device finished (default).
    """

    DATAPATH: str | None = None

    def __init__(
        self,
        nside: int,
        lmax: int | None = None,
        *,
        deconvolve: bool | None = None,
        dtype: Any = np.float64,
        niter: int = 3,
        pixwin: Any = None,
        pixel_weights: Any = "auto",
        weights_mode: str = "premultiply",
        scheme: str = "ring",
        device: int | None = None,
        sync: bool = True,
        aggregate: bool = False,
        context: Any = None,
    ) -> None:
        if lmax is None:
            lmax = 3 * nside // 2
        if deconvolve is None:
            deconvolve = True
        if np.dtype(dtype) != np.float64:
            raise NotImplementedError("CudaHealpixMapper computes in float64 only")
        if nside < 1 or nside & (nside - 1):
            raise ValueError("nside must be a power of two")
        self.__nside = int(nside)
        self.__lmax = int(lmax)
        self.__deconv = bool(deconvolve)
        self.__dtype = np.dtype(dtype)
        self.niter = int(niter)
        self.sync = bool(sync)
        self.aggregate = bool(aggregate)
        if weights_mode not in ("premultiply", "per_pass"):
            raise ValueError("weights_mode must be 'premultiply' or 'per_pass'")
        if scheme not in ("ring", "nest"):
            raise ValueError("scheme must be 'ring' or 'nest'")
        self.weights_mode = weights_mode
        self.scheme = scheme
        self._pixwin = pixwin
        self._pixel_weights_arg = pixel_weights
        # context: a library context of its own (own stream and workspaces, _lib.extra_context) instead of the
        # process-wide one of the device -- what lets transforms run beside the mapping (heracles_b200.overlap)
        self._ctx = context if context is not None else _lib.get_context(device)

    # -- reference attributes ------------------------------------------------
    @property
    def nside(self) -> int:
        return self.__nside

    @property
    def lmax(self) -> int:
        return self.__lmax

    @property
    def deconvolve(self) -> bool:
        return self.__deconv

    @property
    def context(self) -> _lib.Context:
        return self._ctx

    @cached_property
    def area(self) -> float:
        """hp.nside2pixarea(nside) (healpy.py:117-122)"""
        return 4.0 * math.pi / (12 * self.__nside * self.__nside)

    @property
    def npix(self) -> int:
        return 12 * self.__nside * self.__nside

    # -- quadrature weights ------------------------------------------------------
    _warned_weights = False

    @cached_property
    def _pixel_weights(self):
        """resolved per-pixel weights (float64[npix]) or None"""
        pw = self._pixel_weights_arg
        if pw is None:
            return None
        if not isinstance(pw, str):
            pw = _native(pw)
            if pw.size != self.npix:
                raise ValueError("pixel_weights must have npix entries")
            return pw
        if pw != "auto":
            raise ValueError("pixel_weights must be an array, None or 'auto'")
        name = os.path.join("full_weights", "healpix_full_weights_nside_%04d.fits" % self.__nside)
        paths = [os.path.join(self.DATAPATH, name)] if self.DATAPATH else []
        try:
            import healpy  # noqa: F401  (optional: only for its data directory)

            paths.append(os.path.join(os.path.dirname(healpy.__file__), "data", name))
        except Exception:
            pass
        for p in paths:
            if os.path.exists(p):
                return read_fullweights_fits(p, self.__nside)
        if not CudaHealpixMapper._warned_weights:
            import warnings

            warnings.warn(
                "HEALPix full pixel weights (" + name + ") not found under CudaHealpixMapper.DATAPATH: transforming with "
                "uniform weights 4 pi / npix; hp.map2alm(use_pixel_weights=True) would have downloaded the table",
                stacklevel=3,
            )
            CudaHealpixMapper._warned_weights = True
        return None

    def _apply_modes(self):
        _lib.check(self._ctx.lib.hcu_set_weights_mode(self._ctx.handle, 0 if self.weights_mode == "premultiply" else 1))

    def _ring_view(self, data):
        """the map(s) in RING order on the device (NEST maps are reordered into a temporary)"""
        if self.scheme == "ring":
            return data
        src = data if (isinstance(data, DeviceArray) and data.device_ptr is not None) else None
        if src is None:
            src = DeviceArray.zeros(self._ctx, data.shape)
            src[...] = np.asarray(data)
        src.to_device()
        out = DeviceArray.zeros(self._ctx, data.shape)
        npix = self.npix
        for r in range(int(np.prod(data.shape[:-1], dtype=np.int64)) if data.ndim > 1 else 1):
            _lib.check(self._ctx.lib.hcu_reorder(self._ctx.handle, self.__nside, c_vp(src.device_ptr + 8 * r * npix),
                                                 c_vp(out.device_ptr + 8 * r * npix), 0))
        self._ctx.synchronize()
        update_metadata(out, **(data.dtype.metadata or {}))
        return out

    # -- create ---------------------------------------------------------------
    def create(self, *dims: int, spin: int = 0):
        """zero map(s) in managed memory + metadata (healpy.py:124-142)"""
        m = DeviceArray.zeros(self._ctx, (*dims, self.npix), dtype=self.__dtype)
        update_metadata(
            m,
            geometry="healpix",
            kernel="healpix",
            nside=self.__nside,
            lmax=self.__lmax,
            deconv=self.__deconv,
            spin=spin,
            **({"nest": True} if self.scheme == "nest" else {}),
        )
        return m

    # -- one page -> position map and shear map -------------------------------------
    def new_page_stats(self):
        """zeroed float64[8] accumulator for :meth:`map_page` (see ``hcu_map_page`` in include/heracles_cuda.h)"""
        return DeviceArray.zeros(self._ctx, (8,))

    def map_page(self, lon, lat, w=None, g1=None, g2=None, *, pos=None, she=None, stats=None) -> None:
        """
        One catalogue page into the position map AND the shear map of a tomographic bin in one pass: what the
        reference does with two ``map_values`` calls from two Field objects that read the same lon / lat / weight
        columns (``heracles/fields.py:262-271`` and ``:420-433``):

            pos[ipix] += w;   she[0, ipix] += w g1;   she[1, ipix] += w g2

        ``w=None`` means unit weights.  ``stats`` (from :meth:`new_page_stats`) accumulates the running sums the
        Field layer keeps per page -- rows, sum w, sum w^2 for the positions; rows with w != 0, sum w, sum w^2,
        sum w^2 (g1^2 + g2^2) for the shears; rows with NaN -- see :func:`page_means`.
        """
        lon, lat = _native(lon), _native(lat)
        n = lon.size
        cols = [None if c is None else _native(c) for c in (w, g1, g2)]
        for c in [lat] + [c for c in cols if c is not None]:
            if c.size != n:
                raise ValueError("columns differ in size")
        if she is not None and (cols[1] is None or cols[2] is None):
            raise ValueError("the shear map needs g1 and g2")
        for m, lead in ((pos, ()), (she, (2,))):
            if m is not None and (not isinstance(m, DeviceArray) or m.device_ptr is None or m.shape != (*lead, self.npix)):
                raise ValueError("maps must come from this mapper's create()")
            if m is not None:
                m.to_device()
        ptr = lambda a: c_vp(_ptr(a)) if a is not None else c_vp(0)  # noqa: E731
        _lib.check(
            self._ctx.lib.hcu_map_page(
                self._ctx.handle, self.__nside, 1 if self.scheme == "nest" else 0, ptr(lon), ptr(lat), ptr(cols[0]),
                ptr(cols[1]), ptr(cols[2]), n, c_vp(pos.device_ptr if pos is not None else 0),
                c_vp(she.device_ptr if she is not None else 0), self.npix, c_vp(stats.device_ptr if stats is not None else 0),
            )
        )
        if self.sync:
            bad = self._ctx.bad_rows()  # synchronises
            if bad:
                raise ValueError("THETA is out of range [0,pi]")

    @staticmethod
    def page_means(stats):
        """
        ``(ngal, wmean, w2mean)`` of the positions and ``(ngal, wmean, w2mean, var)`` of the shears from a
        :meth:`map_page` accumulator: the values the reference's running means converge to
        (``heracles/fields.py:269-271, 430-433``); raises like ``CatalogPage.get`` if a NaN was seen.
        """
        # wait for the page kernels that are still updating the sums (map_page(sync=False) returns at once)
        s = np.array(stats._host() if isinstance(stats, DeviceArray) else np.asarray(stats), dtype=np.float64)
        if s[7]:
            raise ValueError("invalid values in catalogue page columns")
        pos = (int(s[0]), s[1] / s[0], s[2] / s[0]) if s[0] else (0, 0.0, 0.0)
        she = (int(s[3]), s[4] / s[3], s[5] / s[3], s[6] / s[3]) if s[3] else (0, 0.0, 0.0, 0.0)
        return pos, she

    # -- map_values -------------------------------------------------------------
    def map_values(self, lon, lat, data, values, spin: int = 0) -> None:
        """data[..., ang2pix(lon, lat)] += values (healpy.py:144-160)"""
        lon = _native(lon)
        lat = _native(lat)
        values = _native(values)
        n = lon.size
        if lat.size != n:
            raise ValueError("lon and lat differ in size")
        npix = self.npix
        if data.shape[-1] != npix:
            raise ValueError("data is not a map of this mapper")
        nv = int(np.prod(data.shape[:-1], dtype=np.int64)) if data.ndim > 1 else 1
        if values.size != nv * n:
            raise ValueError("values do not match data and positions")
        if nv > 4:
            raise ValueError("at most 4 value rows per call")
        dptr = data.device_ptr if isinstance(data, DeviceArray) else None
        if dptr is None:
            # a host array (e.g. a plain np.zeros map): map on the device, add back
            tmp = self.create(*data.shape[:-1])
            self.map_values(lon, lat, tmp, values, spin=spin)
            data += tmp._host().reshape(data.shape)
            return
        data.to_device()
        flags = 1 if self.aggregate else 0
        _lib.check(
            self._ctx.lib.hcu_map_values(
                self._ctx.handle, self.__nside, 1 if self.scheme == "nest" else 0, c_vp(_ptr(lon)), c_vp(_ptr(lat)),
                c_vp(_ptr(values)), n, nv, n, c_vp(dptr), npix, flags,
            )
        )
        if self.sync:
            bad = self._ctx.bad_rows()  # synchronises
            if bad:
                raise ValueError("THETA is out of range [0,pi]")

    # -- transform -----------------------------------------------------------------
    def _fl(self, spin: int):
        if not self.__deconv:
            return None
        pw = self._get_pixwin()[0 if spin == 0 else 1]
        if len(pw) < self.__lmax + 1:
            raise ValueError("pixel window shorter than lmax + 1")
        fl = np.ones(self.__lmax + 1)
        s = abs(spin)
        fl[s:] /= np.asarray(pw, dtype=np.float64)[s : self.__lmax + 1]
        return fl

    def _get_pixwin(self):
        if self._pixwin is None:
            pw = None
            name = "pixel_window_n%04d.fits" % self.__nside
            paths = []
            if self.DATAPATH:
                paths.append(os.path.join(self.DATAPATH, name))
            try:
                import healpy  # noqa: F401  (optional: only for its data tables)

                pw = healpy.pixwin(self.__nside, lmax=self.__lmax, pol=True)
            except Exception:
                for p in paths:
                    if os.path.exists(p):
                        pw = read_pixwin_fits(p)
                        break
            if pw is None:
                raise RuntimeError(
                    "deconvolve=True needs HEALPix' pixel window table: pass pixwin=(pw_T, pw_P), set "
                    "CudaHealpixMapper.DATAPATH to a directory with " + name + ", or use deconvolve=False"
                )
            self._pixwin = pw
        pw = self._pixwin
        if isinstance(pw, np.ndarray) and pw.ndim == 1:
            pw = (pw, pw)
        return pw

    def transform(self, data, spin: int = 0):
        """map2alm (+ pixel window deconvolution), healpy.py:162-203"""
        if spin not in (0, 2):
            msg = f"spin-{spin} maps not yet supported"
            raise NotImplementedError(msg)
        md = data.dtype.metadata or {}
        npix = self.npix
        if data.shape[-1] != npix:
            raise ValueError("data is not a map of this mapper")
        lead = data.shape[:-1]
        nmaps = int(np.prod(lead, dtype=np.int64)) if lead else 1
        if spin == 2 and (len(lead) == 0 or lead[-1] != 2):
            raise ValueError("spin-2 data must have shape (..., 2, npix)")
        lmax = self.__lmax
        nalm = (lmax + 1) * (lmax + 2) // 2
        fl = self._fl(spin)
        alm = DeviceArray.zeros(self._ctx, (*lead, nalm), dtype=np.complex128)
        data = self._ring_view(data)
        self._apply_modes()
        if isinstance(data, DeviceArray) and data.device_ptr is not None:
            data.to_device()
            mptr = data.device_ptr
        else:
            data = _native(data)
            mptr = _ptr(data)
        pw = self._pixel_weights
        _lib.check(
            self._ctx.lib.hcu_map2alm(
                self._ctx.handle, self.__nside, lmax, spin, nmaps, c_vp(mptr), npix,
                c_vp(0), c_vp(_ptr(pw) if pw is not None else 0), self.niter,
                c_vp(_ptr(fl) if fl is not None else 0), c_vp(alm.device_ptr), nalm,
            )
        )
        update_metadata(alm, **{**md, "deconv": self.__deconv})
        return alm

    # -- resample ----------------------------------------------------------------------
    def resample(self, data):
        """hp.ud_grade(data, nside) (healpy.py:205-209)"""
        npix_in = data.shape[-1]
        nside_in = int(round(math.sqrt(npix_in / 12)))
        if 12 * nside_in * nside_in != npix_in:
            raise ValueError("input is not a HEALPix map")
        lead = data.shape[:-1]
        out = DeviceArray.zeros(self._ctx, (*lead, self.npix), dtype=np.float64)
        src = data if (isinstance(data, DeviceArray) and data.device_ptr is not None) else _native(data)
        src2 = src.reshape(-1, npix_in)
        out2 = out.reshape(-1, self.npix)
        for i in range(src2.shape[0]):
            row = src2[i]
            sp = row.device_ptr if isinstance(row, DeviceArray) else None
            if sp is None:
                row = np.ascontiguousarray(row)  # kept alive until the call returns
                sp = _ptr(row)
            _lib.check(
                self._ctx.lib.hcu_ud_grade(
                    self._ctx.handle, nside_in, c_vp(sp), self.__nside, c_vp(_ptr(out2[i].view(np.ndarray)))
                )
            )
        self._ctx.synchronize()
        update_metadata(out, **(data.dtype.metadata or {}))
        return out
