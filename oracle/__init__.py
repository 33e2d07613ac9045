"""
oracle -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import this package; ``heracles_b200`` (the
product) never does and fails loudly without its CUDA library.

Parity status (see healpix_oracle.c header): ang2pix and map2alm are
restatements of the published HEALPix algorithms that `healpy` (third-party,
unpinned in /root/reference/pyproject.toml:27, not installable here) wraps --
"parity unpinned" against healpy itself; pinned against independent closed
forms in tests/test_oracle_*.py.  alm2cl is pinned against the reference's own
``heracles/twopoint.py:63-101`` through tests/golden/.

Python side: ring FFTs with scipy.fft; everything else in liboracle.so.
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_dbl = ctypes.c_double
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int64)


def build(force: bool = False) -> str:
    """compile liboracle.so with gcc (Makefile in this directory)"""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("healpix_oracle.c", "legendre_core.inc")]
    if (
        force
        or not os.path.exists(so)
        or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        L.orc_ang2pix.restype = c_i64
        L.orc_ang2pix.argtypes = [c_i64, c_int, c_dbl, c_dbl]
        L.orc_ang2pix_lonlat.restype = c_i64
        L.orc_ang2pix_lonlat.argtypes = [c_i64, c_int, c_i64, _dp, _dp, _ip]
        L.orc_scatter_add.restype = None
        L.orc_scatter_add.argtypes = [c_i64, _ip, c_int, _dp, c_i64, _dp, c_i64]
        L.orc_ring2nest.argtypes = [c_i64, c_i64, _ip, _ip]
        L.orc_nest2ring.argtypes = [c_i64, c_i64, _ip, _ip]
        L.orc_pix2ang_lonlat.argtypes = [c_i64, c_int, c_i64, _ip, _dp, _dp]
        L.orc_ring_info.argtypes = [c_i64, c_i64, _ip, _ip, _dp, _dp, _dp]
        L.orc_phase2alm.argtypes = [c_int, c_i64, c_int, c_int, c_int, _dp, _dp]
        L.orc_alm2phase.argtypes = [c_int, c_i64, c_int, c_int, c_int, _dp, _dp]
        L.orc_lambda.argtypes = [c_int, c_int, c_int, c_int, c_dbl, c_dbl, _dp]
        L.orc_alm2cl.argtypes = [c_int, c_int, c_int, _dp, _dp, _dp]
        L.orc_num_threads.restype = c_int
        L.orc_set_num_threads.argtypes = [c_int]
        _LIB = L
    return _LIB


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


# ---------------------------------------------------------------------------
# pixelisation  (hp.ang2pix, heracles/healpy.py:157)
# ---------------------------------------------------------------------------


def nside2npix(nside: int) -> int:
    return 12 * nside * nside


def ang2pix(nside, lon, lat, nest=False):
    """healpy.ang2pix(nside, lon, lat, nest=nest, lonlat=True)"""
    lon = np.ascontiguousarray(lon, dtype=np.float64)
    lat = np.ascontiguousarray(lat, dtype=np.float64)
    out = np.empty(lon.shape, dtype=np.int64)
    bad = lib().orc_ang2pix_lonlat(nside, int(nest), lon.size, _d(lon), _d(lat), _i(out))
    if bad:
        raise ValueError("THETA is out of range [0,pi]")
    return out


def pix2ang(nside, ipix, nest=False):
    """healpy.pix2ang(nside, ipix, nest=nest, lonlat=True) -> (lon, lat)"""
    ipix = np.ascontiguousarray(ipix, dtype=np.int64)
    lon = np.empty(ipix.shape)
    lat = np.empty(ipix.shape)
    lib().orc_pix2ang_lonlat(nside, int(nest), ipix.size, _i(ipix), _d(lon), _d(lat))
    return lon, lat


def ring2nest(nside, ipix):
    ipix = np.ascontiguousarray(ipix, dtype=np.int64)
    out = np.empty_like(ipix)
    lib().orc_ring2nest(nside, ipix.size, _i(ipix), _i(out))
    return out


def nest2ring(nside, ipix):
    ipix = np.ascontiguousarray(ipix, dtype=np.int64)
    out = np.empty_like(ipix)
    lib().orc_nest2ring(nside, ipix.size, _i(ipix), _i(out))
    return out


def map_values(nside, lon, lat, data, values, nest=False):
    """HealpixMapper.map_values (heracles/healpy.py:144-160): ang2pix + `_map`"""
    ipix = ang2pix(nside, lon, lat, nest=nest)
    values = np.ascontiguousarray(values, dtype=np.float64)
    assert data.flags.c_contiguous and data.dtype == np.float64
    npix = data.shape[-1]
    nv = data.size // npix
    v2 = values.reshape(nv, -1)
    lib().orc_scatter_add(ipix.size, _i(ipix), nv, _d(v2), v2.shape[1], _d(data), npix)


def ring_info(nside, iring):
    """(startpix, ringpix, cos theta, sin theta, phi0) for ring iring in 1..4nside-1"""
    s, n = c_i64(), c_i64()
    c, sn, p0 = c_dbl(), c_dbl(), c_dbl()
    lib().orc_ring_info(
        nside, iring, ctypes.byref(s), ctypes.byref(n), ctypes.byref(c), ctypes.byref(sn), ctypes.byref(p0)
    )
    return s.value, n.value, c.value, sn.value, p0.value


def ring_table(nside):
    nr = 4 * nside - 1
    start = np.empty(nr, np.int64)
    npx = np.empty(nr, np.int64)
    cth = np.empty(nr)
    sth = np.empty(nr)
    phi0 = np.empty(nr)
    for i in range(nr):
        start[i], npx[i], cth[i], sth[i], phi0[i] = ring_info(nside, i + 1)
    return start, npx, cth, sth, phi0


# ---------------------------------------------------------------------------
# spherical harmonic transforms (hp.map2alm, heracles/healpy.py:183-189)
# ---------------------------------------------------------------------------


def nalm(lmax: int) -> int:
    return (lmax + 1) * (lmax + 2) // 2


def almidx(lmax, l, m):
    return m * (2 * lmax + 1 - m) // 2 + l


def _fft_workers():
    return max(1, len(os.sched_getaffinity(0)))


def map2phase(nside, lmax, maps, ring_weights=None):
    """ring FFT stage: phase[c, ring, m] = w_r e^{-i m phi0} sum_j f_j e^{-2 pi i m j / n_r}"""
    import scipy.fft as sfft

    maps = np.atleast_2d(np.asarray(maps, dtype=np.float64))
    ncomp, npix = maps.shape
    assert npix == 12 * nside * nside
    start, npx, _, _, phi0 = ring_table(nside)
    nr = 4 * nside - 1
    w = np.full(nr, 4 * np.pi / npix)
    if ring_weights is not None:
        w = w * np.asarray(ring_weights, dtype=np.float64)
    m = np.arange(lmax + 1)
    phase = np.empty((ncomp, nr, lmax + 1), dtype=np.complex128)
    # equatorial belt in one batched FFT
    n4 = 4 * nside
    ncap = 2 * nside * (nside - 1)
    belt = maps[:, ncap : npix - ncap].reshape(ncomp, 2 * nside + 1, n4)
    X = sfft.rfft(belt, axis=-1, workers=_fft_workers())
    k = m % n4
    kk = np.where(k <= n4 // 2, k, n4 - k)
    Xm = X[..., kk]
    Xm = np.where(k <= n4 // 2, Xm, np.conj(Xm))
    r0 = nside - 1
    ph = np.exp(-1j * np.outer(phi0[r0 : r0 + 2 * nside + 1], m))
    phase[:, r0 : r0 + 2 * nside + 1, :] = Xm * (ph * w[r0 : r0 + 2 * nside + 1, None])
    # polar caps, one FFT per ring
    for r in list(range(nside - 1)) + list(range(3 * nside, nr)):
        n = int(npx[r])
        X = sfft.rfft(maps[:, start[r] : start[r] + n], axis=-1)
        k = m % n
        kk = np.where(k <= n // 2, k, n - k)
        Xm = X[:, kk]
        Xm = np.where(k <= n // 2, Xm, np.conj(Xm))
        phase[:, r, :] = Xm * (np.exp(-1j * m * phi0[r]) * w[r])
    return phase


def phase2map(nside, lmax, phase):
    """inverse ring FFT stage: f(r, j) = Re sum_m (2 - delta_m0) b_m(r) e^{i m phi_j}"""
    import scipy.fft as sfft

    ncomp = phase.shape[0]
    npix = 12 * nside * nside
    start, npx, _, _, phi0 = ring_table(nside)
    nr = 4 * nside - 1
    m = np.arange(lmax + 1)
    maps = np.empty((ncomp, npix))
    for r in range(nr):
        n = int(npx[r])
        c = phase[:, r, :] * np.exp(1j * m * phi0[r])
        G = np.zeros((ncomp, n), dtype=np.complex128)
        np.add.at(G, (slice(None), m % n), c)
        np.add.at(G, (slice(None), (-m[1:]) % n), np.conj(c[:, 1:]))
        maps[:, start[r] : start[r] + n] = sfft.ifft(G, axis=-1).real * n
    return maps


def phase2alm(nside, lmax, phase, spin=0, prec=0):
    phase = np.ascontiguousarray(phase, dtype=np.complex128)
    ncomp = phase.shape[0]
    alm = np.empty((ncomp, nalm(lmax)), dtype=np.complex128)
    lib().orc_phase2alm(prec, nside, lmax, spin, ncomp, _d(phase.view(np.float64)), _d(alm.view(np.float64)))
    return alm


def alm2phase(nside, lmax, alm, spin=0, prec=0):
    alm = np.ascontiguousarray(np.atleast_2d(alm), dtype=np.complex128)
    ncomp = alm.shape[0]
    phase = np.empty((ncomp, 4 * nside - 1, lmax + 1), dtype=np.complex128)
    lib().orc_alm2phase(prec, nside, lmax, spin, ncomp, _d(alm.view(np.float64)), _d(phase.view(np.float64)))
    return phase


def alm2map(nside, lmax, alm, spin=0, prec=0):
    """hp.alm2map restated; spin 2: alm rows (E, B) -> maps rows (Q, U)"""
    return phase2map(nside, lmax, alm2phase(nside, lmax, alm, spin=spin, prec=prec))


def map2alm(nside, lmax, maps, spin=0, niter=0, ring_weights=None, pixel_weights=None, prec=0):
    """
    hp.map2alm restated (heracles/healpy.py:183-189).

    maps: (npix,) or (k, npix).  spin 0: k independent scalar maps.  spin 2: k
    must be even, rows are (Q, U) pairs, output rows are (E, B) pairs (the
    reference prepends a zero T map and drops the T alm again,
    healpy.py:174-178,198-199; that is a no-op numerically).

    niter: Jacobi refinement steps alm += A(map - S(alm)) as in HEALPix'
    map2alm_iter (healpy's default iter=3).  pixel_weights multiply the map
    (healpy use_pixel_weights=True reads them from a data file that is not
    available here; None = uniform 4 pi / npix quadrature).
    """
    maps = np.asarray(maps, dtype=np.float64)
    single = maps.ndim == 1
    maps = np.atleast_2d(maps)

    def analysis(mm):
        if pixel_weights is not None:
            mm = mm * pixel_weights
        return phase2alm(nside, lmax, map2phase(nside, lmax, mm, ring_weights), spin=spin, prec=prec)

    alm = analysis(maps)
    for _ in range(niter):
        resid = maps - alm2map(nside, lmax, alm, spin=spin, prec=prec)
        alm = alm + analysis(resid)
    return alm[0] if single else alm


def almxfl(alm, fl):
    """hp.almxfl (heracles/healpy.py:195): alm[l, m] * fl[l]"""
    alm = np.array(alm, dtype=np.complex128, copy=True)
    lmax = alm2lmax(alm)
    for m in range(lmax + 1):
        s = almidx(lmax, m, m)
        alm[..., s : s + lmax - m + 1] *= fl[m : lmax + 1]
    return alm


def lambda_lm(lmax, m, spin, cth, sth, prec=0):
    out = np.zeros(lmax + 1)
    lib().orc_lambda(prec, lmax, m, spin, cth, sth, _d(out))
    return out


# ---------------------------------------------------------------------------
# alm2cl (heracles/twopoint.py:55-101)
# ---------------------------------------------------------------------------


def alm2lmax(alm):
    return (int((8 * np.shape(alm)[-1] + 1) ** 0.5 + 0.01) - 3) // 2


def alm2cl(alm, alm2=None, *, lmax=None):
    """twopoint.alm2cl restated in C with the same running-mean update"""
    if alm2 is None:
        alm2 = alm
    alm = np.ascontiguousarray(alm, dtype=np.complex128)
    alm2 = np.ascontiguousarray(alm2, dtype=np.complex128)
    l1, l2 = alm2lmax(alm), alm2lmax(alm2)
    if lmax is None:
        lmax = step = min(l1, l2)
    else:
        step = min(lmax, l1, l2)
    d1, d2 = alm.shape[:-1], alm2.shape[:-1]
    a = alm.reshape(-1, alm.shape[-1])
    b = alm2.reshape(-1, alm2.shape[-1])
    cl = np.empty((a.shape[0], b.shape[0], step + 1))
    for i in range(a.shape[0]):
        for j in range(b.shape[0]):
            lib().orc_alm2cl(l1, l2, lmax, _d(a[i].view(np.float64)), _d(b[j].view(np.float64)), _d(cl[i, j]))
    return cl.reshape(*d1, *d2, step + 1)
