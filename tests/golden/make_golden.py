"""
Generate the committed golden fixtures under tests/golden/.

Run in the BUILD container only (needs /root/reference, scipy, mpmath):

    python tests/golden/make_golden.py

1. alm2cl_*.npz  -- inputs and outputs of the reference's own
   heracles.twopoint.alm2cl (heracles/twopoint.py:63-101), imported through a
   stub package because `import heracles` needs fitsio/healpy
   (heracles/__init__.py:87).  Shapes follow tests/test_twopoint.py:24-88.
2. sht_direct_nside4.npz -- spin-0 and spin-2 analysis of random nside=4 maps
   by DIRECT summation over pixels with scipy.special.sph_harm_y and
   explicit Wigner-d spin-2 harmonics (mpmath) -- independent of any
   recursion or FFT; pixel centres from the published HEALPix ring formulae.
3. ang2pix_cases.npz -- hand-derivable pixel indices at nside=1,2 and
   adversarial points (poles, cap/belt boundary, lon wrap) whose expected
   values follow from the HEALPix pixel layout, not from code under test.
"""

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference_twopoint():
    pkg = types.ModuleType("heracles")
    pkg.__path__ = [os.path.join(REF, "heracles")]
    sys.modules["heracles"] = pkg
    import importlib

    return importlib.import_module("heracles.twopoint")


def make_alm2cl():
    tp = import_reference_twopoint()
    rng = np.random.default_rng(50)  # tests/conftest.py:21
    lmax = 32
    size = (lmax + 1) * (lmax + 2) // 2
    pos = rng.standard_normal((size, 2)) @ [1, 1j]
    she = rng.standard_normal((2, size, 2)) @ [1, 1j]
    pos2 = rng.standard_normal((size, 2)) @ [1, 1j]
    she2 = rng.standard_normal((2, size, 2)) @ [1, 1j]
    out = dict(pos=pos, she=she, pos2=pos2, she2=she2)
    out["cl_pos_pos"] = tp.alm2cl(pos, pos)
    out["cl_pos_pos2"] = tp.alm2cl(pos, pos2)
    out["cl_pos_she"] = tp.alm2cl(pos, she)
    out["cl_she_she"] = tp.alm2cl(she, she)
    out["cl_she_she2"] = tp.alm2cl(she, she2)
    out["cl_pos_pos_lmax20"] = tp.alm2cl(pos, pos2, lmax=20)
    # unequal sizes, tests/test_twopoint.py:68-88
    l1, l2 = 10, 20
    a = rng.standard_normal(((l1 + 1) * (l1 + 2) // 2, 2)) @ [1, 1j]
    b = rng.standard_normal(((l2 + 1) * (l2 + 2) // 2, 2)) @ [1, 1j]
    out["ua"], out["ub"] = a, b
    out["cl_u"] = tp.alm2cl(a, b)
    out["cl_u_lmax20"] = tp.alm2cl(a, b, lmax=l2)
    np.savez(os.path.join(HERE, "alm2cl_reference.npz"), **out)
    print("alm2cl_reference.npz written")


def pix_centres(nside):
    """(theta, phi) of RING pixel centres from the HEALPix primer formulae"""
    npix = 12 * nside * nside
    th, ph = np.empty(npix), np.empty(npix)
    p = 0
    for i in range(1, 4 * nside):
        if i < nside:
            n, z = 4 * i, 1 - i * i / (3 * nside * nside)
            phis = (np.arange(n) + 0.5) * np.pi / (2 * i)
        elif i <= 3 * nside:
            n, z = 4 * nside, 4 / 3 - 2 * i / (3 * nside)
            s = (i - nside + 1) % 2
            phis = (np.arange(n) + s / 2) * np.pi / (2 * nside)
        else:
            j = 4 * nside - i
            n, z = 4 * j, -(1 - j * j / (3 * nside * nside))
            phis = (np.arange(n) + 0.5) * np.pi / (2 * j)
        th[p : p + n] = np.arccos(z)
        ph[p : p + n] = phis
        p += n
    assert p == npix
    return th, ph


def make_sht_direct():
    import mpmath as mp
    from scipy.special import sph_harm_y

    mp.mp.dps = 30

    def wd(j, mp_, m, beta):
        f = mp.factorial
        s = mp.mpf(0)
        pref = mp.sqrt(f(j + mp_) * f(j - mp_) * f(j + m) * f(j - m))
        for k in range(0, 2 * j + 1):
            a, b, c = j + m - k, j - k - mp_, k - m + mp_
            if a < 0 or b < 0 or c < 0:
                continue
            s += (
                (-1) ** (k - m + mp_)
                / (f(a) * f(k) * f(b) * f(c))
                * mp.cos(beta / 2) ** (2 * j - 2 * k + m - mp_)
                * mp.sin(beta / 2) ** (2 * k - m + mp_)
            )
        return pref * s

    def slam(s, l, m, th):
        return (-1) ** m * mp.sqrt((2 * l + 1) / (4 * mp.pi)) * wd(l, -m, s, th)

    nside, lmax = 4, 8
    npix = 12 * nside * nside
    th, ph = pix_centres(nside)
    rng = np.random.default_rng(51)
    T, Q, U = rng.standard_normal((3, npix))
    w = 4 * np.pi / npix
    nalm = (lmax + 1) * (lmax + 2) // 2
    aT = np.zeros(nalm, complex)
    aE = np.zeros(nalm, complex)
    aB = np.zeros(nalm, complex)
    uth = np.unique(th)
    for l in range(lmax + 1):
        for m in range(l + 1):
            i = m * (2 * lmax + 1 - m) // 2 + l
            aT[i] = w * np.sum(T * np.conj(sph_harm_y(l, m, th, ph)))
            if l >= 2:
                lp = {t: float(slam(2, l, m, mp.mpf(t))) for t in uth}
                lm = {t: float(slam(-2, l, m, mp.mpf(t))) for t in uth}
                Yp = np.array([lp[t] for t in th]) * np.exp(1j * m * ph)
                Ym = np.array([lm[t] for t in th]) * np.exp(1j * m * ph)
                a2 = w * np.sum((Q + 1j * U) * np.conj(Yp))
                am2 = w * np.sum((Q - 1j * U) * np.conj(Ym))
                aE[i] = -(a2 + am2) / 2
                aB[i] = 1j * (a2 - am2) / 2
    np.savez(
        os.path.join(HERE, "sht_direct_nside4.npz"),
        nside=nside, lmax=lmax, T=T, Q=Q, U=U, aT=aT, aE=aE, aB=aB, theta=th, phi=ph,
    )
    print("sht_direct_nside4.npz written")


def make_ang2pix_cases():
    # nside=1: 12 base pixels.  RING order: ring 1 = pixels 0..3 (z>2/3),
    # ring 2 = 4..7 (equator, centred on lon = 0, 90, 180, 270), ring 3 = 8..11.
    lon, lat, nside, ring = [], [], [], []

    def add(ns, lo, la, pix):
        nside.append(ns), lon.append(lo), lat.append(la), ring.append(pix)

    for q in range(4):
        add(1, 45.0 + 90 * q, 60.0, q)  # north cap faces
        add(1, 45.0 + 90 * q, -60.0, 8 + q)  # south cap faces
        add(1, 90.0 * q + 1.0, 0.0, 4 + q)  # equatorial faces centred on 0,90,...
        add(1, 90.0 * q - 1.0 + (360 if q == 0 else 0), 0.0, 4 + q)
    add(1, 0.0, 90.0, 0)  # north pole
    add(1, 0.0, -90.0, 8)  # south pole
    add(1, 360.0 + 45.0, 60.0, 0)  # lon wrap
    add(1, -315.0, 60.0, 0)
    add(1, 720.0 + 135.0, -60.0, 9)
    # nside=2: ring 1 has 4 pixels (0..3), ring 2 has 8 (4..11), rings 3-5 have 8
    # each (12..19, 20..27, 28..35), ring 6 has 8 (36..43), ring 7 has 4 (44..47)
    for q in range(4):
        add(2, 45.0 + 90 * q, 80.0, q)
        add(2, 45.0 + 90 * q, -80.0, 44 + q)
    # ring 4 (equator) of nside=2: z=0, phi_j = j*pi/4 (shifted ring: s = (4-2+1)%2 = 1 -> (j+0.5)pi/4)
    for j in range(8):
        add(2, (j + 0.5) * 45.0, 0.0, 20 + j)
    # ring 2 of nside 2: z = 1 - 4/12 = 2/3, phi_j = (j+0.5) pi/4; lat = asin(2/3) = 41.81 deg
    for j in range(8):
        add(2, (j + 0.5) * 45.0, 43.0, 4 + j)
    np.savez(
        os.path.join(HERE, "ang2pix_cases.npz"),
        nside=np.array(nside), lon=np.array(lon), lat=np.array(lat), ring=np.array(ring),
    )
    print("ang2pix_cases.npz written", len(ring))


if __name__ == "__main__":
    make_ang2pix_cases()
    make_alm2cl()
    make_sht_direct()
