"""heracles_b200.io: heracles.io.read_vmap (heracles/io.py:360-381) without healpy -- FITS parsing on the CPU, the
resolution change and the transform on the GPU"""
import os

import numpy as np
import numpy.testing as npt
import pytest


def write_healpix_fits(path, cols, ordering="RING", fmt="D", rows_of=1024):
    """a HEALPix map file as healpy.write_map lays it out: one binary table, `rows_of` values per row and column"""
    npix = cols[0].size
    rep = rows_of if npix % rows_of == 0 else 1
    nrow = npix // rep
    code = {"D": ">f8", "E": ">f4"}[fmt]
    width = np.dtype(code).itemsize * rep

    def card(key, val, quote=False):
        v = f"'{val}'" if quote else str(val)
        return f"{key:<8}= {v:>20}".ljust(80)

    def block(cards):
        s = "".join(cards) + "END".ljust(80)
        return (s + " " * (-len(s) % 2880)).encode("ascii")

    primary = block([card("SIMPLE", "T"), card("BITPIX", 8), card("NAXIS", 0), card("EXTEND", "T")])
    cards = [card("XTENSION", "BINTABLE", True), card("BITPIX", 8), card("NAXIS", 2), card("NAXIS1", width * len(cols)),
             card("NAXIS2", nrow), card("PCOUNT", 0), card("GCOUNT", 1), card("TFIELDS", len(cols))]
    for i in range(len(cols)):
        cards += [card(f"TTYPE{i + 1}", f"MAP{i}", True), card(f"TFORM{i + 1}", f"{rep}{fmt}", True)]
    cards += [card("PIXTYPE", "HEALPIX", True), card("ORDERING", ordering, True), card("NSIDE", int(round((npix / 12) ** 0.5))),
              card("INDXSCHM", "IMPLICIT", True)]
    dt = np.dtype([(f"c{i}", code, (rep,)) for i in range(len(cols))])
    tab = np.zeros(nrow, dtype=dt)
    for i, c in enumerate(cols):
        tab[f"c{i}"] = c.reshape(nrow, rep)
    data = tab.tobytes()
    with open(path, "wb") as f:
        f.write(primary + block(cards) + data + b"\0" * (-len(data) % 2880))


def test_read_map_and_unseen(tmp_path):
    from heracles_b200 import io

    nside = 16
    rng = np.random.default_rng(2)
    a, b = rng.uniform(0, 1, 12 * nside**2), rng.uniform(0, 1, 12 * nside**2)
    a[5:40] = io.UNSEEN
    path = os.path.join(tmp_path, "vmap.fits")
    write_healpix_fits(path, [a, b])
    npt.assert_array_equal(io.read_map(path, field=1), b)
    got = io.read_vmap(path)  # no resolution change, no transform: nothing touches the GPU
    exp = a.copy()
    exp[5:40] = 0.0
    npt.assert_array_equal(got, exp)
    write_healpix_fits(path, [a.astype(np.float32).astype(np.float64)], fmt="E", rows_of=1)
    npt.assert_array_equal(io.read_map(path), a.astype(np.float32).astype(np.float64))
    with pytest.raises(IndexError):
        io.read_map(path, field=3)


@pytest.mark.gpu
def test_read_vmap_transform(tmp_path, oracle):
    """NEST file at nside 32 -> RING, ud_grade to nside 16, map2alm (niter 3) / pixel window: against the oracle"""
    from heracles_b200 import io

    nside_in, nside, lmax = 32, 16, 40
    rng = np.random.default_rng(4)
    ring = rng.uniform(0, 1, 12 * nside_in**2)
    ring[100:200] = io.UNSEEN
    nest = ring[oracle.nest2ring(nside_in, np.arange(ring.size))]
    path = os.path.join(tmp_path, "vmap_nest.fits")
    write_healpix_fits(path, [nest], ordering="NESTED")
    npt.assert_array_equal(io.read_map(path), ring)
    clean = np.where(ring == io.UNSEEN, 0.0, ring)
    low = clean[oracle.nest2ring(nside_in, np.arange(ring.size))].reshape(-1, 4).mean(axis=1)[oracle.ring2nest(nside, np.arange(12 * nside**2))]
    with pytest.warns(UserWarning, match="changing NSIDE"):
        got = np.asarray(io.read_vmap(path, nside=nside))
    npt.assert_allclose(got, low, rtol=1e-14, atol=1e-15)
    pw = 1.0 / (1.0 + 1e-3 * np.arange(lmax + 1) ** 2)
    with pytest.warns(UserWarning):
        alm = np.asarray(io.read_vmap(path, nside=nside, transform=True, lmax=lmax, pixwin=(pw, pw), pixel_weights=None))
    ref = oracle.almxfl(oracle.map2alm(nside, lmax, low[None], spin=0, niter=3), 1.0 / pw)[0]
    assert np.linalg.norm(alm - ref) <= 1e-10 * np.linalg.norm(ref)
