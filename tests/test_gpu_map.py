"""
catalogue -> map parity on the GPU (through the C ABI via CudaHealpixMapper).
Mirrors the reference's tests/test_healpy.py:17-78 with the oracle standing in
for healpy.  Bit-exact pixel indices; maps equal to the sequential scatter up to
the order of the floating-point adds (<= 1e-12 relative, exact when no pixel is
hit twice).
"""

import numpy as np
import numpy.testing as npt
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu


def gpu_ang2pix(ctx, nside, lon, lat, nest=False):
    from heracles_b200 import _lib

    lon = np.ascontiguousarray(lon, dtype=np.float64)
    lat = np.ascontiguousarray(lat, dtype=np.float64)
    out = np.empty(lon.size, dtype=np.int64)
    _lib.check(
        ctx.lib.hcu_ang2pix(
            ctx.handle, nside, int(nest), lon.ctypes.data, lat.ctypes.data, lon.size, out.ctypes.data
        )
    )
    return out


def test_mapper_protocol(hb, rng):
    nside = 1 << rng.integers(1, 10)
    npix = 12 * nside * nside
    lmax = 1 << rng.integers(1, 10)
    deconv = False
    mapper = hb.CudaHealpixMapper(nside, lmax, deconvolve=deconv)
    for attr in ("area", "create", "map_values", "transform", "resample"):
        assert hasattr(mapper, attr)
    assert mapper.nside == nside
    assert mapper.lmax == lmax
    assert mapper.deconvolve == deconv
    assert mapper.area == 4 * np.pi / npix
    m = mapper.create(1, 2, 3, spin=-3)
    assert isinstance(m, np.ndarray)
    assert m.shape == (1, 2, 3, npix)
    assert m.dtype.metadata == {
        "geometry": "healpix",
        "kernel": "healpix",
        "nside": nside,
        "lmax": lmax,
        "deconv": deconv,
        "spin": -3,
    }
    npt.assert_array_equal(m, 0.0)


@pytest.mark.parametrize("nside", [1, 2, 4, 64, 512, 4096, 8192])
@pytest.mark.parametrize("nest", [False, True])
def test_ang2pix_bit_exact(ctx, oracle, nside, nest):
    rng = np.random.default_rng(nside)
    n = 200_000
    lon = rng.uniform(-360, 720, n)
    lat = np.degrees(np.arcsin(rng.uniform(-1, 1, n)))
    # adversarial rows: poles, cap/belt boundary, wrap-arounds, pixel centres and edges
    b = np.degrees(np.arcsin(2 / 3))
    adv_lon = np.array([0, 0, 0, 360, -360, 720, 90, 180, 270, 45, 1e-300, -1e-300, 359.99999999999994, 0, 0])
    adv_lat = np.array([90, -90, 0, 0, 0, 0, b, -b, np.nextafter(b, 100), np.nextafter(-b, -100), 89.9999, -89.9999, 0, 89.5, -89.5])
    ip = rng.integers(0, 12 * nside * nside, 20000)
    clon, clat = oracle.pix2ang(nside, ip, nest=nest)
    lon = np.concatenate([lon, adv_lon, clon])
    lat = np.concatenate([lat, adv_lat, clat])
    got = gpu_ang2pix(ctx, nside, lon, lat, nest)
    exp = oracle.ang2pix(nside, lon, lat, nest=nest)
    mism = int(np.count_nonzero(got != exp))
    assert mism == 0, f"{mism} of {lon.size} pixel indices differ"


def test_ang2pix_golden_cases(ctx):
    g = golden("ang2pix_cases.npz")
    for ns in np.unique(g["nside"]):
        sel = g["nside"] == ns
        got = gpu_ang2pix(ctx, int(ns), g["lon"][sel], g["lat"][sel])
        npt.assert_array_equal(got, g["ring"][sel])


def test_ang2pix_invalid_rows(ctx):
    lon = np.array([0.0, 10.0, np.nan, 5.0])
    lat = np.array([91.0, -90.5, 0.0, np.nan])
    got = gpu_ang2pix(ctx, 16, lon, lat)
    npt.assert_array_equal(got, -1)


def test_healpix_maps(hb, oracle, rng):
    # tests/test_healpy.py:17-78
    nside = 1 << rng.integers(1, 10)
    npix = 12 * nside * nside
    mapper = hb.CudaHealpixMapper(nside, 16, deconvolve=False)
    size = 1000
    lon = rng.uniform(0, 360, size=size)
    lat = np.degrees(np.arcsin(rng.uniform(-1, 1, size=size)))
    x = rng.standard_normal(size=size)
    y = rng.standard_normal(size=size)
    ipix = oracle.ang2pix(nside, lon, lat)

    m = mapper.create()
    mapper.map_values(lon, lat, m, x)
    expected = np.zeros(npix)
    np.add.at(expected, ipix, x)
    npt.assert_allclose(m, expected, rtol=0, atol=1e-12 * np.abs(expected).max())

    m = mapper.create(2)
    mapper.map_values(lon, lat, m, np.stack([x, y]))
    expected = np.zeros((2, npix))
    np.add.at(expected[0], ipix, x)
    np.add.at(expected[1], ipix, y)
    npt.assert_allclose(m, expected, rtol=0, atol=1e-12 * np.abs(expected).max())


@pytest.mark.parametrize("aggregate", [False, True])
def test_map_values_many_pages(hb, oracle, aggregate):
    nside = 256
    npix = 12 * nside * nside
    mapper = hb.CudaHealpixMapper(nside, 16, deconvolve=False, aggregate=aggregate)
    pos = mapper.create()
    she = mapper.create(2)
    ref_pos = np.zeros(npix)
    ref_she = np.zeros((2, npix))
    for page in range(3):
        rng = np.random.default_rng(50 + page)
        n = 700_001 if page == 0 else 300_000  # > one staging slot, ragged
        lon = rng.uniform(0, 360, n)
        lat = np.degrees(np.arcsin(rng.uniform(-1, 1, n)))
        if page == 2:  # spatially clustered page: many duplicate pixels per warp
            lon = np.sort(lon % 3.0)
            lat = np.sort(lat % 2.0)
        w = rng.uniform(0.5, 1.5, n)
        g = rng.normal(0, 0.3, (2, n))
        mapper.map_values(lon, lat, pos, w)
        mapper.map_values(lon, lat, she, g, spin=2)
        oracle.map_values(nside, lon, lat, ref_pos, w)
        oracle.map_values(nside, lon, lat, ref_she, g)
    npt.assert_allclose(pos, ref_pos, rtol=0, atol=1e-12 * ref_pos.max())
    npt.assert_allclose(she, ref_she, rtol=0, atol=1e-12 * np.abs(ref_she).max())
    # checksum of checksums: total weight is conserved
    assert abs(np.asarray(pos).sum() - ref_pos.sum()) <= 1e-9 * ref_pos.sum()


def test_map_values_edge_cases(hb):
    mapper = hb.CudaHealpixMapper(8, 8, deconvolve=False)
    m = mapper.create()
    # empty page
    mapper.map_values(np.empty(0), np.empty(0), m, np.empty(0))
    npt.assert_array_equal(m, 0.0)
    # read-only, non-contiguous and big-endian inputs (catalog/base.py:67-69, healpy.py:43-55)
    lon = np.array([10.0, 20.0, 30.0, 40.0])[::2]
    lat = np.array([0.0, 1.0, 2.0, 3.0])[::2].astype(">f8")
    w = np.array([1.0, 2.0])
    w.flags.writeable = False
    mapper.map_values(lon, lat, m, w)
    assert np.asarray(m).sum() == 3.0
    # invalid latitude raises like healpy's check
    with pytest.raises(ValueError):
        mapper.map_values(np.array([0.0]), np.array([95.0]), m, np.array([1.0]))
    # plain numpy map as target
    h = np.zeros(12 * 64)
    mapper.map_values(np.array([0.0]), np.array([0.0]), h, np.array([2.5]))
    assert h.sum() == 2.5


def test_inplace_field_ops(hb):
    # the Field layer's in-place normalisation (fields.py:296,304,446) runs on the device
    mapper = hb.CudaHealpixMapper(16, 8, deconvolve=False)
    m = mapper.create()
    m += 4.0
    m /= 4.0
    npt.assert_array_equal(m, 1.0)
    v = mapper.create()
    v[:] = 1.0
    m -= v
    npt.assert_array_equal(m, 0.0)
    m2 = mapper.create(2, spin=2)
    m2 += 3.0
    m2 /= 2.0
    m2 *= 2.0
    npt.assert_array_equal(m2, 3.0)
    assert m2.dtype.metadata["spin"] == 2
    # generic numpy still works on the managed array
    assert float(np.mean(m2)) == 3.0
    assert (m2 * 2).sum() == 6.0 * m2.size


def test_resample(hb, oracle):
    mapper = hb.CudaHealpixMapper(8, 8, deconvolve=False)
    rng = np.random.default_rng(3)
    big = rng.standard_normal(12 * 32 * 32)
    out = np.asarray(mapper.resample(big))
    # mean over the 16 NEST children of each nside=8 pixel
    nest = big[oracle.nest2ring(32, np.arange(big.size))].reshape(-1, 16).mean(axis=1)
    exp = nest[oracle.ring2nest(8, np.arange(12 * 64))]
    npt.assert_allclose(out, exp, rtol=1e-14, atol=1e-15)
    small = rng.standard_normal(12 * 4 * 4)
    up = np.asarray(mapper.resample(small))
    parent = oracle.nest2ring(4, oracle.ring2nest(8, np.arange(12 * 64)) // 4)
    npt.assert_array_equal(up, small[parent])


@pytest.mark.parametrize("pinned", [False, True])
def test_map_page_fused(hb, oracle, pinned):
    """hcu_map_page: POS + SHE maps of a tomographic bin from ONE pass over the page (5 columns instead of 7),
    with the Field layer's running sums reduced on the device (heracles/fields.py:262-271, 420-433)"""
    nside, n = 64, 300_000
    rng = np.random.default_rng(77)
    lon = rng.uniform(-180, 540, n)
    lat = np.degrees(np.arcsin(rng.uniform(-1, 1, n)))
    w = rng.uniform(0.5, 1.5, n)
    w[rng.integers(0, n, 1000)] = 0.0  # page.delete(page[wcol] == 0) for the shears, fields.py:420-421
    g1, g2 = rng.normal(0, 0.3, (2, n))
    mapper = hb.CudaHealpixMapper(nside, deconvolve=False)
    ctx = mapper.context
    cols = [lon, lat, w, g1, g2]
    if pinned:  # pinned pages go straight to cudaMemcpyAsync, pageable ones through the staging slots
        from heracles_b200 import _lib
        import ctypes

        keep = []
        for i, c in enumerate(cols):
            p = ctx.malloc_pinned(c.nbytes)
            buf = np.ctypeslib.as_array((ctypes.c_double * c.size).from_address(p))
            buf[:] = c
            keep.append(p)
            cols[i] = buf
    pos, she, stats = mapper.create(), mapper.create(2, spin=2), mapper.new_page_stats()
    # two "pages"
    h = n // 3
    mapper.map_page(*(c[:h] for c in cols), pos=pos, she=she, stats=stats)
    mapper.map_page(*(c[h:] for c in cols), pos=pos, she=she, stats=stats)
    ref_pos = np.zeros(12 * nside**2)
    ref_she = np.zeros((2, 12 * nside**2))
    oracle.map_values(nside, lon, lat, ref_pos, w)
    nz = w != 0
    oracle.map_values(nside, lon[nz], lat[nz], ref_she, np.stack([w * g1, w * g2])[:, nz])
    npt.assert_allclose(np.asarray(pos), ref_pos, rtol=0, atol=1e-12 * ref_pos.max())
    npt.assert_allclose(np.asarray(she), ref_she, rtol=0, atol=1e-12 * np.abs(ref_she).max())
    (ng0, wm0, w2m0), (ng2, wm2, w2m2, var) = hb.CudaHealpixMapper.page_means(stats)
    assert ng0 == n and ng2 == int(nz.sum())
    npt.assert_allclose([wm0, w2m0], [w.mean(), (w**2).mean()], rtol=1e-13)
    npt.assert_allclose([wm2, w2m2, var], [w[nz].mean(), (w[nz] ** 2).mean(), ((w * g1) ** 2 + (w * g2) ** 2)[nz].mean()], rtol=1e-13)
    # positions only, unit weights (fields.py:265); NaN rows are what CatalogPage.get raises on
    pos2, st2 = mapper.create(), mapper.new_page_stats()
    lon2 = lon.copy()
    lon2[5] = np.nan
    mapper.map_page(lon2, lat, pos=pos2, stats=st2)
    ref2 = np.zeros(12 * nside**2)
    oracle.map_values(nside, np.delete(lon, 5), np.delete(lat, 5), ref2, np.ones(n - 1))
    npt.assert_array_equal(np.asarray(pos2), ref2)
    with pytest.raises(ValueError, match="invalid values"):
        hb.CudaHealpixMapper.page_means(st2)
    if pinned:
        for p in keep:
            ctx.free(p)
