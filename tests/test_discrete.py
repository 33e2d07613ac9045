"""CudaDiscreteMapper against the reference's DiscreteMapper contract (heracles/ducc.py:40-162; reference tests
tests/test_ducc.py): exact adjoint synthesis at free points"""
import numpy as np
import numpy.testing as npt
import pytest


def test_resample_and_properties():
    """resample follows ducc.py:145-162; no device is touched"""
    import heracles_b200 as hb

    mapper = hb.CudaDiscreteMapper(5)
    assert mapper.lmax == 5 and mapper.area == 1.0
    lmax_in = 8
    rng = np.random.default_rng(1)
    alm = rng.standard_normal((2, (lmax_in + 1) * (lmax_in + 2) // 2)) + 0j
    out = mapper.resample(alm)
    assert out.shape == (2, 21)
    i = j = 0
    for m in range(6):
        npt.assert_array_equal(out[:, j : j + 6 - m], alm[:, i : i + 6 - m])
        i += lmax_in - m + 1
        j += 5 - m + 1
    up = hb.CudaDiscreteMapper(10).resample(alm)
    assert up.shape == (2, 66) and np.count_nonzero(up) == alm.size
    assert mapper.transform(alm) is alm


@pytest.mark.gpu
def test_map_values_spin0_direct_sum(oracle):
    import heracles_b200 as hb

    lmax, n = 40, 3000
    rng = np.random.default_rng(8)
    lon = rng.uniform(-400, 800, n)
    lat = np.degrees(np.arcsin(rng.uniform(-1, 1, n)))
    lat[:4] = [90.0, -90.0, 89.999, 0.0]
    vals = rng.standard_normal((3, n))
    mapper = hb.CudaDiscreteMapper(lmax)
    alm = mapper.create(3)
    md = alm.dtype.metadata
    assert md["geometry"] == "discrete" and md["kernel"] == "none" and md["lmax"] == lmax and md["spin"] == 0
    mapper.map_values(lon, lat, alm, vals)
    mapper.map_values(lon[:100], lat[:100], alm, 2.0 * vals[:, :100])  # accumulates
    got = np.asarray(alm)
    theta, phi = np.radians(90.0 - lat), np.radians(lon % 360.0)
    w = vals.copy()
    w[:, :100] *= 3.0
    ref = np.zeros_like(got)
    for m in range(lmax + 1):
        lam = np.array([oracle.lambda_lm(lmax, m, 0, np.cos(t), np.sin(t)) for t in theta])  # [point, l]
        f = w * np.exp(-1j * m * phi)
        ref[:, oracle.almidx(lmax, m, m) : oracle.almidx(lmax, m, m) + lmax - m + 1] = (f @ lam)[:, m:]
    assert np.linalg.norm(got - ref) <= 1e-11 * np.linalg.norm(ref)
    # a plain ndarray target (the reference's create() returns one) and 1-d values
    host = np.zeros(alm.shape[-1], dtype=complex)
    mapper.map_values(lon, lat, host, vals[0])
    mapper.map_values(lon[:100], lat[:100], host, 2.0 * vals[0, :100])
    assert np.linalg.norm(host - ref[0]) <= 1e-11 * np.linalg.norm(ref[0])


@pytest.mark.gpu
@pytest.mark.parametrize("spin", [0, 2])
def test_map_values_at_pixel_centres_is_unweighted_map2alm(oracle, spin):
    """values at HEALPix pixel centres: the adjoint synthesis is the (unweighted, niter 0) map analysis / (4 pi / npix),
    spin 0 and spin 2 (E, B from Q, U in the convention of the HEALPix path)"""
    import heracles_b200 as hb

    nside, lmax = 8, 20
    npix = 12 * nside * nside
    rng = np.random.default_rng(5)
    ipix = rng.choice(npix, 200, replace=False)
    lon, lat = oracle.pix2ang(nside, ipix)
    nrow = 2 if spin == 0 else 4
    vals = rng.standard_normal((nrow, ipix.size))
    maps = np.zeros((nrow, npix))
    maps[:, ipix] = vals
    ref = oracle.map2alm(nside, lmax, maps, spin=spin, niter=0) / (4 * np.pi / npix)
    mapper = hb.CudaDiscreteMapper(lmax)
    shape = (nrow,) if spin == 0 else (nrow // 2, 2)
    alm = mapper.create(*shape, spin=spin)
    mapper.map_values(lon, lat, alm, vals.reshape(*shape, -1), spin=spin)
    got = np.asarray(alm).reshape(nrow, -1)
    assert np.linalg.norm(got - ref) <= 1e-11 * np.linalg.norm(ref)


def test_resample_like_the_reference_test():
    """the reference's own check, tests/test_ducc.py:13-45, on CudaDiscreteMapper"""
    import heracles_b200 as hb

    lmax = 200
    alm = np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)], dtype=complex)
    out = hb.CudaDiscreteMapper(lmax).resample(alm)
    npt.assert_array_equal(out, alm)
    lmax_out = lmax // 2
    out = hb.CudaDiscreteMapper(lmax_out).resample(alm)
    assert out.shape == ((lmax_out + 1) * (lmax_out + 2) // 2,)
    i = j = 0
    for m in range(lmax_out + 1):
        i, j = j, j + lmax_out - m + 1
        npt.assert_array_equal(out[i:j], np.arange(m, lmax_out + 1))
    lmax_out = lmax * 2
    out = hb.CudaDiscreteMapper(lmax_out).resample(alm)
    assert out.shape == ((lmax_out + 1) * (lmax_out + 2) // 2,)
    i = j = 0
    for m in range(lmax + 1):
        i, j = j, j + lmax_out - m + 1
        expected = np.pad(np.arange(m, lmax + 1), (0, lmax_out - lmax))
        npt.assert_array_equal(out[i:j], expected)
    npt.assert_array_equal(out[j:], 0.0)
