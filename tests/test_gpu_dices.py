"""
DICES jackknife on the device: batched region transforms and delete-d spectra by alm subtraction.
Mirrors the reference's tests/test_dices.py:30-55 (alm subtraction and map masking must give identical Cls) and checks
the batched region alm against one-map-at-a-time transforms and against the oracle.
"""
from itertools import combinations

import numpy as np
import numpy.testing as npt
import pytest

pytestmark = pytest.mark.gpu


class F:  # the two attributes heracles.mapping.transform reads from a Field
    def __init__(self, mapper, spin):
        self.mapper_or_error, self.spin = mapper, spin


@pytest.fixture
def setup(hb):
    nside, lmax, njk = 32, 48, 5
    rng = np.random.default_rng(50)  # tests/conftest.py:20-22
    npix = 12 * nside**2
    mapper = hb.CudaHealpixMapper(nside, lmax, deconvolve=False, niter=1, pixel_weights=None)
    fields = {"POS": F(mapper, 0), "SHE": F(mapper, 2)}
    jk_map = rng.integers(0, njk + 1, npix).astype(float)  # 0 = outside the footprint
    maps = {}
    for i in (1, 2):
        p = mapper.create(spin=0)
        p[:] = rng.standard_normal(npix) * (jk_map != 0)
        hb.update_metadata(p, fsky=0.8, musq=1.0, dens=2.0 + i)
        s = mapper.create(2, spin=2)
        s[:] = rng.standard_normal((2, npix)) * (jk_map != 0)
        hb.update_metadata(s, fsky=0.8, musq=0.5, dens=3.0)
        maps["POS", i], maps["SHE", i] = p, s
    return hb, mapper, fields, maps, jk_map, njk


def test_region_alms_match_single_transforms_and_oracle(setup, oracle):
    hb, mapper, fields, maps, jk_map, njk = setup
    alms = hb.dices.region_alms(fields, maps, jk_map, [0] + hb.dices.region_ids(jk_map))
    assert list(alms) == list(range(njk + 1))
    for k in range(njk + 1):
        assert list(alms[k]) == list(maps)
        for (name, i), m in maps.items():
            spin = fields[name].spin
            masked = np.asarray(m) * ((jk_map == k) if k else 1.0)
            one = np.asarray(mapper.transform(masked, spin=spin))  # the reference's one map2alm per region map
            got = alms[k][name, i]
            assert got.dtype.metadata["spin"] == spin and got.dtype.metadata["deconv"] is False
            assert got.dtype.metadata["fsky"] == 0.8  # the region alm carry the full-footprint metadata
            npt.assert_allclose(np.asarray(got), one, rtol=0, atol=1e-13 * np.abs(one).max())
    m = np.asarray(maps["SHE", 2]) * (jk_map == 3)
    ref = oracle.map2alm(mapper.nside, mapper.lmax, m, spin=2, niter=1)
    err = np.linalg.norm(np.asarray(alms[3]["SHE", 2]) - ref) / np.linalg.norm(ref)
    assert err < 1e-10


def test_region_alm_cls(setup):
    """tests/test_dices.py:30-55: ALM subtraction and map masking must give identical Cls"""
    hb, mapper, fields, maps, jk_map, njk = setup
    for nd in (0, 1, 2):
        cls = hb.dices.jackknife_cls(maps, jk_map, fields, nd=nd, debias=False)
        tuples = list(combinations(range(1, njk + 1), nd))
        assert list(cls) == tuples
        for regions in tuples[:4]:
            keep = ~np.isin(jk_map, regions)
            removed = {}
            for key, m in maps.items():
                r = mapper.create(*m.shape[:-1], spin=fields[key[0]].spin)
                r[:] = np.asarray(m) * keep
                hb.update_metadata(r, **m.dtype.metadata)
                removed[key] = r
            ref = hb.angular_power_spectra(hb.transform(fields, removed), debias=False)
            assert list(cls[regions]) == list(ref)
            for key in ref:
                npt.assert_allclose(np.asarray(cls[regions][key]), np.asarray(ref[key]), rtol=1e-7, atol=1e-10,
                                    err_msg=f"nd={nd}, regions={regions}, key={key}")


def test_jackknife_bias_hook(setup):
    hb, mapper, fields, maps, jk_map, njk = setup
    seen = []

    def correct(cls, regions):
        seen.append(regions)
        return cls

    cls = hb.dices.jackknife_cls(maps, jk_map, fields, nd=1, correct=correct, debias=False)
    assert seen == [(k,) for k in range(1, njk + 1)] == list(cls)
    assert cls[(1,)]["POS", "POS", 1, 1].dtype.metadata["bias"] == 0.8 * 1.0 / 3.0
