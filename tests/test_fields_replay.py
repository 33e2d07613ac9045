"""
tests/golden/fields_pipeline.npz was produced by the reference's own Positions / Shears /
Weights, map_catalogs, transform and angular_power_spectra over an oracle-backed Mapper.
CPU: the replay drivers in fields_replay.py reproduce those maps with the same oracle mapper
(so the drivers mirror the reference).  GPU: the same drivers over CudaHealpixMapper, plus
heracles_b200.transform and angular_power_spectra, reproduce maps, alm and every Cl.
"""
import math

import numpy as np
import pytest

import fields_replay as fr
from conftest import golden


class OracleMapper:
    def __init__(self, oracle, nside, lmax):
        self.o, self.nside, self.lmax = oracle, nside, lmax

    area = property(lambda self: 4 * math.pi / (12 * self.nside**2))

    def create(self, *dims, spin=0):
        return np.zeros((*dims, 12 * self.nside**2))

    def map_values(self, lon, lat, data, values, spin=0):
        self.o.map_values(self.nside, lon, lat, data, np.ascontiguousarray(values))


def test_replay_drivers_mirror_the_reference(oracle):
    g = golden("fields_pipeline.npz")
    maps, md = fr.run_all(OracleMapper(oracle, int(g["nside"]), int(g["lmax"])), g)
    for (k, b), m in maps.items():
        ref = g[f"map_{k}_{b}"]
        assert np.array_equal(np.asarray(m), ref), (k, b)  # same arithmetic, same order: bit identical
        for key, val in md[k, b].items():
            assert val == pytest.approx(float(g[f"md_{k}_{b}_{key}"]), rel=1e-15)


@pytest.mark.gpu
def test_cuda_mapper_reproduces_the_reference_pipeline(hb):
    g = golden("fields_pipeline.npz")
    nside, lmax, niter = int(g["nside"]), int(g["lmax"]), int(g["niter"])
    mapper = hb.CudaHealpixMapper(nside, lmax, deconvolve=False, niter=niter)
    maps, md = fr.run_all(mapper, g)
    for (k, b), m in maps.items():
        ref = g[f"map_{k}_{b}"]
        assert isinstance(m, hb.DeviceArray)
        assert np.abs(np.asarray(m) - ref).max() <= 1e-12 * np.abs(ref).max(), (k, b)
        hb.update_metadata(m, spin=2 if k == "SHE" else 0)

    class F:
        def __init__(self, spin):
            self.mapper_or_error, self.spin = mapper, spin

    fields = {"POS": F(0), "SHE": F(2), "WHT": F(0)}
    alms = hb.transform(fields, maps)
    assert list(alms.keys()) == list(maps.keys())  # reference insertion order (mapping.py:151-171)
    for (k, b), a in alms.items():
        ref = g[f"alm_{k}_{b}"]
        a = np.asarray(a)
        assert a.shape == ref.shape
        for x, y in zip(a.reshape(-1, a.shape[-1]), ref.reshape(-1, ref.shape[-1])):
            assert np.linalg.norm(x - y) <= 1e-10 * np.linalg.norm(y), (k, b)
    cls = hb.angular_power_spectra(alms, debias=False)
    names = ["cl_" + "_".join(str(x) for x in key) for key in cls.keys()]
    assert names == [str(n) for n in g["cl_keys"]]  # same pairs, same canonical order (twopoint.py:198-242)
    auto = {}
    for key, c in cls.items():
        if key[0] == key[1] and key[2] == key[3]:
            auto[key[0], key[2]] = np.asarray(c)
    for key, c in cls.items():
        ref = g["cl_" + "_".join(str(x) for x in key)]
        c = np.asarray(c)
        assert c.shape == ref.shape
        scale = np.sqrt(np.abs(auto[key[0], key[2]]).max() * np.abs(auto[key[1], key[3]]).max())
        assert np.abs(c - ref).max() <= 1e-10 * scale, key
