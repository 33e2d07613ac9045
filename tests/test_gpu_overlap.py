"""heracles_b200.OverlappedTransform: transforms on a second library context beside the mapping -- same results, same
order as heracles_b200.transform (the reference's heracles.mapping.transform, heracles/mapping.py:130-174)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class _Field:  # the two attributes heracles.mapping.transform reads from a Field
    def __init__(self, mapper, spin):
        self.mapper_or_error, self.spin = mapper, spin


def test_overlapped_transform_matches_transform(hb):
    nside, lmax, nbins = 64, 128, 5
    mapper = hb.CudaHealpixMapper(nside, lmax, deconvolve=False, niter=2, pixel_weights=None, sync=False)
    rng = np.random.default_rng(21)
    n = 20000
    maps = {}
    ov = hb.OverlappedTransform(mapper, batch={0: 3, 2: 2})  # small batches: several groups per spin + a remainder
    for b in range(nbins):
        lon, lat = rng.uniform(0, 360, n), np.degrees(np.arcsin(rng.uniform(-1, 1, n)))
        w, g1, g2 = rng.uniform(0.5, 1.5, n), rng.normal(0, 0.3, n), rng.normal(0, 0.3, n)
        pos, she = mapper.create(spin=0), mapper.create(2, spin=2)
        mapper.map_page(lon, lat, w, g1, g2, pos=pos, she=she)   # on the (now high-priority) mapping stream
        pos /= 0.37
        she /= 1.7
        maps["POS", b], maps["SHE", b] = pos, she
        ov.submit(("POS", b), pos, spin=0)
        ov.submit(("SHE", b), she, spin=2)
    alms = ov.finish()
    fields = {"POS": _Field(mapper, 0), "SHE": _Field(mapper, 2)}
    ref = hb.transform(fields, maps)
    assert list(alms) == list(ref)
    for key in ref:
        a, r = np.asarray(alms[key]), np.asarray(ref[key])
        assert a.shape == r.shape
        assert np.linalg.norm(a - r) <= 1e-12 * np.linalg.norm(r), key
        assert (alms[key].dtype.metadata or {}).get("spin") == (0 if key[0] == "POS" else 2)
    # both contexts are back on their own streams and still work
    m = mapper.create(spin=0)
    mapper.map_values(np.array([10.0]), np.array([20.0]), m, np.array([2.0]))
    assert float(np.asarray(m).sum()) == 2.0
