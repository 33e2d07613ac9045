"""
alm -> Cl parity on the GPU: against the reference's own alm2cl outputs
(tests/golden/alm2cl_reference.npz, produced by heracles/twopoint.py:63-101)
and the oracle.  Tolerance: 1e-10 of sqrt(C_l^aa C_l^bb) (north_star; plain
relative error is ill-defined for cross spectra near zero).
"""
import numpy as np
import numpy.testing as npt
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu


def norm_err(cl, ref, a, b, alm2cl):
    return np.abs(cl - ref).max() / np.sqrt(np.abs(alm2cl(a, a)).max() * np.abs(alm2cl(b, b)).max())


@pytest.mark.parametrize(
    "key,a,b,kw",
    [
        ("cl_pos_pos", "pos", "pos", {}),
        ("cl_pos_pos2", "pos", "pos2", {}),
        ("cl_pos_she", "pos", "she", {}),
        ("cl_she_she", "she", "she", {}),
        ("cl_she_she2", "she", "she2", {}),
        ("cl_pos_pos_lmax20", "pos", "pos2", {"lmax": 20}),
        ("cl_u", "ua", "ub", {}),
        ("cl_u_lmax20", "ua", "ub", {"lmax": 20}),
    ],
)
def test_alm2cl_reference_golden(hb, oracle, key, a, b, kw):
    g = golden("alm2cl_reference.npz")
    cl = hb.alm2cl(g[a], g[b], **kw)
    assert cl.shape == g[key].shape
    assert norm_err(cl, g[key], g[a], g[b], oracle.alm2cl) < 1e-13


def test_alm2cl_auto_default_arg(hb):
    g = golden("alm2cl_reference.npz")
    npt.assert_allclose(hb.alm2cl(g["she"]), g["cl_she_she"], rtol=0, atol=1e-13 * np.abs(g["cl_she_she"]).max())


def test_alm2cl_large_block(hb, oracle):
    lmax = 300
    rng = np.random.default_rng(11)
    na = (lmax + 1) * (lmax + 2) // 2
    a = rng.standard_normal((7, na)) + 1j * rng.standard_normal((7, na))
    b = rng.standard_normal((3, 2, na)) + 1j * rng.standard_normal((3, 2, na))
    cl = hb.alm2cl(a, b)
    ref = oracle.alm2cl(a, b)
    assert cl.shape == (7, 3, 2, lmax + 1)
    assert np.abs(cl - ref).max() < 1e-13 * np.abs(ref).max()
    auto = hb.alm2cl(a)
    refa = oracle.alm2cl(a, a)
    assert np.abs(auto - refa).max() < 1e-13 * np.abs(refa).max()
    npt.assert_allclose(auto, np.swapaxes(auto, 0, 1), rtol=0, atol=1e-14 * np.abs(auto).max())


def test_angular_power_spectra(hb, oracle):
    # tests/test_twopoint.py:24-138: key set, shapes, metadata and bias
    lmax = 32
    rng = np.random.default_rng(50)
    size = (lmax + 1) * (lmax + 2) // 2
    alms = {}
    for n, s in {"POS": 0, "SHE": 2}.items():
        shape = (size, 2) if s == 0 else (2, size, 2)
        for i in (0, 1):
            a = rng.standard_normal(shape) @ [1, 1j]
            a.dtype = np.dtype(a.dtype, metadata={"nside": 32, "spin": s, "fsky": 0.5, "musq": 2.0, "dens": 10.0, "kernel": "other"})
            alms[n, i] = a
    comb = {
        ("POS", "POS", 0, 0): (lmax + 1,),
        ("POS", "POS", 0, 1): (lmax + 1,),
        ("POS", "POS", 1, 1): (lmax + 1,),
        ("POS", "SHE", 0, 0): (2, lmax + 1),
        ("POS", "SHE", 0, 1): (2, lmax + 1),
        ("POS", "SHE", 1, 0): (2, lmax + 1),
        ("POS", "SHE", 1, 1): (2, lmax + 1),
        ("SHE", "SHE", 0, 0): (2, 2, lmax + 1),
        ("SHE", "SHE", 0, 1): (2, 2, lmax + 1),
        ("SHE", "SHE", 1, 1): (2, 2, lmax + 1),
    }
    cls = hb.angular_power_spectra(alms, debias=False)
    assert set(cls.keys()) == set(comb.keys())
    for key, shape in comb.items():
        assert np.shape(cls[key]) == shape
        k1, k2, i1, i2 = key
        ref = oracle.alm2cl(alms[k1, i1], alms[k2, i2])
        assert np.abs(np.asarray(cls[key]) - ref).max() < 1e-12 * 4
        md = cls[key].dtype.metadata
        assert md["spin_1"] == (0 if k1 == "POS" else 2)
        assert ("bias" in md) == (k1 == k2 and i1 == i2)
    # bias = factor fsky musq / dens, removed from l >= max spin (twopoint.py:260-273,104-170)
    deb = hb.angular_power_spectra(alms, debias=True, include=[("SHE", "SHE", 0, 0), ("POS", "POS", 1, 1)])
    assert set(deb.keys()) == {("SHE", "SHE", 0, 0), ("POS", "POS", 1, 1)}
    b = 0.5 * 0.5 * 2.0 / 10.0
    raw, d = np.asarray(cls["SHE", "SHE", 0, 0]), np.asarray(deb["SHE", "SHE", 0, 0])
    npt.assert_allclose(raw[0, 0, 2:] - d[0, 0, 2:], b)
    npt.assert_allclose(raw[1, 1, 2:] - d[1, 1, 2:], b)
    npt.assert_allclose(raw[0, 1], d[0, 1], rtol=0, atol=1e-14)
    npt.assert_allclose(raw[0, 0, :2], d[0, 0, :2], rtol=0, atol=1e-14)
    raw, d = np.asarray(cls["POS", "POS", 1, 1]), np.asarray(deb["POS", "POS", 1, 1])
    npt.assert_allclose(raw - d, 2 * b)


def test_alm2cl_mslice_parts_sum_to_the_whole(hb, ctx):
    # the multi-GPU path: every rank sums its own m = rank (mod world); the parts add up to alm2cl
    import ctypes

    from heracles_b200 import DeviceArray, _lib

    lmax, n, world = 37, 3, 4
    nalm = (lmax + 1) * (lmax + 2) // 2
    rng = np.random.default_rng(12)
    alm = rng.standard_normal((n, nalm)) + 1j * rng.standard_normal((n, nalm))
    dev = DeviceArray.zeros(ctx, alm.shape, dtype=np.complex128)
    dev[:] = alm
    dev.to_device()
    full = np.asarray(hb.alm2cl(alm, alm))
    parts = np.zeros((n, n, lmax + 1))
    for r in range(world):
        cl = DeviceArray.zeros(ctx, (n, n, lmax + 1))
        _lib.check(ctx.lib.hcu_alm2cl_mslice(ctx.handle, n, ctypes.c_void_p(dev.device_ptr), nalm, lmax, n,
                                             ctypes.c_void_p(dev.device_ptr), nalm, lmax, lmax, world, r,
                                             ctypes.c_void_p(cl.device_ptr)))
        ctx.synchronize()
        parts += np.asarray(cl)
    scale = np.abs(full).max()
    assert np.abs(parts - full).max() < 1e-13 * scale
