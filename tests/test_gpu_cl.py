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


def test_angular_power_spectra_one_gram_and_lazy_alms(hb, oracle):
    """every pair from ONE hcu_alm2cl_rows launch; alm mappings that load lazily (a NEW array per access, as the
    reference's AlmFits does -- 'alms might lazy-load from file', heracles/twopoint.py:233) must not confuse the staging"""
    lmax = 40
    na = (lmax + 1) * (lmax + 2) // 2
    rng = np.random.default_rng(31)
    store = {}
    for k, i, spin in [("POS", 0, 0), ("POS", 1, 0), ("SHE", 0, 2), ("SHE", 1, 2)]:
        a = rng.standard_normal((2, na) if spin else na) + 1j * rng.standard_normal((2, na) if spin else na)
        hb.update_metadata(a, spin=spin, nside=16)
        store[k, i] = a

    class Lazy(dict):
        loads = 0

        def __getitem__(self, key):
            Lazy.loads += 1
            a = np.array(dict.__getitem__(self, key))  # a fresh array every time
            hb.update_metadata(a, **dict.__getitem__(self, key).dtype.metadata)
            return a

    cls = hb.angular_power_spectra(Lazy(store), debias=False)
    assert Lazy.loads == 4  # one load per alm, not one per pair
    assert list(cls) == [("POS", "POS", 0, 0), ("POS", "POS", 0, 1), ("POS", "SHE", 0, 0), ("POS", "SHE", 0, 1),
                         ("POS", "POS", 1, 1), ("POS", "SHE", 1, 0), ("POS", "SHE", 1, 1), ("SHE", "SHE", 0, 0),
                         ("SHE", "SHE", 0, 1), ("SHE", "SHE", 1, 1)]
    for (k1, k2, i1, i2), cl in cls.items():
        ref = oracle.alm2cl(store[k1, i1], store[k2, i2])
        assert np.asarray(cl).shape == ref.shape
        assert np.abs(np.asarray(cl) - ref).max() < 1e-13 * np.abs(ref).max() + 1e-30
    # two mappings (product) and a smaller lmax
    cl2 = hb.angular_power_spectra(store, {("POS", 7): store["POS", 1]}, lmax=20, debias=False)
    assert list(cl2) == [("POS", "POS", 0, 7), ("POS", "POS", 1, 7), ("SHE", "POS", 0, 7), ("SHE", "POS", 1, 7)]
    npt.assert_allclose(np.asarray(cl2["SHE", "POS", 0, 7]), oracle.alm2cl(store["SHE", 0], store["POS", 1], lmax=20), rtol=0, atol=1e-13)


def test_debias_uses_the_callers_pixel_window(hb):
    """the bias of deconvolved spectra is divided by the window the MAPPER used, passed by the caller;
    no mapper / device context is created behind the scenes and no healpy table is needed"""
    lmax = 12
    na = (lmax + 1) * (lmax + 2) // 2
    rng = np.random.default_rng(3)
    a = rng.standard_normal(na) + 0j
    hb.update_metadata(a, spin=0, nside=8, kernel="healpix", deconv=True, fsky=1.0, musq=1.0, dens=4.0)
    pw = 1.0 / (1.0 + 0.05 * np.arange(lmax + 1))
    raw = hb.angular_power_spectra({("P", 0): a}, debias=False)["P", "P", 0, 0]
    deb = hb.angular_power_spectra({("P", 0): a}, pixwin=(pw, pw))["P", "P", 0, 0]
    # both sides of the auto spectrum were deconvolved: the bias is divided by the window twice (twopoint.py:149-165)
    npt.assert_allclose(np.asarray(raw) - np.asarray(deb), 0.25 / pw**2, rtol=1e-12)
    assert np.asarray(deb).dtype.metadata["bias"] == 0.25
    with pytest.raises(RuntimeError, match="pixel window"):
        hb.angular_power_spectra({("P", 0): a})
