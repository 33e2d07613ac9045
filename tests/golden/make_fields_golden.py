"""
Generate tests/golden/fields_pipeline.npz: the reference's OWN driver code
(heracles.fields.Positions / Shears / Weights, heracles.mapping.map_catalogs and
transform, heracles.twopoint.angular_power_spectra) run in the BUILD container over a
Mapper whose four methods are evaluated by the CPU oracle.  The GPU test
(tests/test_gpu_fields.py) replays the same mapper calls against CudaHealpixMapper and
must reproduce these maps / alm / Cl.

Run in the build container only (needs /root/reference):

    python tests/golden/make_fields_golden.py

`import heracles` needs fitsio / healpy / coroutines, none of which is installable here, so
the modules are imported through a stub package and a 20-line `coroutines` shim
(run / gather / sleep, the three functions heracles/mapping.py:113 and fields.py:188-194 use).
"""

import importlib
import math
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def import_reference():
    co = types.ModuleType("coroutines")

    class _Sleep:
        def __await__(self):
            yield

    def sleep():
        return _Sleep()

    async def gather(*coros):
        its = [c.__await__() for c in coros]
        res, done = [None] * len(its), [False] * len(its)
        while not all(done):
            for i, it in enumerate(its):
                if not done[i]:
                    try:
                        next(it)
                    except StopIteration as e:
                        res[i], done[i] = e.value, True
            await sleep()
        return res

    def run(coro):
        it = coro.__await__()
        while True:
            try:
                next(it)
            except StopIteration as e:
                return e.value

    co.sleep, co.gather, co.run = sleep, gather, run
    sys.modules["coroutines"] = co
    pkg = types.ModuleType("heracles")
    pkg.__path__ = [os.path.join(REF, "heracles")]
    sys.modules["heracles"] = pkg
    cat = types.ModuleType("heracles.catalog")
    cat.__path__ = [os.path.join(REF, "heracles", "catalog")]
    sys.modules["heracles.catalog"] = cat
    base = importlib.import_module("heracles.catalog.base")
    for n in dir(base):
        if n[0].isupper():
            setattr(cat, n, getattr(base, n))
    arr = importlib.import_module("heracles.catalog.array")
    mods = {m: importlib.import_module("heracles." + m) for m in ("core", "fields", "mapping", "twopoint")}
    return arr.ArrayCatalog, mods


class OracleMapper:
    """heracles.mapper.Mapper protocol evaluated by the CPU oracle (no deconvolution)"""

    def __init__(self, core, nside, lmax, niter):
        import oracle

        self.o, self.core = oracle, core
        self.nside, self.lmax, self.niter = nside, lmax, niter
        self.deconvolve = False

    @property
    def area(self):
        return 4 * math.pi / (12 * self.nside**2)

    def create(self, *dims, spin=0):
        m = np.zeros((*dims, 12 * self.nside**2))
        self.core.update_metadata(m, geometry="healpix", kernel="healpix", nside=self.nside, lmax=self.lmax, deconv=False, spin=spin)
        return m

    def map_values(self, lon, lat, data, values, spin=0):
        self.o.map_values(self.nside, lon, lat, data, np.ascontiguousarray(values))

    def transform(self, data, spin=0):
        md = data.dtype.metadata or {}
        alm = self.o.map2alm(self.nside, self.lmax, np.asarray(data), spin=spin, niter=self.niter)
        self.core.update_metadata(alm, **{**md, "deconv": False})
        return alm

    def resample(self, data):
        raise NotImplementedError


def main():
    ArrayCatalog, M = import_reference()
    fields, mapping, twopoint, core = M["fields"], M["mapping"], M["twopoint"], M["core"]
    nside, lmax, niter = 16, 24, 3
    rng = np.random.default_rng(50)  # the reference's tests/conftest.py:21 seed
    nbins, nrows = 2, 8_000
    cats, raw = {}, {}
    vis = np.ones(12 * nside**2)
    core.update_metadata(vis, nside=nside)
    for b in range(nbins):
        rows = np.empty(nrows, dtype=[("ra", float), ("dec", float), ("g1", float), ("g2", float), ("w", float)])
        rows["ra"] = rng.uniform(-180, 180, nrows)  # negative longitudes as in tests/test_catalog.py:329
        rows["dec"] = np.degrees(np.arcsin(rng.uniform(-1, 1, nrows)))
        rows["g1"] = rng.normal(0, 0.3, nrows)
        rows["g2"] = rng.normal(0, 0.3, nrows)
        rows["w"] = rng.uniform(0.5, 1.5, nrows)
        rows["w"][rng.integers(0, nrows, 40)] = 0.0  # zero weights are deleted by the shear field (fields.py:420)
        c = ArrayCatalog(rows)
        c.page_size = 3_000  # three pages per catalogue
        c.visibility = vis
        cats[b] = c
        raw[b] = rows
    mapper = OracleMapper(core, nside, lmax, niter)
    fs = {
        "POS": fields.Positions(mapper, "ra", "dec", "w"),
        "SHE": fields.Shears(mapper, "ra", "dec", "g1", "g2", "w"),
        "WHT": fields.Weights(mapper, "ra", "dec", "w"),
    }
    maps = mapping.map_catalogs(fs, cats)
    alms = mapping.transform(fs, maps)
    cls = twopoint.angular_power_spectra(alms, debias=False)
    out = {"nside": nside, "lmax": lmax, "niter": niter, "nbins": nbins, "page_size": 3000}
    for b in range(nbins):
        for col in raw[b].dtype.names:
            out[f"cat{b}_{col}"] = raw[b][col]
    for (k, i), m in maps.items():
        out[f"map_{k}_{i}"] = np.asarray(m)
        md = m.dtype.metadata
        for key in ("nbar", "wbar", "musq", "dens", "fsky"):
            if key in md:
                out[f"md_{k}_{i}_{key}"] = md[key]
    for (k, i), a in alms.items():
        out[f"alm_{k}_{i}"] = np.asarray(a)
    keys = []
    for key, c in cls.items():
        name = "cl_" + "_".join(str(x) for x in key)
        out[name] = np.asarray(c)
        keys.append(name)
    out["cl_keys"] = np.array(keys)
    np.savez_compressed(os.path.join(HERE, "fields_pipeline.npz"), **out)
    print("wrote fields_pipeline.npz:", len(maps), "maps,", len(alms), "alms,", len(cls), "spectra")
    for k in sorted(out):
        if k.startswith("md_"):
            print(k, out[k])


if __name__ == "__main__":
    main()
