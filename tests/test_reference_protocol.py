"""
The drop-in boundary against the reference's OWN interface definitions (build container only: /root/reference does not
exist on the GPU box, and nothing here needs a GPU): ``isinstance(CudaHealpixMapper(...), heracles.mapper.Mapper)``
(reference tests/test_healpy.py:29), the constructor / attribute surface of ``heracles.healpy.HealpixMapper`` and the
signatures of ``heracles.mapping.transform`` / ``heracles.twopoint.angular_power_spectra``.  The reference's Field
classes driving this mapper are covered by the recorded-call replay (tests/fields_replay.py, golden produced by
tests/golden/make_fields_golden.py from the reference's own fields.py / mapping.py / twopoint.py).
"""
import ast
import importlib
import inspect
import os
import sys
import types

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "heracles")), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    """the reference's pure-Python modules through a stub package (heracles/__init__.py needs fitsio / healpy)"""
    saved = {k: v for k, v in sys.modules.items() if k == "heracles" or k.startswith("heracles.")}
    pkg = types.ModuleType("heracles")
    pkg.__path__ = [os.path.join(REF, "heracles")]
    sys.modules["heracles"] = pkg
    mods = {name: importlib.import_module("heracles." + name) for name in ("core", "mapper", "twopoint")}
    yield mods
    for k in [k for k in sys.modules if k == "heracles" or k.startswith("heracles.")]:
        del sys.modules[k]
    sys.modules.update(saved)


@pytest.fixture
def mapper(monkeypatch):
    import heracles_b200 as hb
    from heracles_b200 import _lib

    class NoDevice:  # the protocol / attribute checks need no CUDA context
        device = 0

    monkeypatch.setattr(_lib, "get_context", lambda device=None: NoDevice())
    return hb.CudaHealpixMapper(64, 100, deconvolve=False)


def test_isinstance_of_the_reference_protocol(ref, mapper):
    Mapper = ref["mapper"].Mapper
    assert isinstance(mapper, Mapper)  # reference tests/test_healpy.py:29
    for name in ("area", "create", "map_values", "transform", "resample"):
        assert hasattr(mapper, name)
    import math

    assert mapper.area == 4 * math.pi / (12 * 64 * 64)  # hp.nside2pixarea, healpy.py:117-122
    assert (mapper.nside, mapper.lmax, mapper.deconvolve) == (64, 100, False)


def test_discrete_mapper_satisfies_the_protocol(ref):
    """heracles/ducc.py:40-162: same constructor keywords (minus nthreads), properties and methods; isinstance of Mapper"""
    import heracles_b200 as hb

    m = hb.CudaDiscreteMapper(30)
    assert isinstance(m, ref["mapper"].Mapper)
    assert (m.lmax, m.area) == (30, 1.0)
    src = open(os.path.join(REF, "heracles", "ducc.py")).read()
    cls = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "DiscreteMapper")
    for fn in (n for n in cls.body if isinstance(n, ast.FunctionDef) and not n.name.startswith("_")):
        assert hasattr(hb.CudaDiscreteMapper, fn.name), fn.name
        if fn.name in ("create", "map_values", "transform", "resample"):
            ours = [p for p in inspect.signature(getattr(hb.CudaDiscreteMapper, fn.name)).parameters]
            assert ours == [a.arg for a in fn.args.args] + ([fn.args.vararg.arg] if fn.args.vararg else []) + [a.arg for a in fn.args.kwonlyargs] or \
                sorted(ours) == sorted([a.arg for a in fn.args.args] + ([fn.args.vararg.arg] if fn.args.vararg else []) + [a.arg for a in fn.args.kwonlyargs]), fn.name


def test_constructor_and_method_signatures_match_healpixmapper(ref):
    """heracles/healpy.py cannot be imported (healpy), so its class is read from the source"""
    import heracles_b200 as hb

    tree = ast.parse(open(os.path.join(REF, "heracles", "healpy.py")).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "HealpixMapper")
    ref_methods = {n.name: n for n in cls.body if isinstance(n, ast.FunctionDef)}
    for name in ("__init__", "create", "map_values", "transform", "resample"):
        fn = ref_methods[name]
        ref_pos = [a.arg for a in fn.args.args]
        ref_kw = [a.arg for a in fn.args.kwonlyargs]
        sig = inspect.signature(getattr(hb.CudaHealpixMapper, name))
        ours = list(sig.parameters)
        assert ours[: len(ref_pos)] == ref_pos, (name, ours, ref_pos)
        for k in ref_kw:
            assert k in sig.parameters and sig.parameters[k].kind is inspect.Parameter.KEYWORD_ONLY, (name, k)
    # defaults of the reference constructor: lmax = 3 nside // 2, deconvolve = True (healpy.py:75-96)
    assert hb.CudaHealpixMapper.__init__.__kwdefaults__["deconvolve"] is None
    assert "DATAPATH" in vars(hb.CudaHealpixMapper)  # cli.py:536-538 sets it on the class


def test_driver_signatures(ref):
    import heracles_b200 as hb

    tp = ref["twopoint"]
    for ours, theirs in ((hb.angular_power_spectra, tp.angular_power_spectra), (hb.alm2cl, tp.alm2cl), (hb.alm2lmax, tp.alm2lmax)):
        p_ref = inspect.signature(theirs).parameters
        p_our = inspect.signature(ours).parameters
        assert list(p_our)[: len(p_ref)] == list(p_ref) or set(p_ref) <= set(p_our), (ours.__name__, list(p_our), list(p_ref))
        for name, p in p_ref.items():
            assert p_our[name].kind == p.kind, (ours.__name__, name)
    # heracles.mapping.transform(fields, data, *, out=None, progress=None) -- mapping.py:130-136 (needs coroutines: read the source)
    tree = ast.parse(open(os.path.join(REF, "heracles", "mapping.py")).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "transform")
    sig = inspect.signature(hb.transform)
    assert [a.arg for a in fn.args.args] == [n for n, p in sig.parameters.items() if p.kind is inspect.Parameter.POSITIONAL_OR_KEYWORD]
    assert {a.arg for a in fn.args.kwonlyargs} == {n for n, p in sig.parameters.items() if p.kind is inspect.Parameter.KEYWORD_ONLY}


def test_update_metadata_matches_the_reference(ref):
    import numpy as np

    import heracles_b200 as hb

    class Catalog:  # fields.py:312 passes the catalogue, whose `metadata` mapping is merged
        metadata = {"catalog": "cat.fits", "x": 1}

    a, b = np.zeros(4), np.zeros(4)
    for f, arr in ((ref["core"].update_metadata, a), (hb.update_metadata, b)):
        f(arr, z=0)
        f(arr, Catalog(), y=2)
    assert a.dtype.metadata == b.dtype.metadata == {"z": 0, "catalog": "cat.fits", "x": 1, "y": 2}
    c = np.zeros(3, dtype=complex)
    hb.update_metadata(c, a)  # arrays contribute their dtype metadata (used by the transform)
    assert c.dtype.metadata == a.dtype.metadata
