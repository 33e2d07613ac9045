"""
CPU tests of the drop-in boundary: the C-ABI library loads, exports every
symbol include/heracles_cuda.h declares, and fails loudly without a device.
No compute calls are made here.
"""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "heracles_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hcu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from heracles_b200 import _lib

    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    # and the Python binding declares a prototype for each
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)
    assert lib.hcu_version() == 100


def test_no_cpu_fallback():
    from heracles_b200 import _lib

    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.HeraclesCudaError, match="no CPU path"):
        _lib.Context(0)
    import heracles_b200 as hb

    with pytest.raises(_lib.HeraclesCudaError):
        hb.CudaHealpixMapper(8)


def test_product_does_not_import_oracle():
    # the oracle is test infrastructure; the product must never route through it
    pkg = os.path.join(ROOT, "heracles_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "liboracle" not in src, f


def test_sass_is_sm100a():
    import subprocess

    so = os.path.join(ROOT, "heracles_b200", "lib", "libheracles_cuda.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_pixwin_fits_reader(tmp_path):
    # minimal FITS binary table like HEALPix' pixel_window_n0004.fits
    from heracles_b200.mapper import read_pixwin_fits

    def card(k, v):
        return f"{k:<8}= {v:>20}".ljust(80).encode()

    def block(cards):
        raw = b"".join(cards) + "END".ljust(80).encode()
        return raw + b" " * (-len(raw) % 2880)

    pw = np.linspace(1, 0.5, 17)
    primary = block([card("SIMPLE", "T"), card("BITPIX", 8), card("NAXIS", 0), card("EXTEND", "T")])
    tab = np.zeros(17, dtype=[("T", ">f8"), ("P", ">f8")])
    tab["T"], tab["P"] = pw, pw**2
    ext = block(
        [
            card("XTENSION", "'BINTABLE'"), card("BITPIX", 8), card("NAXIS", 2), card("NAXIS1", 16),
            card("NAXIS2", 17), card("PCOUNT", 0), card("GCOUNT", 1), card("TFIELDS", 2),
            card("TFORM1", "'1D'"), card("TFORM2", "'1D'"),
        ]
    )
    data = tab.tobytes()
    p = tmp_path / "pixel_window_n0004.fits"
    p.write_bytes(primary + ext + data + b"\0" * (-len(data) % 2880))
    t, pol = read_pixwin_fits(str(p))
    np.testing.assert_array_equal(t, pw)
    np.testing.assert_array_equal(pol, pw**2)
