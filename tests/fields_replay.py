"""
Replay of the mapper calls the reference's Field classes make, page by page, against any
object with the Mapper protocol (heracles/mapper.py:33-74).  The sequence, the running means
and the normalisations are those of heracles/fields.py (lines cited); the expected maps in
tests/golden/fields_pipeline.npz come from the reference's own classes
(tests/golden/make_fields_golden.py), so these drivers are checked against the reference
before they are used to check CudaHealpixMapper.
"""
import numpy as np


def pages(g, b):
    """the pages ArrayCatalog yields (heracles/catalog/array.py:53-65): consecutive row blocks"""
    n, size = len(g[f"cat{b}_ra"]), int(g["page_size"])
    for s in range(0, n, size):
        yield {c: g[f"cat{b}_{c}"][s:s + size] for c in ("ra", "dec", "g1", "g2", "w")}


def positions(mapper, g, b, vis):
    """Positions.__call__ with overdensity=True, fields.py:235-316"""
    pos = mapper.create(spin=0)
    ngal, wmean, w2mean = 0, 0.0, 0.0
    for p in pages(g, b):
        lon, lat, w = p["ra"], p["dec"], p["w"]
        mapper.map_values(lon, lat, pos, w, spin=0)          # fields.py:267
        ngal += len(w)
        wmean += (w - wmean).sum() / ngal                    # fields.py:270
        w2mean += (w**2 - w2mean).sum() / ngal
    npix = 4 * np.pi / mapper.area
    nbar = ngal * wmean / 1.0 / npix                         # fields.py:283
    pos /= nbar                                              # fields.py:296
    pos -= vis                                               # fields.py:304
    return pos, dict(nbar=nbar)


def shears(mapper, g, b):
    """Shears (ComplexField/Spin2Field).__call__, fields.py:392-457"""
    val = mapper.create(2, spin=2)
    ngal, wmean, w2mean, var = 0, 0.0, 0.0, 0.0
    for p in pages(g, b):
        keep = p["w"] != 0                                   # page.delete(page[wcol] == 0), fields.py:420
        lon, lat, re, im, w = (p[c][keep] for c in ("ra", "dec", "g1", "g2", "w"))
        re, im = w * re, w * im
        mapper.map_values(lon, lat, val, np.r_[[re, im]], spin=2)   # fields.py:428
        ngal += len(w)
        wmean += (w - wmean).sum() / ngal
        w2mean += (w**2 - w2mean).sum() / ngal
        var += (re**2 + im**2 - var).sum() / ngal
    wbar = ngal / (4 * np.pi) * wmean * mapper.area          # fields.py:440
    val /= wbar                                              # fields.py:446
    return val, dict(wbar=wbar, musq=var / w2mean)


def weights(mapper, g, b):
    """Weights.__call__, fields.py:496-559"""
    wht = mapper.create(spin=0)
    ngal, wmean = 0, 0.0
    for p in pages(g, b):
        keep = p["w"] != 0
        lon, lat, w = p["ra"][keep], p["dec"][keep], p["w"][keep]
        mapper.map_values(lon, lat, wht, w, spin=0)          # fields.py:531
        ngal += len(w)
        wmean += (w - wmean).sum() / ngal
    wbar = ngal / (4 * np.pi) * wmean * mapper.area
    wht /= wbar                                              # fields.py:548
    return wht, dict(wbar=wbar)


def run_all(mapper, g):
    """maps of map_catalogs(fields, catalogs) (mapping.py:61-127) for the fixture's two catalogues"""
    vis = mapper.create(spin=0)
    vis += 1.0
    maps, md = {}, {}
    for b in range(int(g["nbins"])):
        maps["POS", b], md["POS", b] = positions(mapper, g, b, vis)
        maps["SHE", b], md["SHE", b] = shears(mapper, g, b)
        maps["WHT", b], md["WHT", b] = weights(mapper, g, b)
    return maps, md
