"""
``CudaDiscreteMapper`` -- the pixel-free mapper of Heracles (``heracles.ducc.DiscreteMapper``, ``heracles/ducc.py:40-162``)
on the B200: ``map_values`` adds ``sum_i v_i conj(sY_lm(theta_i, phi_i))`` to the alm directly.

The reference calls ``ducc0.sht.adjoint_synthesis_general`` (a NUFFT, epsilon 1e-12).  Here every catalogue point is a
ring of its own for the Legendre analysis kernels of the HEALPix path (``hcu_points2alm``): the sum is exact and costs
O(points x lmax^2) -- fine for the catalogue sizes and band limits the discrete mapper is used with, not a replacement
of the NUFFT for 1e9 rows.  Same constructor, properties, ``create / map_values / transform / resample`` and metadata
as the reference class.
"""

from __future__ import annotations

from typing import Any

import numpy as np

from . import _lib
from .arrays import DeviceArray, update_metadata
from .mapper import _native, _ptr

__all__ = ["CudaDiscreteMapper"]

c_vp = _lib.c_vp


class CudaDiscreteMapper:
    """Mapper that creates alms directly."""

    def __init__(self, lmax: int, *, dtype: Any = np.complex128, device: int | None = None, context: Any = None) -> None:
        if np.dtype(dtype) != np.complex128:
            raise NotImplementedError("CudaDiscreteMapper computes in complex128 only")
        self.__lmax = int(lmax)
        self.__dtype = np.dtype(dtype)
        self.__device, self.__ctx = device, context  # the library context is created on first use

    @property
    def lmax(self) -> int:
        """The maximum angular mode number."""
        return self.__lmax

    @property
    def area(self) -> float:
        """The effective area for this mapper."""
        return 1.0

    @property
    def context(self):
        if self.__ctx is None:
            self.__ctx = _lib.get_context(self.__device)
        return self.__ctx

    _ctx = context

    def create(self, *dims: int, spin: int = 0):
        """Create zero alms (managed memory: the Field layer's in-place arithmetic keeps working)."""
        lmax = self.__lmax
        m = DeviceArray.zeros(self._ctx, (*dims, (lmax + 1) * (lmax + 2) // 2), dtype=self.__dtype)
        update_metadata(m, geometry="discrete", kernel="none", lmax=lmax, spin=spin)
        return m

    def map_values(self, lon, lat, data, values, spin: int = 0) -> None:
        """Add values to alms (ducc.py:92-133)."""
        if spin not in (0, 2):
            msg = f"spin-{spin} values not yet supported"
            raise NotImplementedError(msg)
        lon, lat = _native(lon), _native(lat)
        values = np.asarray(values)
        flatten = values.ndim == 1
        vals = _native(values.reshape(1, -1) if flatten else values.reshape(-1, values.shape[-1]))
        n = lon.size
        if lat.size != n or vals.shape[-1] != n:
            raise ValueError("columns differ in size")
        nalm = (self.__lmax + 1) * (self.__lmax + 2) // 2
        if data.shape[-1] != nalm or int(np.prod(data.shape[:-1], dtype=np.int64)) != vals.shape[0]:
            raise ValueError("data does not match the values")
        if spin == 2 and vals.shape[0] % 2:
            raise ValueError("spin-2 values must come as (2, n)")
        if isinstance(data, DeviceArray) and data.device_ptr is not None and data.dtype == np.complex128:
            data.to_device()
            target, tmp = data, None
        else:  # a plain ndarray: accumulate in a device array and add it on the host
            tmp = DeviceArray.zeros(self._ctx, (vals.shape[0], nalm), dtype=np.complex128)
            target = tmp
        cap = int(self._ctx.lib.hcu_legendre_batch_size(spin))
        rows = vals.shape[0]
        tptr = target.device_ptr
        for r0 in range(0, rows, cap):
            nb = min(cap, rows - r0)
            _lib.check(
                self._ctx.lib.hcu_points2alm(
                    self._ctx.handle, self.__lmax, spin, nb, n, c_vp(_ptr(lon)), c_vp(_ptr(lat)),
                    c_vp(_ptr(vals) + 8 * r0 * n), n, c_vp(tptr + 16 * r0 * nalm), nalm,
                )
            )
        if tmp is not None:
            data += np.asarray(tmp).reshape(data.shape)

    def transform(self, data, spin: int = 0):
        """Does nothing, since inputs are alms already."""
        return data

    def resample(self, data):
        """Change LMAX of alm."""
        *dims, n = data.shape
        lmax_in = (int((8 * n + 1) ** 0.5 + 0.01) - 3) // 2
        lmax_out = self.__lmax
        lmax = min(lmax_in, lmax_out)
        out = np.zeros((*dims, (lmax_out + 1) * (lmax_out + 2) // 2), dtype=self.__dtype)
        src = np.asarray(data)
        i = j = 0
        for m in range(lmax + 1):
            out[..., j : j + lmax - m + 1] = src[..., i : i + lmax - m + 1]
            i += lmax_in - m + 1
            j += lmax_out - m + 1
        return out
