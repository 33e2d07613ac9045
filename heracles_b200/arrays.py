"""
``DeviceArray`` -- an ``np.ndarray`` subclass whose buffer is CUDA managed memory.

Why: the reference's Field layer treats the object returned by
``Mapper.create()`` as a plain numpy array -- it divides and subtracts in place
(``heracles/fields.py:296,304,373,446,548``) and ``update_metadata`` assigns
``array.dtype`` (``heracles/core.py:102-122``).  A managed buffer wrapped as a
genuine ndarray satisfies all of that while the scatter and transform kernels
work on the same memory on the device; the common in-place ufuncs are
intercepted and run as device kernels so the map never migrates to the host.
"""

from __future__ import annotations

import numpy as np

from . import _lib


class _ManagedBuffer:
    """owns one hcu_malloc_managed allocation"""

    __slots__ = ("ctx", "ptr", "nbytes", "host_dirty", "__weakref__")

    def __init__(self, ctx: _lib.Context, nbytes: int):
        self.ctx = ctx
        self.nbytes = int(nbytes)
        self.ptr = ctx.malloc_managed(max(self.nbytes, 8))
        self.host_dirty = False

    @property
    def __array_interface__(self):
        # numpy keeps THIS object as the base of every view of the buffer, so the
        # allocation lives exactly as long as any array that points into it
        return {"data": (self.ptr, False), "shape": (max(self.nbytes, 8),), "typestr": "|u1", "version": 3}

    def __del__(self):
        try:
            if self.ptr and self.ctx.handle:
                self.ctx.free(self.ptr)
        except Exception:  # pragma: no cover - interpreter shutdown
            pass
        self.ptr = None


def update_metadata(array, *sources, **metadata):
    """same contract as heracles.core.update_metadata (core.py:100-122): ``sources`` are objects with a ``metadata``
    mapping (catalogues, fields.py:312); arrays are accepted too and contribute their ``dtype.metadata``"""
    md = {}
    if array.dtype.metadata is not None:
        md.update(array.dtype.metadata)
    for source in sources:
        smd = getattr(source, "metadata", None)
        if smd is None and isinstance(source, np.ndarray):
            smd = source.dtype.metadata
        if smd:
            md.update(smd)
    md.update(metadata)
    if array.dtype.fields is not None:
        dt = array.dtype.fields
    else:
        dt = array.dtype.str
    dt = np.dtype(dt, metadata=md)
    if not np.can_cast(dt, array.dtype, casting="no"):
        msg = "array with unsupported dtype"
        raise ValueError(msg)
    array.dtype = dt


class DeviceArray(np.ndarray):
    """ndarray over CUDA managed memory; see module docstring"""

    _hcu: _ManagedBuffer | None = None

    # -- construction -----------------------------------------------------
    @classmethod
    def zeros(cls, ctx: _lib.Context, shape, dtype=np.float64) -> "DeviceArray":
        dtype = np.dtype(dtype)
        shape = tuple(int(s) for s in np.atleast_1d(shape)) if not isinstance(shape, tuple) else shape
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        owner = _ManagedBuffer(ctx, nbytes)
        ctx.prefetch(owner.ptr, max(nbytes, 8), True)
        ctx.memset_zero(owner.ptr, max(nbytes, 8))
        ctx.synchronize()
        count = int(np.prod(shape, dtype=np.int64))
        arr = np.asarray(owner)[: count * dtype.itemsize].view(dtype).reshape(shape)
        out = arr.view(cls)
        out._hcu = owner
        return out

    def __array_finalize__(self, obj):
        owner = getattr(obj, "_hcu", None)
        if owner is not None and owner.ptr:
            # views share the buffer; results allocated by numpy do not
            p = self.__array_interface__["data"][0]
            if not (owner.ptr <= p < owner.ptr + max(owner.nbytes, 8)):
                owner = None
        self._hcu = owner

    # -- helpers ----------------------------------------------------------
    @property
    def device_ptr(self) -> int | None:
        """address usable by the kernels, or None when not managed / not contiguous"""
        if self._hcu is None or not self.flags.c_contiguous:
            return None
        return self.__array_interface__["data"][0]

    def _host(self, write: bool = False) -> np.ndarray:
        """plain ndarray view for host-side numpy; synchronises the device first"""
        if self._hcu is not None:
            self._hcu.ctx.synchronize()
            if write:
                self._hcu.host_dirty = True
        return self.view(np.ndarray)

    def to_device(self) -> None:
        """bring pages back to the device after host writes"""
        if self._hcu is not None and self._hcu.host_dirty:
            self._hcu.ctx.prefetch(self._hcu.ptr, max(self._hcu.nbytes, 8), True)
            self._hcu.host_dirty = False

    # -- in-place arithmetic on the device ----------------------------------
    def _device_inplace(self, ufunc, other) -> bool:
        ptr = self.device_ptr
        if ptr is None or self.dtype != np.float64:
            return False
        ctx = self._hcu.ctx
        lib = ctx.lib
        n = self.size
        h = ctx.handle
        if np.isscalar(other) or (isinstance(other, np.ndarray) and other.ndim == 0):
            a = float(other)
            self.to_device()
            if ufunc is np.true_divide:
                _lib.check(lib.hcu_divide(h, ptr, n, a))
            elif ufunc is np.multiply:
                _lib.check(lib.hcu_scale(h, ptr, n, a))
            elif ufunc is np.add:
                _lib.check(lib.hcu_add_scalar(h, ptr, n, a))
            elif ufunc is np.subtract:
                _lib.check(lib.hcu_add_scalar(h, ptr, n, -a))
            else:
                return False
            ctx.synchronize()
            return True
        if (
            isinstance(other, DeviceArray)
            and other.device_ptr is not None
            and other.dtype == np.float64
            and other.shape == self.shape
            and ufunc in (np.add, np.subtract)
        ):
            self.to_device()
            other.to_device()
            a = 1.0 if ufunc is np.add else -1.0
            _lib.check(lib.hcu_axpy(h, ptr, other.device_ptr, a, n))
            ctx.synchronize()
            return True
        return False

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        if (
            method == "__call__"
            and out is not None
            and len(out) == 1
            and out[0] is self
            and len(inputs) == 2
            and inputs[0] is self
            and not kwargs
        ):
            if self._device_inplace(ufunc, inputs[1]):
                return self
        # generic path: run numpy on host views of the managed memory
        conv = [x._host() if isinstance(x, DeviceArray) else x for x in inputs]
        if out is not None:
            kwargs["out"] = tuple(x._host(write=True) if isinstance(x, DeviceArray) else x for x in out)
        res = getattr(ufunc, method)(*conv, **kwargs)
        if out is not None and len(out) == 1 and res is kwargs["out"][0]:
            return out[0]
        return res

    def __array_function__(self, func, types, args, kwargs):
        if self._hcu is not None:
            self._hcu.ctx.synchronize()
        return super().__array_function__(func, types, args, kwargs)

    def __getitem__(self, key):
        if self._hcu is not None:
            self._hcu.ctx.synchronize()
        return super().__getitem__(key)

    def __setitem__(self, key, value):
        if self._hcu is not None:
            self._hcu.ctx.synchronize()
            self._hcu.host_dirty = True
        if isinstance(value, DeviceArray):
            value = value.view(np.ndarray)
        super().__setitem__(key, value)

    def __reduce__(self):
        # pickling copies to a plain array
        return np.asarray(self._host()).copy().__reduce__()

    def copy(self, order="C"):
        return np.array(self._host(), copy=True, order=order)


def as_device_pointer(arr) -> int | None:
    """device-usable address of ``arr`` if it is a contiguous DeviceArray"""
    if isinstance(arr, DeviceArray):
        return arr.device_ptr
    return None
