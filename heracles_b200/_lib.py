"""
ctypes binding of ``libheracles_cuda.so`` (C ABI declared in
``include/heracles_cuda.h``).

There is no CPU fallback: if the shared library is missing, or no CUDA device
is present when a context is requested, this module raises.
"""

from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# HERACLES_CUDA_LIB points at an alternative build of the SAME library (kernel tuning experiments)
LIB_PATH = os.environ.get("HERACLES_CUDA_LIB") or os.path.join(_HERE, "lib", "libheracles_cuda.so")

c_int = ctypes.c_int
c_i64 = ctypes.c_int64
c_size = ctypes.c_size_t
c_dbl = ctypes.c_double
c_vp = ctypes.c_void_p

# name -> (restype, argtypes); every symbol include/heracles_cuda.h declares
SIGNATURES = {
    "hcu_version": (c_int, []),
    "hcu_last_error": (ctypes.c_char_p, []),
    "hcu_device_count": (c_int, [ctypes.POINTER(c_int)]),
    "hcu_create": (c_int, [c_int, ctypes.POINTER(c_vp)]),
    "hcu_destroy": (c_int, [c_vp]),
    "hcu_set_stream": (c_int, [c_vp, c_vp]),
    "hcu_synchronize": (c_int, [c_vp]),
    "hcu_launch_count": (c_int, [c_vp, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]),
    "hcu_trim": (c_int, [c_vp]),
    "hcu_malloc_managed": (c_int, [c_vp, c_size, ctypes.POINTER(c_vp)]),
    "hcu_malloc_device": (c_int, [c_vp, c_size, ctypes.POINTER(c_vp)]),
    "hcu_malloc_pinned": (c_int, [c_vp, c_size, ctypes.POINTER(c_vp)]),
    "hcu_free": (c_int, [c_vp, c_vp]),
    "hcu_prefetch": (c_int, [c_vp, c_vp, c_size, c_int]),
    "hcu_memset_zero": (c_int, [c_vp, c_vp, c_size]),
    "hcu_memcpy": (c_int, [c_vp, c_vp, c_vp, c_size]),
    "hcu_ang2pix": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_i64, c_vp]),
    "hcu_map_values": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_i64, c_int, c_i64, c_vp, c_i64, c_int]),
    "hcu_bad_rows": (c_int, [c_vp, ctypes.POINTER(c_i64)]),
    "hcu_scale": (c_int, [c_vp, c_vp, c_i64, c_dbl]),
    "hcu_divide": (c_int, [c_vp, c_vp, c_i64, c_dbl]),
    "hcu_axpy": (c_int, [c_vp, c_vp, c_vp, c_dbl, c_i64]),
    "hcu_add_scalar": (c_int, [c_vp, c_vp, c_i64, c_dbl]),
    "hcu_ud_grade": (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp]),
    "hcu_map2alm": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_vp, c_i64, c_vp, c_vp, c_int, c_vp, c_vp, c_i64]),
    "hcu_map2alm_many": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp]),
    "hcu_alm2map": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_vp, c_i64, c_vp, c_i64]),
    "hcu_legendre_batch_size": (c_int, [c_int]),
    "hcu_map2phase": (c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_i64, c_vp, c_i64, c_i64, c_vp, c_int, c_vp]),
    "hcu_alm2phase": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_vp, c_i64, c_vp, c_int, c_i64, c_i64, c_vp]),
    "hcu_phase2alm_blocks": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_i64]),
    "hcu_alm2phase_blocks": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_vp, c_i64, c_vp, c_int, c_int, c_vp, c_vp]),
    "hcu_phase2map": (c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_i64, c_i64, c_vp, c_i64]),
    "hcu_map2phase_peers": (c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_i64, c_vp, c_i64, c_i64, c_vp, c_int, c_int, c_vp, c_vp]),
    "hcu_alm2phase_peers": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_vp, c_i64, c_vp, c_int, c_int, c_vp, c_vp]),
    "hcu_ipc_export": (c_int, [c_vp, c_vp, c_vp]),
    "hcu_ipc_open": (c_int, [c_vp, c_vp, ctypes.POINTER(c_vp)]),
    "hcu_ipc_close": (c_int, [c_vp, c_vp]),
    "hcu_points2alm": (c_int, [c_vp, c_int, c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64]),
    "hcu_phase2alm": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp, c_int, c_i64, c_i64, c_vp, c_vp, c_i64]),
    "hcu_alm2cl": (c_int, [c_vp, c_int, c_vp, c_i64, c_int, c_int, c_vp, c_i64, c_int, c_int, c_vp]),
    "hcu_alm2cl_rows": (c_int, [c_vp, c_int, ctypes.POINTER(c_vp), c_int, c_int, c_vp]),
    "hcu_alm2cl_mslice": (c_int, [c_vp, c_int, c_vp, c_i64, c_int, c_int, c_vp, c_i64, c_int, c_int, c_int, c_int, c_vp]),
    "hcu_map_page": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp]),
    "hcu_reorder": (c_int, [c_vp, c_i64, c_vp, c_vp, c_int]),
    "hcu_set_timing": (c_int, [c_vp, c_int]),
    "hcu_set_start_table": (c_int, [c_vp, c_int]),
    "hcu_set_weights_mode": (c_int, [c_vp, c_int]),
    "hcu_multiply": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64]),
    "hcu_region_select": (c_int, [c_vp, c_vp, c_vp, c_vp, c_dbl, c_i64]),
    "hcu_last_sht_timing": (c_int, [c_vp, ctypes.POINTER(ctypes.c_float * 4)]),
    "hcu_last_sht_work": (c_int, [c_vp, ctypes.POINTER(c_dbl), ctypes.POINTER(c_dbl)]),
    "hcu_measure_fp64_peak": (c_int, [c_vp, ctypes.POINTER(c_dbl)]),
}

_lib = None
_lock = threading.Lock()


class HeraclesCudaError(RuntimeError):
    """an hcu_* call returned a failure status"""


def load():
    """dlopen libheracles_cuda.so and declare every prototype"""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "or `make -C heracles_b200/csrc` (there is no CPU fallback)"
                )
            lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(status: int) -> None:
    if status != 0:
        msg = load().hcu_last_error().decode("utf-8", "replace")
        if status == -4:
            raise NotImplementedError(msg)
        if status == -2:
            raise ValueError(msg)
        if status == -3:
            raise MemoryError(msg)
        raise HeraclesCudaError(f"heracles_cuda error {status}: {msg}")


def device_count() -> int:
    n = c_int(0)
    check(load().hcu_device_count(ctypes.byref(n)))
    return n.value


class Context:
    """one hcu_ctx: a device, a stream, cached tables and workspaces"""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = c_vp()
        check(self.lib.hcu_create(int(device), ctypes.byref(h)))
        self.handle = h
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.hcu_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover - interpreter shutdown order
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing -------------------------------------------------------
    def set_stream(self, cuda_stream: int | None):
        check(self.lib.hcu_set_stream(self.handle, c_vp(cuda_stream or 0)))

    def synchronize(self):
        check(self.lib.hcu_synchronize(self.handle))

    def launch_count(self):
        a, b = c_i64(0), c_i64(0)
        check(self.lib.hcu_launch_count(self.handle, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def trim(self):
        check(self.lib.hcu_trim(self.handle))

    def malloc_managed(self, nbytes: int) -> int:
        p = c_vp()
        check(self.lib.hcu_malloc_managed(self.handle, nbytes, ctypes.byref(p)))
        return p.value

    def malloc_device(self, nbytes: int) -> int:
        p = c_vp()
        check(self.lib.hcu_malloc_device(self.handle, nbytes, ctypes.byref(p)))
        return p.value

    def malloc_pinned(self, nbytes: int) -> int:
        p = c_vp()
        check(self.lib.hcu_malloc_pinned(self.handle, nbytes, ctypes.byref(p)))
        return p.value

    def free(self, ptr: int):
        if self.handle:
            check(self.lib.hcu_free(self.handle, c_vp(ptr)))

    def prefetch(self, ptr: int, nbytes: int, to_device: bool = True):
        check(self.lib.hcu_prefetch(self.handle, c_vp(ptr), nbytes, int(to_device)))

    def memset_zero(self, ptr: int, nbytes: int):
        check(self.lib.hcu_memset_zero(self.handle, c_vp(ptr), nbytes))

    def memcpy(self, dst: int, src: int, nbytes: int):
        check(self.lib.hcu_memcpy(self.handle, c_vp(dst), c_vp(src), nbytes))

    def bad_rows(self) -> int:
        n = c_i64(0)
        check(self.lib.hcu_bad_rows(self.handle, ctypes.byref(n)))
        return n.value

    def set_timing(self, enabled: bool = True) -> None:
        """CUDA-event timing of the SHT stages (costs a host wait per Legendre batch; off by default)"""
        check(self.lib.hcu_set_timing(self.handle, 1 if enabled else 0))

    def sht_timing(self):
        arr = (ctypes.c_float * 4)()
        check(self.lib.hcu_last_sht_timing(self.handle, ctypes.byref(arr)))
        return list(arr)

    def sht_work(self):
        a, b = c_dbl(0), c_dbl(0)
        check(self.lib.hcu_last_sht_work(self.handle, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def fp64_peak(self) -> float:
        f = c_dbl(0)
        check(self.lib.hcu_measure_fp64_peak(self.handle, ctypes.byref(f)))
        return f.value


_contexts: dict[int, Context] = {}


_extra: dict = {}


def extra_context(device: int, index: int) -> Context:
    """additional library contexts on a device (own stream, workspaces, cuFFT plans): the lanes of the multi-GPU path"""
    with _lock:
        ctx = _extra.get((device, index))
    if ctx is None:
        ctx = Context(device)
        with _lock:
            _extra[(device, index)] = ctx
    return ctx


def get_context(device: int | None = None) -> Context:
    """process-wide context per device (LOCAL_RANK selects the default device)"""
    if device is None:
        device = int(os.environ.get("HERACLES_CUDA_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    with _lock:
        ctx = _contexts.get(device)
    if ctx is None:
        ctx = Context(device)
        with _lock:
            _contexts[device] = ctx
    return ctx
