"""ctypes binding of ``libclr_b200.so`` (the C ABI in ``include/clr_b200.h``).

There is no CPU fallback: if the library is missing or a call fails, this raises.  The library is
loaded lazily so that CPU-only tooling (the oracle tests, ``build()``) can import the package.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, c_char_p, c_float, c_int, c_size_t, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libclr_b200.so")

CLR_W_COMPLEMENT = 0
CLR_W_EXPLICIT = 1
CLR_MAX_K = 8

_lock = threading.Lock()
_lib = None


class ClrError(RuntimeError):
    pass


_P = c_void_p  # device pointers travel as integers

_SIGNATURES = {
    "clr_version": (c_int, []),
    "clr_status_string": (c_char_p, [c_int]),
    "clr_device_info": (c_int, [POINTER(c_int), POINTER(c_int)]),
    "clr_pool_ws_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "clr_pool_fwd": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P, _P]),
    "clr_pool_fwd2": (c_int, [_P, _P, c_int, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P, _P, _P]),
    "clr_proto_finalize": (c_int, [_P, c_int, c_int, _P, _P]),
    "clr_pool_bwd": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, c_float, _P, _P, c_int, _P, _P]),
    "clr_pool_bwd2": (c_int, [_P, c_int, c_int, _P, _P, c_float, _P, _P, c_int, _P,
                              _P, c_int, c_int, _P, _P, c_float, _P, c_int, c_int, c_int, _P]),
    "clr_pixel_dots": (c_int, [_P, c_int, c_int, c_int, _P, c_int, _P, _P, _P]),
    "clr_pool_bwd_w_ws_bytes": (c_size_t, [c_int, c_int, c_int]),
    "clr_pool_bwd_w": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, c_float, _P, c_size_t, _P, _P]),
}


def exported_symbols():
    """Every symbol ``include/clr_b200.h`` declares (checked by tests/test_abi.py)."""
    return sorted(_SIGNATURES)


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises :class:`ClrError` if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise ClrError("libclr_b200.so not built: run `python -m uda_clr_b200.build` "
                           "(there is no CPU fallback for the CLR ops)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().clr_status_string(status)
        raise ClrError("%s failed: %s (status %d)" % (what, msg.decode() if msg else "?", status))


def ptr(t):
    """Device pointer of a tensor (or None -> NULL)."""
    return None if t is None else t.data_ptr()
