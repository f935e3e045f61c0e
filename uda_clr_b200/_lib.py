"""ctypes binding of ``libclr_b200.so`` (the C ABI in ``include/clr_b200.h``).

There is no CPU fallback: if the library is missing or a call fails, this raises.  The library is
loaded lazily so that CPU-only tooling (the oracle tests, ``build()``) can import the package.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_size_t, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libclr_b200.so")

CLR_W_COMPLEMENT = 0
CLR_W_EXPLICIT = 1
CLR_MAX_K = 8
CLR_ERR_UNSUPPORTED = -4
CLR_MAX_WORLD = 8

_lock = threading.Lock()
_lib = None


class ClrError(RuntimeError):
    pass


_P = c_void_p  # device pointers travel as integers


class BwdDom(Structure):
    """``clr_bwd_dom`` (include/clr_b200.h)."""
    _fields_ = [("w", _P), ("g", _P), ("sums", _P), ("xcoef", _P), ("xtab", _P), ("grad", _P),
                ("scale_dev", _P), ("scale", c_float), ("fmt", c_int), ("B", c_int), ("Kx", c_int)]


class StepArgs(Structure):
    """``clr_step_args`` (include/clr_b200.h) -- field order must match the header."""
    _fields_ = [
        ("B_s", c_int), ("B_t", c_int), ("C", c_int), ("H", c_int), ("W", c_int), ("K", c_int),
        ("Hi", c_int), ("Wi", c_int), ("T", c_int),
        ("use_retrify", c_int), ("use_disc", c_int), ("use_cons", c_int),
        ("wt_fmt", c_int), ("first_s", c_int), ("first_t", c_int),
        ("decay", c_double), ("npx_global", c_double),
        ("w_intra", c_float), ("w_inter", c_float), ("w_disc", c_float), ("w_aug", c_float), ("margin", c_float),
        ("aug_weight", c_float), ("cons_threshold", c_float), ("pseudo_thr", c_float), ("std_thr", c_float),
        ("grad_scale", c_float),
        ("xs", _P), ("ys", _P), ("xt", _P), ("wt", _P), ("oT_before", _P), ("preds", _P), ("oT", _P), ("oT_aug", _P),
        ("gup", _P),
        ("stored_s", _P), ("stored_t", _P),
        ("packed1", _P), ("packed2", _P), ("P_s", _P), ("P_t", _P), ("g_s", _P), ("g_t", _P), ("losses", _P),
        ("std_map", _P), ("pred_mean", _P), ("wt_retrify", _P), ("masks", _P),
        ("disc_coef", _P), ("disc_vec", _P), ("disc_beta", _P), ("xtab", _P),
        ("gxs", _P), ("gxt", _P), ("g_oT_aug", _P),
        ("ws", _P), ("ws_bytes", c_size_t),
        ("ev_pool_begin", _P), ("ev_pool_end", _P), ("ev_bwd_begin", _P), ("ev_bwd_end", _P),
        ("reserved_ptr", _P * 3),
        ("world", c_int), ("rank", c_int), ("seq", ctypes.c_uint), ("reserved0", c_int), ("peer_rx", _P * 8),
    ]


_SIGNATURES = {
    "clr_version": (c_int, []),
    "clr_status_string": (c_char_p, [c_int]),
    "clr_device_info": (c_int, [POINTER(c_int), POINTER(c_int)]),
    "clr_set_tunable": (c_int, [c_char_p, c_int]),
    "clr_launch_count": (ctypes.c_ulonglong, []),
    "clr_event_create": (c_int, [POINTER(c_void_p)]),
    "clr_event_destroy": (c_int, [c_void_p]),
    "clr_event_elapsed_us": (c_int, [c_void_p, c_void_p, POINTER(c_float)]),
    "clr_peer_alloc": (c_int, [c_size_t, POINTER(c_void_p)]),
    "clr_peer_free": (c_int, [c_void_p]),
    "clr_peer_export": (c_int, [c_void_p, c_char_p]),
    "clr_peer_open": (c_int, [c_char_p, POINTER(c_void_p)]),
    "clr_peer_close": (c_int, [c_void_p]),
    "clr_step_xchg_bytes": (c_size_t, [c_int, c_int, c_int]),
    "clr_trace_enable": (c_int, [c_int]),
    "clr_trace_slots": (c_int, []),
    "clr_trace_name": (c_char_p, [c_int]),
    "clr_trace_read": (c_int, [c_void_p]),
    "clr_pool_ws_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "clr_pool_rows_ws_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "clr_pool_rows_fwd": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, c_size_t, _P, _P]),
    "clr_pool_rows_fwd_ps": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, c_size_t, _P, _P]),
    "clr_bmm_finalize": (c_int, [_P, c_int, c_int, c_int, c_float, _P, _P]),
    "clr_pool_bwd_ps": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, c_float, c_float, _P, _P]),
    "clr_pool_fwd": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P, _P]),
    "clr_pool_fwd_mu": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P, _P, _P]),
    "clr_pool_fwd2": (c_int, [_P, _P, c_int, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P, _P, _P]),
    "clr_proto_finalize": (c_int, [_P, c_int, c_int, _P, _P]),
    "clr_pool_bwd": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, c_float, _P, _P, c_int, _P, _P]),
    "clr_pool_bwd_multi": (c_int, [POINTER(BwdDom), c_int, c_int, c_int, c_int, _P]),
    "clr_pixel_dots": (c_int, [_P, c_int, c_int, c_int, _P, c_int, _P, _P, _P]),
    "clr_pool_bwd_w_ws_bytes": (c_size_t, [c_int, c_int, c_int]),
    "clr_pool_bwd_w": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, c_float, _P, c_size_t, _P, _P]),
    "clr_proto_distance": (c_int, [_P, c_int, c_int, c_int, _P, c_int, _P, _P]),
    "clr_proto_cosine": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, _P]),
    "clr_minmax_normalize": (c_int, [_P, c_size_t, _P, _P]),
    "clr_disc_partials_cap": (c_int, []),
    "clr_disc_fwd": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P, c_float, _P, _P, _P, c_int,
                             POINTER(c_int), _P]),
    "clr_disc_fused_ws_bytes": (c_size_t, [c_int, c_int]),
    "clr_disc_fused_fwd": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P, c_float, _P, _P, _P, c_size_t, _P, _P]),
    "clr_mc_stats": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "clr_retrify_weights": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
                                    _P, _P, _P, _P, _P]),
    "clr_mc_state_floats": (c_size_t, [c_int, c_int, c_int, c_int]),
    "clr_mc_accumulate": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "clr_mc_finalize": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "clr_label_downsample": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "clr_ema_rows": (c_int, [_P, _P, c_int, c_int, c_float, _P, _P]),
    "clr_mc_retrify": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, _P, _P, _P, _P, _P]),
    "clr_seg_loss_ws_bytes": (c_size_t, []),
    "clr_seg_loss_fwd": (c_int, [_P, _P, c_size_t, _P, _P, c_size_t, _P, c_size_t, _P, _P]),
    "clr_seg_loss_bwd": (c_int, [_P, _P, c_size_t, _P, _P, c_size_t, _P, c_float, _P, _P, _P]),
    "clr_entropy_fwd": (c_int, [_P, c_size_t, c_float, _P, _P]),
    "clr_entropy_bwd": (c_int, [_P, _P, c_size_t, c_float, _P, _P]),
    "clr_seg_counts": (c_int, [_P, _P, c_int, c_int, c_size_t, c_float, _P, _P]),
    "clr_tn_ws_bytes": (c_size_t, [c_int]),
    "clr_tn_fwd": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, c_float, c_float, _P, c_size_t, _P, _P, _P]),
    "clr_tn_eval": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, c_float, _P, c_size_t, _P, _P, _P]),
    "clr_tn_bwd": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, c_int, _P, c_size_t, _P, _P, _P, _P]),
    "clr_cons_ws_bytes": (c_size_t, []),
    "clr_cons_fwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, _P, c_size_t,
                             _P, _P]),
    "clr_cons_bwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, _P, _P,
                             c_float, _P, _P]),
    "clr_align_finalize": (c_int, [_P, _P, c_int, c_int, _P, _P, c_int, c_int, c_double, c_float, c_float,
                                   _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "clr_disc_finalize": (c_int, [_P, _P, c_int, c_int, c_double, c_float, c_float, c_float, _P, _P,
                                  c_float, c_float, c_float, c_float, c_int, c_int, _P, _P]),
    "clr_step_ws_bytes": (c_size_t, [POINTER(StepArgs)]),
    "clr_step_schedule": (c_int, [POINTER(StepArgs)]),
    "clr_step_fwd_a": (c_int, [POINTER(StepArgs), _P]),
    "clr_step_fwd_b": (c_int, [POINTER(StepArgs), _P]),
    "clr_step_fwd_c": (c_int, [POINTER(StepArgs), _P]),
    "clr_step_fwd": (c_int, [POINTER(StepArgs), _P]),
    "clr_step_bwd": (c_int, [POINTER(StepArgs), _P]),
    "clr_step_run": (c_int, [POINTER(StepArgs), _P]),
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
        # a library older than its sources (edited tree, interrupted build) is rebuilt when nvcc is at hand; a missing
        # library without a compiler is an error -- there is no CPU fallback for the CLR ops
        try:
            from . import build as _build
            if not _build.is_fresh():
                _build.build()
        except Exception as e:
            if not os.path.isfile(LIB_PATH):
                raise ClrError("libclr_b200.so not built and could not be built (%s): run `python -m uda_clr_b200.build` "
                               "(there is no CPU fallback for the CLR ops)" % str(e)[:200])
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


class Event:
    """A CUDA timing event owned by the library (recorded by the library around its own launches)."""

    def __init__(self):
        h = c_void_p()
        check(load().clr_event_create(ctypes.byref(h)), "clr_event_create")
        self.handle = h.value

    def elapsed_us(self, end: "Event") -> float:
        us = c_float()
        check(load().clr_event_elapsed_us(self.handle, end.handle, ctypes.byref(us)), "clr_event_elapsed_us")
        return float(us.value)

    def __del__(self):
        try:
            if self.handle and _lib is not None:
                _lib.clr_event_destroy(self.handle)
        except Exception:
            pass


def ptr(t):
    """Device pointer of a tensor (or None -> NULL)."""
    return None if t is None else t.data_ptr()
