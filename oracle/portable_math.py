"""ctypes front-end of oracle/portable_math.c (host build of die_b200/csrc/die_math.h).
TEST INFRASTRUCTURE -- see oracle/die_ref.py."""
import ctypes
import os

import numpy as np

_lib = None
_dp = ctypes.POINTER(ctypes.c_double)


def _load():
    global _lib
    if _lib is None:
        from oracle import build_oracle
        path = build_oracle.LIB if os.path.exists(build_oracle.LIB) else build_oracle.build()
        _lib = ctypes.CDLL(path)
    return _lib


def sincos(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    s, c = np.empty_like(x), np.empty_like(x)
    _load().die_sincos_array(x.ctypes.data_as(_dp), s.ctypes.data_as(_dp), c.ctypes.data_as(_dp), ctypes.c_long(x.size))
    return s, c


def atan2(y, x, fast=False):
    y, x = np.broadcast_arrays(np.asarray(y, dtype=np.float64), np.asarray(x, dtype=np.float64))
    y, x = np.ascontiguousarray(y), np.ascontiguousarray(x)
    out = np.empty_like(x)
    fn = _load().die_atan2_fast_array if fast else _load().die_atan2_array
    fn(y.ctypes.data_as(_dp), x.ctypes.data_as(_dp), out.ctypes.data_as(_dp), ctypes.c_long(x.size))
    return out


def sincos_angle(x):
    """(sin x, cos x, atan2(sin x, cos x)) for |x| <= pi -- die_sincos_angle."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    s, c, a = np.empty_like(x), np.empty_like(x), np.empty_like(x)
    _load().die_sincos_angle_array(x.ctypes.data_as(_dp), s.ctypes.data_as(_dp), c.ctypes.data_as(_dp),
                                   a.ctypes.data_as(_dp), ctypes.c_long(x.size))
    return s, c, a
