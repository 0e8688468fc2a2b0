"""Loader for oracle/cmath_ref.c (see its header) -- TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "cmath_ref.c")
_SO = os.path.join(_HERE, "libdfb_oracle_cmath.so")
_lib = None


def build(force=False):
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fno-builtin", _SRC, "-o", _SO, "-lm"])
    return _SO


def powf2(x):
    """Elementwise libm powf(x, 2.0f) on a float32 array (== numpy float32 scalar `x ** 2`)."""
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.dfb_oracle_powf2.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    a = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(a)
    _lib.dfb_oracle_powf2(a.ctypes.data, out.ctypes.data, a.size)
    return out


def pow2(x):
    """Elementwise libm pow(x, 2.0) on a float64 array (== numpy float64 scalar `x ** 2`)."""
    global _lib
    if _lib is None:
        powf2(np.zeros(1, np.float32))
    _lib.dfb_oracle_pow2.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    a = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(a)
    _lib.dfb_oracle_pow2(a.ctypes.data, out.ctypes.data, a.size)
    return out
