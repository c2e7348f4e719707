"""lib/decimal/decimal.go mirror."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib


def append_decimal_to_float(va, e, ctx=None):
    """decimal.AppendDecimalToFloat decimal.go:100 (GPU)"""
    ctx = ctx or _lib.default_context()
    a = np.ascontiguousarray(va, dtype=np.int64)
    dst = np.empty(a.size, dtype=np.float64)
    check(lib().vmb_decimal_to_float(ctx.h, dst.ctypes.data_as(_lib.f64p), a.ctypes.data_as(_lib.i64p), a.size, int(e)))
    return dst


def append_float_to_decimal(src):
    """decimal.AppendFloatToDecimal decimal.go:173 (host, write path) -> (np.int64[n], scale)"""
    f = np.ascontiguousarray(src, dtype=np.float64)
    dst = np.empty(f.size, dtype=np.int64)
    e = C.c_int16(0)
    check(lib().vmb_float_to_decimal(dst.ctypes.data_as(_lib.i64p), C.byref(e), f.ctypes.data_as(_lib.f64p), f.size))
    return dst, e.value


def append_float_to_decimal_columns(src2d, ctx=None):
    """decimal.AppendFloatToDecimal for equal-length columns on the GPU (vmb_float_to_decimal_columns): src2d [ncols x rows] float64
    -> (np.int64[ncols, rows], np.int16[ncols] scales)"""
    ctx = ctx or _lib.default_context()
    f = np.ascontiguousarray(src2d, dtype=np.float64)
    ncols, rows = f.shape
    dst = np.empty((ncols, rows), dtype=np.int64)
    scales = np.zeros(ncols, dtype=np.int16)
    check(lib().vmb_float_to_decimal_columns(ctx.h, dst.ctypes.data_as(_lib.i64p), scales.ctypes.data_as(C.POINTER(C.c_int16)),
                                             f.ctypes.data_as(_lib.f64p), ncols, rows))
    return dst, scales


def calibrate_scale(a, ae, b, be):
    """decimal.CalibrateScale decimal.go:13 (host, merge path) -> (a', b', e): both arrays rescaled to the common exponent e"""
    a = np.array(a, dtype=np.int64)
    b = np.array(b, dtype=np.int64)
    e = C.c_int16(0)
    check(lib().vmb_calibrate_scale(a.ctypes.data_as(_lib.i64p), a.size, int(ae), b.ctypes.data_as(_lib.i64p), b.size, int(be),
                                    C.byref(e)))
    return a, b, e.value
