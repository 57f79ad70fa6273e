import ctypes

import numpy as np

from pyseistr_b200 import _lib

fp = ctypes.POINTER(ctypes.c_float)


def f32(a):
    """contiguous float32 1-D view/copy, like PyArray_FROM_OTF(NPY_FLOAT, IN_ARRAY) + flat indexing"""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))


def ptr(a):
    return a.ctypes.data_as(fp)


def ctx():
    return _lib.default_context()


def check(rc):
    _lib.check(rc)
