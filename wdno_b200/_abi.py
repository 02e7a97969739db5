"""Signatures of the plain (non-struct) C-ABI entry points, one line per symbol in include/wdno_b200.h."""
import ctypes as C

P = C.c_void_p
I = C.c_int32
L64 = C.c_int64
F = C.c_float

# name -> argtypes ; restype is always int
SIGNATURES = {}


def bind(lib):
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = argtypes
