"""ctypes binding of libwdno_b200.so (the C ABI in include/wdno_b200.h).

The product path has no fallback: if the library is missing or cannot be loaded, importing any
engine entry point raises.  `lib()` builds nothing; use `wdno_b200.build.build()` / __graft_entry__.build().
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwdno_b200.so")


class Tap(C.Structure):
    _fields_ = [("kz", C.c_int32), ("shift", C.c_int32)]


class KSet(C.Structure):
    _fields_ = [("src", C.c_int32), ("ch_off", C.c_int32), ("ph_y", C.c_int32), ("ph_x", C.c_int32),
                ("tap_begin", C.c_int32), ("tap_count", C.c_int32)]


class NChunk(C.Structure):
    _fields_ = [("out_ch_off", C.c_int32), ("n_valid", C.c_int32), ("ph_y", C.c_int32), ("ph_x", C.c_int32),
                ("set_begin", C.c_int32), ("set_count", C.c_int32), ("n_tiles", C.c_int32), ("pad_", C.c_int32),
                ("w_tile_off", C.c_int64)]


class TapGemmParams(C.Structure):
    _fields_ = [
        ("src", C.c_void_p * 2), ("src_c", C.c_int32 * 2),
        ("coef_a", C.c_void_p * 2), ("coef_c", C.c_void_p * 2),
        ("src_mode", C.c_int32),
        ("B", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("KD", C.c_int32), ("pz", C.c_int32), ("py", C.c_int32), ("px", C.c_int32),
        ("Wp", C.c_int32), ("maxshift", C.c_int32),
        ("ZT", C.c_int32), ("PT", C.c_int32), ("KC", C.c_int32), ("N", C.c_int32), ("n_chunks", C.c_int32),
        ("chunks", C.c_void_p), ("sets", C.c_void_p), ("taps", C.c_void_p), ("wpacked", C.c_void_p),
        ("out_mode", C.c_int32), ("out_c", C.c_int32), ("out", C.c_void_p),
        ("bias", C.c_void_p), ("resid", C.c_void_p), ("stats", C.c_void_p),
        ("G", C.c_int32), ("cpg", C.c_int32),
        ("NSLOT", C.c_int32), ("NBST", C.c_int32), ("S_pad", C.c_int32), ("grid", C.c_int32),
        ("TPS", C.c_int32), ("reuse", C.c_int32), ("n_taps", C.c_int32), ("bias_len", C.c_int32),
        ("n_sets", C.c_int32), ("zstack", C.c_int32), ("strips", C.c_int32), ("Wfull", C.c_int32),
        ("fold", C.c_int32), ("cluster", C.c_int32),
    ]


_lib = None


def lib():
    """Load (once) and return the ctypes handle; raises RuntimeError if the CUDA extension is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the wdno_b200 CUDA extension is not built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.wdno_last_error.restype = C.c_char_p
    L.wdno_version.restype = C.c_int
    L.wdno_device_cc.restype = C.c_int
    L.wdno_tapgemm_smem_bytes.restype = C.c_int64
    L.wdno_tapgemm_smem_bytes.argtypes = [C.POINTER(TapGemmParams)]
    L.wdno_tapgemm_max_cluster_ctas.restype = C.c_int
    L.wdno_tapgemm_max_cluster_ctas.argtypes = [C.c_int64]
    L.wdno_tapgemm.restype = C.c_int
    L.wdno_tapgemm.argtypes = [C.POINTER(TapGemmParams), C.c_void_p]
    _bind_rest(L)
    _lib = L
    return L


def _bind_rest(L):
    """argtypes for the non-struct entry points (filled as the ABI grows; see include/wdno_b200.h)."""
    from . import _abi
    _abi.bind(L)


def check(rc, what=""):
    if rc != 0:
        msg = lib().wdno_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"wdno_b200 {what}: {msg}")
        raise RuntimeError(f"wdno_b200 {what} failed (rc={rc}): {msg}")


def current_stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
