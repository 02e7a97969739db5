"""Signatures of the plain (non-struct) C-ABI entry points, one line per symbol in include/wdno_b200.h."""
import ctypes as C

P = C.c_void_p
I = C.c_int32
L64 = C.c_int64
F = C.c_float
D = C.c_double


class CondOp(C.Structure):
    _fields_ = [("f0", I), ("f1", I), ("c0", I), ("c1", I), ("y0", I), ("y1", I), ("x0", I), ("x1", I),
                ("src", P), ("sb", L64), ("sf", L64), ("sc", L64), ("sy", L64), ("sx", L64),
                ("of", I), ("oc", I), ("oy", I), ("ox", I)]


MAX_COND_OPS = 8

# name -> argtypes ; restype is always int
SIGNATURES = {
    "wdno_pack_bfchw_f16": [P, P, I, I, I, I, I, I, P],
    "wdno_gn_finalize": [P, P, P, P, I, P, P, I, I, I, D, F, P],
    "wdno_gn_silu_add": [P, P, P, P, P, I, I, L64, P],
    "wdno_chan_layernorm": [P, P, P, P, L64, I, F, P],
    "wdno_time_mlp": [P, P, P, P, P, P, P, I, I, I, F, P],
    "wdno_small_linear": [P, P, P, P, I, I, I, P],
    "wdno_conv1x1": [P, I, P, I, P, I, P, P, P, L64, I, I, L64, P],
    "wdno_softmax_attn": [P, P, P, P, P, L64, I, L64, L64, L64, L64, F, P],
    "wdno_linear_attn": [P, P, L64, I, F, P],
    "wdno_linattn_block": [P, P, P, P, P, P, P, P, L64, I, I, F, F, P],
    "wdno_linattn_block_tc": [P, P, P, P, P, P, P, L64, I, I, F, F, P],
    "wdno_linattn_tc_supported": [L64, I, I],
    "wdno_tattn_block_tc": [P, P, P, P, P, P, P, P, L64, I, L64, I, F, F, P],
    "wdno_tattn_block_row": [P, P, P, P, P, P, P, L64, I, L64, I, F, F, P],
    "wdno_tattn_block": [P, P, P, P, P, P, P, P, P, L64, I, L64, I, F, F, P],
    "wdno_ddim_step": [P, P, P, P, P, P, I, I, I, I, I, I, I, P],
    "wdno_ddpm_step": [P, P, P, P, P, P, I, I, I, I, I, I, I, P],
    "wdno_apply_conditions": [P, P, I, I, I, I, I, I, P],
    "wdno_predict_x0": [P, P, P, P, L64, I, P],
    "wdno_q_sample": [P, P, P, P, P, P, I, L64, P],
    "wdno_mse_weighted": [P, P, P, I, I, I, I, I, I, P, P],
    "wdno_step_begin": [P, P, P, P, P, I, I, P],
    "wdno_gn_bwd_reduce": [P, P, P, P, P, I, I, L64, P],
    "wdno_gn_bwd_finalize": [P, P, P, P, P, I, P, P, P, I, P, P, I, I, I, D, F, F, P],
    "wdno_gn_bwd_apply": [P, P, P, P, P, P, P, P, I, I, I, L64, P],
    "wdno_pack_grad_f16": [P, P, I, I, I, I, I, I, F, P],
    "wdno_add_f16": [P, P, P, L64, P],
    "wdno_chan_layernorm_bwd": [P, P, P, P, P, P, L64, I, F, F, P],
    "wdno_softmax_attn_bwd": [P, P, P, P, P, P, P, L64, I, L64, L64, L64, L64, F, P],
    "wdno_linear_attn_bwd": [P, P, P, P, L64, I, F, P],
    "wdno_upsample2x_f16": [P, P, L64, I, I, I, P],
    "wdno_sumpool2x2_f16": [P, P, L64, I, I, I, P],
    "wdno_colsum_f16": [P, L64, I, P, F, P],
    "wdno_sumsq": [P, L64, P, P],
    "wdno_adam_clip_ema": [P, P, P, P, P, L64, P, F, F, F, F, F, F, F, F, I, P],
    "wdno_randn_slice": [P, L64, L64, L64, I, C.c_uint64, C.c_uint64, P],
    "wdno_dwt_analysis_axis": [P, P, P, L64, I, L64, I, L64, L64, L64, P, P, I, I, I, P],
    "wdno_dwt_synthesis_axis": [P, P, P, L64, I, L64, I, L64, L64, L64, P, P, I, I, I, P],
    "wdno_dwt3d_supported": [I, I, I],
    "wdno_dwt3d_synthesis": [P, L64, P, L64, I, I, I, I, I, I, P, P, I, I, P],
    "wdno_dwt3d_analysis": [P, P, L64, L64, I, I, I, I, I, I, P, P, I, I, P],
    "wdno_dwt2d_supported": [I, I, I, I, I, I],
    "wdno_dwt2d_analysis": [P, P, P, L64, I, I, I, I, P, P, I, I, I, I, P],
    "wdno_dwt2d_synthesis": [P, P, P, L64, I, I, I, I, P, P, I, I, I, I, P],
}


def bind(lib):
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = argtypes
    lib.wdno_linattn_work_bytes.restype = C.c_int64
    lib.wdno_linattn_work_bytes.argtypes = [L64, I, I]
    lib.wdno_linear_attn_bwd_work_bytes.restype = C.c_int64
    lib.wdno_linear_attn_bwd_work_bytes.argtypes = [L64]
