"""The C-ABI library loads without a GPU and exports every symbol include/wdno_b200.h declares; argument
validation errors surface as ValueError with the library's message (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from wdno_b200 import build
    build.build()
    from wdno_b200 import _lib
    return _lib.lib()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "wdno_b200.h")).read()
    names = sorted(set(re.findall(r"\b(wdno_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    from wdno_b200 import _abi
    declared = {n for n in set(names) - {"wdno_last_error", "wdno_version", "wdno_device_cc", "wdno_tapgemm", "wdno_tapgemm_max_cluster_ctas", "wdno_wgrad", "wdno_wgrad_tc"} if not n.endswith("_bytes")}   # struct-taking entry points are bound next to their ctypes structs
    assert declared == set(_abi.SIGNATURES), declared ^ set(_abi.SIGNATURES)


def test_struct_layouts_match_header():
    from wdno_b200 import _abi, _lib
    assert C.sizeof(_lib.Tap) == 8 and C.sizeof(_lib.KSet) == 24 and C.sizeof(_lib.NChunk) == 40
    assert C.sizeof(_lib.TapGemmParams) == 256   # + strips, Wfull, fold, cluster
    assert C.sizeof(_abi.CondOp) == 96
    from wdno_b200 import training
    assert C.sizeof(training.WgradGroup) == 56          # 5 + 3 int32, 3 int64
    assert C.sizeof(training.WgradTcJob) == 184 and C.sizeof(training.WgradTcParams) == 128 and training.WgradTcParams.jobs.offset == 104
    assert C.sizeof(training.WgradParams) == 128 and training.WgradParams.groups.offset == 112   # == the C struct (g++ sizeof / offsetof)


def test_invalid_arguments_are_rejected_without_a_gpu(lib):
    from wdno_b200 import _lib
    p = _lib.TapGemmParams()
    rc = lib.wdno_tapgemm(C.byref(p), None)
    assert rc == -1 and b"KC" in lib.wdno_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "tapgemm")
    assert lib.wdno_chan_layernorm(None, None, None, None, 0, 64, 1e-5, None) == -1
    assert lib.wdno_dwt_analysis_axis(None, None, None, 1, 1, 1, 1, 1, 1, 1, None, None, 6, 4, 0, None) == -1


def test_product_path_has_no_cpu_fallback():
    import torch
    from wdno_b200.unet3d import Unet3D_with_Conv3D
    from wdno_b200 import wavelets
    m = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 24, 42, 40, 40), torch.zeros(1))
    with pytest.raises(RuntimeError):
        wavelets.wavedec3(torch.zeros(1, 8, 8, 8), "bior1.3")


def test_fused_wavelet_envelopes_are_host_decisions(lib):
    """wdno_dwt3d_supported / wdno_dwt2d_supported are pure host code: the shapes of the reference's configurations are inside
    the fused kernels' envelopes, degenerate ones fall back to the per-axis entry points"""
    # 3-D: (taps, coefficient width, signal width): smoke base 34 / 64, super-resolution 66 / 128 (bior1.3, 6 taps)
    assert lib.wdno_dwt3d_supported(6, 34, 64) == 1 and lib.wdno_dwt3d_supported(6, 66, 128) == 1
    assert lib.wdno_dwt3d_supported(5, 34, 64) == 0 and lib.wdno_dwt3d_supported(6, 0, 64) == 0
    # 2-D: (taps, H, W, nh, nw, periodic): Burgers 81 x 120 -> 41 x 60 and its inverse 82 x 120, cascade levels, smoke rho0
    assert lib.wdno_dwt2d_supported(10, 81, 120, 41, 60, 1) == 1 and lib.wdno_dwt2d_supported(10, 82, 120, 41, 60, 1) == 1
    assert lib.wdno_dwt2d_supported(10, 162, 240, 81, 120, 1) == 1 and lib.wdno_dwt2d_supported(6, 64, 64, 34, 34, 0) == 1
    assert lib.wdno_dwt2d_supported(10, 11, 15, 6, 8, 1) == 0      # fewer coefficients than taps under 'periodization'
    assert lib.wdno_dwt2d_supported(4, 64, 64, 32, 32, 1) == 0     # tap counts other than 2 / 6 / 10
    assert lib.wdno_dwt2d_supported(6, 2000, 20000, 1002, 10002, 0) == 0   # a single output row does not fit shared memory
