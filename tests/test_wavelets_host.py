"""Host-side logic of wdno_b200/wavelets.py that needs no GPU: filter banks against the oracle's, per-axis geometry against the
shapes the reference documents, and the view / stride analysis that decides whether a sub-band view can be handed to the
2-D kernels as is (image stride) or has to be made contiguous."""
import numpy as np
import pytest
import torch

from wdno_b200 import wavelets as W


def test_filter_banks_match_oracle():
    from oracle import wavelets as O
    for name in ("bior1.3", "bior2.4", "haar"):
        w = W.Wavelet(name)
        try:
            dl, dh, rl, rh = O.filter_bank(name)
        except Exception:
            pytest.skip(f"oracle has no {name}")
        for got, want in ((w.dec_lo, dl), (w.dec_hi, dh), (w.rec_lo, rl), (w.rec_hi, rh)):
            assert np.allclose(np.asarray(got), np.asarray(want), atol=1e-15), name
        assert w.dec_len == len(dl) and W.Wavelet(w).name == name
    with pytest.raises(ValueError):
        W.Wavelet("db4")


def test_geometry_matches_reference_documented_shapes():
    # (nout, off, periodic): wave_trans_2d.py:172,176,183 -> 32x64x64 gives 18x34x34 ('zero', 6 taps); data_burgers_1d.py:53 ->
    # 81x120 gives 41x60 ('periodization', 10 taps)
    assert [W._geom(n, 6, "zero")[0] for n in (32, 64, 16, 8)] == [18, 34, 10, 6]
    assert W._geom(64, 6, "zero")[1:] == (4, 0)
    assert [W._geom(n, 10, "periodization")[0] for n in (81, 120, 41, 21)] == [41, 60, 21, 11]
    assert W._geom(81, 10, "periodization")[1:] == (4, 1)
    with pytest.raises(ValueError):
        W._geom(8, 6, "symmetric")


def test_image_stride_of_sub_band_views():
    hw = 5 * 7
    assert W._img_stride(torch.zeros(3, 2, 5, 7)) == hw
    assert W._img_stride(torch.zeros(3, 2, 3, 5, 7).select(-3, 1)) == 3 * hw        # a detail band of DWTForward's Yh
    assert W._img_stride(torch.zeros(3, 2, 4, 5, 7)[:, :, 2]) == 4 * hw             # a slot of the packed builder layout
    assert W._img_stride(torch.zeros(3, 1, 3, 5, 7).select(-3, 0)) == 3 * hw        # size-1 channel dim is skipped
    assert W._img_stride(torch.zeros(1, 1, 5, 7)) == hw and W._img_stride(torch.zeros(5, 7)) == hw
    assert W._img_stride(torch.zeros(3, 2, 7, 5).transpose(-1, -2)) is None         # planes not contiguous
    assert W._img_stride(torch.zeros(2, 3, 5, 7).transpose(0, 1)) is None           # leading dims do not collapse uniformly
    assert W._img_stride(torch.zeros(4, 6, 5, 7)[:, :2]) is None
    assert W._img_stride(torch.zeros(3, 2, 5, 8)[..., :7]) is None                  # cropped rows (DWTInverse's ll[..., :-1])
    t, st = W._planes(torch.zeros(3, 2, 5, 8)[..., :7])
    assert t.is_contiguous() and st == hw


def test_cpu_tensors_are_rejected():
    with pytest.raises(RuntimeError):
        W.wavedec3(torch.zeros(1, 8, 8, 8), "bior1.3")
    with pytest.raises(RuntimeError):
        W.DWTForward(J=1, wave="bior2.4", mode="periodization")(torch.zeros(1, 1, 16, 16))
    with pytest.raises(RuntimeError):
        W.dwt2_packed(torch.zeros(1, 1, 16, 16), "bior2.4", "periodization")
