"""Offline coefficient builders against tests/golden/coef_builders.pt, which holds summaries of the
files the reference's OWN `__main__` blocks (smoke/wave_trans_2d.py:61-189, burgers/wave_trans.py:66-127) wrote for seeded
synthetic raw data (tests/golden/make_golden.py builders).

CPU: the host logic (down-sampling strides, packing, per-simulation records, file names, key names, types) with the torch
restatement of the wavelets plugged in.  GPU: the same jobs end to end on the DWT kernels, plus the packed-layout launches
bit-compared with transform-then-pack and checked against the float64 oracle at the full 32x64x64 size."""
import os
import types

import numpy as np
import pytest
import torch

from tests.golden.make_golden import (BUILDER_STRIDE, builder_inputs_burgers, builder_inputs_smoke, sub,
                                      write_smoke_sims)

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-5  # relative l2 per tensor: fp32 kernels vs the fp32 conv-form restatement the golden run used


def golden():
    return torch.load(os.path.join(HERE, "golden", "coef_builders.pt"), weights_only=False)


def check_record(rec, gold, tol=TOL):
    """a saved dictionary against its golden summary: same keys, list lengths, tensor shapes, container types, values"""
    assert set(rec.keys()) == {k for k in gold if not k.endswith("_type")}
    for k, g in gold.items():
        if k.endswith("_type"):
            continue
        v = rec[k]
        if isinstance(g, list) and g and isinstance(g[0], dict):
            assert isinstance(v, list) and len(v) == len(g)
            for t, gt in zip(v, g):
                assert tuple(t.shape) == gt["shape"] and t.dtype == torch.float32 and t.device.type == "cpu"
                s = sub(t, BUILDER_STRIDE)
                assert float((s - gt["sub"]).norm() / (gt["sub"].norm() + 1e-30)) < tol, k
                assert abs(float(t.double().norm()) - gt["norm"]) <= tol * gt["norm"], k
        elif isinstance(g, list):
            assert [tuple(x) for x in v] == g and type(v[0]).__name__ == gold[k + "_type"], k
        else:
            assert tuple(v) == g and type(v).__name__ == gold[k + "_type"], k


def oracle_namespace():
    """the three entry points coef_builders calls, on the CPU restatement (oracle/wavelets_torch.py)"""
    from oracle import wavelets_torch as wt

    def wavedec3_packed(x, wave, mode="zero"):
        yl, yh = wt.wavedec3(x, wave, mode=mode)
        return torch.cat((yl[:, None], torch.stack(list(yh.values()), dim=1)), dim=1)

    def dwt2_packed(x, wave, mode):
        yl, yh = wt.DWTForward(J=1, wave=wave, mode=mode)(x)
        return torch.cat((yl[:, :, None], yh[0]), dim=2)

    return types.SimpleNamespace(wavedec3_packed=wavedec3_packed, dwt2_packed=dwt2_packed, afb1d=wt.afb1d)


def run_builders(tmp, device):
    from wdno_b200 import coef_builders as CB
    sims = builder_inputs_smoke()
    write_smoke_sims(str(tmp), sims)
    root = os.path.join(str(tmp), "data", "2d") + "/"
    mx = CB.build_smoke_coef_files(root, "train/", range(len(sims)), batch_sims=2, device=device)
    mx1 = CB.build_smoke_coef_files(root, "train/", range(1), batch_sims=1, device=device)  # ragged last batch of one
    assert len(mx) == 42 and all(isinstance(v, int) for v in mx) and all(a >= b for a, b in zip(mx, mx1))
    g = golden()
    wave_dir = os.path.join(root, "train", "bior1.3_zero")
    for kind in ("time", "space"):
        assert sorted(os.listdir(os.path.join(wave_dir, kind + "_downsample"))) == ["000000", "000001"]
        for i in range(len(sims)):
            rec = torch.load(os.path.join(wave_dir, kind + "_downsample", "{:06d}".format(i)), weights_only=False)
            check_record(rec, g["smoke"][kind][i])
    os.makedirs(os.path.join(str(tmp), "data", "1d"), exist_ok=True)
    train = os.path.join(str(tmp), "data", "1d", "train")
    torch.save(builder_inputs_burgers(), train)
    out = CB.build_burgers_coef_file(train, batch_size=2, device=device)  # 3 trajectories in batches of 2 + 1
    assert out == os.path.join(str(tmp), "data", "1d", "coef_bior2.4_periodization_super")
    check_record(torch.load(out, weights_only=False), g["burgers"])


def test_builders_host_logic_reproduces_reference_files_cpu(tmp_path, monkeypatch):
    from wdno_b200 import coef_builders as CB
    monkeypatch.setattr(CB, "W", oracle_namespace())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)  # no CUDA runtime in the build container
    run_builders(tmp_path, "cpu")


def test_mirror_modules_export_builders():
    import wdno_b200.burgers.wave_trans as bw
    import wdno_b200.smoke.wave_trans_2d as sw
    assert callable(sw.build_smoke_coef_files) and callable(sw.smoke_sims_to_coef)
    assert callable(bw.build_burgers_coef_file) and callable(bw.burgers_data_to_coef)


def test_oracle_dwt_max_level():
    from oracle import wavelets_torch as wt
    # pywt.dwt_max_level(80, 'bior2.4') = 3 and (32, 'bior1.3') = 2: the values the reference scripts print / loop over
    assert wt.dwt_max_level(80, "bior2.4") == 3 and wt.dwt_max_level(32, "bior1.3") == 2 and wt.dwt_max_level(4, 10) == 0


@pytest.mark.gpu
def test_builders_reproduce_reference_files_gpu(tmp_path):
    run_builders(tmp_path, "cuda")


@pytest.mark.gpu
def test_packed_launches_bit_equal_to_transform_then_pack_and_oracle():
    from oracle import wavelets as O
    from wdno_b200 import packing as P
    from wdno_b200 import wavelets as W
    rng = np.random.default_rng(5)
    for shape in ((10, 32, 64, 64), (5, 8, 64, 64), (5, 32, 16, 16), (3, 7, 9, 11)):
        x = rng.standard_normal(shape)
        xc = torch.tensor(x, dtype=torch.float32, device="cuda")
        packed = W.wavedec3_packed(xc, "bior1.3")
        ref = P.smoke_coef_to_tensor(W.wavedec3(xc, "bior1.3"))
        assert packed.is_contiguous() and torch.equal(packed, ref), shape
        aaa_o, d_o = O.wavedec3(x, "bior1.3")
        want = np.concatenate([aaa_o[:, None]] + [d_o[k][:, None] for k in O.KEYS3], axis=1)
        assert float((packed.cpu().double() - torch.from_numpy(want)).abs().max()) < 2e-6 * np.abs(want).max()
    for shape, wave, mode in (((7, 2, 81, 120), "bior2.4", "periodization"), ((4, 2, 41, 60), "bior2.4", "periodization"),
                              ((5, 1, 64, 64), "bior1.3", "zero"), ((2, 2, 11, 15), "bior2.4", "periodization")):
        xc = torch.tensor(rng.standard_normal(shape), dtype=torch.float32, device="cuda")
        packed = W.dwt2_packed(xc, wave, mode)
        yl, yh = W.DWTForward(J=1, wave=wave, mode=mode)(xc)
        assert torch.equal(packed, P.burgers_coef_to_tensor(yl, yh)), shape


@pytest.mark.gpu
def test_smoke_builder_full_size_roundtrip_and_batch_independence():
    """BASELINE-size fields (5 x 32 x 64 x 64 per simulation): level-0 coefficients reconstruct the raw fields, and a
    simulation's coefficients do not depend on which batch it was transformed in"""
    from wdno_b200 import coef_builders as CB
    from wdno_b200 import packing as P
    from wdno_b200 import wavelets as W
    g = torch.Generator().manual_seed(3)
    X = torch.randn(6, 5, 32, 64, 64, generator=g).cuda()
    s = (torch.rand(6, 32, generator=g) + 0.1).cuda()
    res = CB.smoke_sims_to_coef(X, s)
    assert [tuple(c.shape[2:]) for c in res["time"]["coef"]] == [(8, 18, 34, 34), (8, 10, 34, 34), (8, 6, 34, 34)]
    assert [tuple(c.shape[2:]) for c in res["space"]["coef"]] == [(8, 18, 34, 34), (8, 18, 18, 18), (8, 18, 10, 10)]
    assert [tuple(c.shape[2:]) for c in res["time"]["init_coef"]] == [(4, 34, 34)] * 3
    assert [tuple(c.shape[1:]) for c in res["time"]["smokeout"]] == [(2, 18), (2, 10), (2, 6)]
    assert [tuple(c.shape[1:]) for c in res["space"]["smokeout"]] == [(2, 18)] * 3
    c0 = res["time"]["coef"][0]
    rec = W.waverec3(P.smoke_tensor_to_coef(c0.reshape(6, 40, 18, 34, 34), [18, 34, 34]), "bior1.3")
    assert float((rec.view_as(X) - X).abs().max()) < 1e-5
    one = CB.smoke_sims_to_coef(X[4:5], s[4:5])
    for kind in CB.SMOKE_KINDS:
        for name in ("coef", "init_coef", "smokeout"):
            for a, b in zip(res[kind][name], one[kind][name]):
                assert torch.equal(a[4:5], b), (kind, name)


@pytest.mark.skipif(not __import__("oracle.ref_loader", fromlist=["available"]).available(), reason="/root/reference not mounted")
def test_reference_dataset_classes_read_our_files_like_their_own(tmp_path, monkeypatch):
    """the consumers of the on-disk format -- smoke `Smoke_wave.__getitem__` (smoke/ddpm/data_2d.py:156-221) and the Burgers
    `get_wavelet_super_preprocess` (burgers/ddpm_burgers/data_burgers_1d.py:20-85), imported from the reference unchanged --
    produce the same training samples from the files our builders write as from the files the reference's scripts write
    (full-size fields, base and super-resolution variants)"""
    import importlib
    from oracle import ref_loader
    from wdno_b200 import coef_builders as CB
    monkeypatch.setattr(CB, "W", oracle_namespace())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    ref_dir, our_dir = tmp_path / "ref", tmp_path / "ours"
    sims = builder_inputs_smoke(n_sims=1, T=32, H=64, seed=91)
    for d in (ref_dir, our_dir):
        write_smoke_sims(str(d), sims)
    err = ref_loader.run_reference_main("smoke/wave_trans_2d.py", str(ref_dir))
    assert isinstance(err, FileNotFoundError)   # the script's hard-coded 20 000 ids end at the first missing simulation
    CB.build_smoke_coef_files(os.path.join(str(our_dir), "data", "2d") + "/", "train/", range(1), batch_sims=1, device="cpu")
    ref_loader.smoke()
    data_2d = importlib.import_module("ddpm.data_2d")
    real_load = torch.load
    monkeypatch.setattr(torch, "load", lambda *a, **k: real_load(*a, **{**k, "weights_only": False}))  # files hold torch.Size
    for sup, kind, n in ((False, "time", 0), (False, "space", 0), (True, "time", 0), (True, "space", 0), (True, "space", 1)):
        items = []
        for d in (ref_dir, our_dir):
            ds = data_2d.Smoke_wave(os.path.join(str(d), "data", "2d"), "bior1.3", "zero", is_super_model=sup,
                                    downsample_type=kind, N_downsample=n)
            items.append(ds[0])
        (sa, sha, oa, ida), (sb, shb, ob, idb) = items
        assert sa.shape == sb.shape and sha == shb and oa == ob and ida == idb, (sup, kind, n)
        assert float((sa - sb).abs().max()) <= 1e-5 * float(sa.abs().max()), (sup, kind, n)
    # Burgers
    for d in (ref_dir, our_dir):
        os.makedirs(os.path.join(str(d), "data", "1d"), exist_ok=True)
        torch.save(builder_inputs_burgers(N=3, seed=92), os.path.join(str(d), "data", "1d", "train"))
    assert ref_loader.run_reference_main("burgers/wave_trans.py", str(ref_dir)) is None
    CB.build_burgers_coef_file(os.path.join(str(our_dir), "data", "1d", "train"), device="cpu")
    ref_loader.burgers()
    ref_loader.install_wavelet_shims()
    data_1d = importlib.import_module("ddpm_burgers.data_burgers_1d")
    dbs = [real_load(os.path.join(str(d), "data", "1d", "coef_bior2.4_periodization_super"), weights_only=False)
           for d in (ref_dir, our_dir)]
    for sup, n in ((False, 0), (True, 0), (True, 1)):
        outs = [data_1d.get_wavelet_super_preprocess(rescaler=70, is_super_model=sup, N_downsample=n, mode="periodization",
                                                     wave_type="bior2.4", is_condition_u0=True, is_condition_uT=True)(
                    {k: ([t.clone() for t in v] if k == "coef" else v) for k, v in db.items()}) for db in dbs]
        (da, sha, oa), (db_, shb, ob) = outs
        assert da.shape == db_.shape and sha == shb and oa == ob, (sup, n)
        assert float((da - db_).abs().max()) <= 1e-5 * float(da.abs().max()), (sup, n)
