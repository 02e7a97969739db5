"""TEST INFRASTRUCTURE ONLY -- import the reference's own modules from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  It is used by
tests/golden/make_golden.py to generate the committed fixtures and by the `-m "not gpu"` tests to
pin the oracle restatement (oracle/unet3d.py, oracle/unet2d.py, oracle/diffusion.py) against the
real reference code.  The third-party packages the reference imports but this image lacks are
replaced by the minimal shims below; none of the shims is on the
numeric path except RotaryEmbedding and rearrange_many, which restate the published algorithms of
rotary-embedding-torch / einops-exts.
"""
import importlib
import importlib.util
import math
import os
import sys
import types

import torch
from torch import nn

REF_ROOT = os.environ.get("WDNO_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "smoke")) and os.path.isdir(os.path.join(REF_ROOT, "burgers"))


# ------------------------------------------------------------------ shims
class RotaryEmbedding(nn.Module):
    """rotary_embedding_torch.RotaryEmbedding(dim): freqs = theta^-(0,2,..,dim-2)/dim as a parameter named
    `freqs`; rotate_queries_or_keys(t) rotates interleaved pairs of the last dim by position * freq."""

    def __init__(self, dim, theta=10000):
        super().__init__()
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
        self.freqs = nn.Parameter(freqs, requires_grad=False)

    def rotate_queries_or_keys(self, t, seq_dim=-2):
        n = t.shape[seq_dim]
        pos = torch.arange(n, device=t.device, dtype=self.freqs.dtype)
        ang = torch.einsum("i,j->ij", pos, self.freqs)
        ang = ang.repeat_interleave(2, dim=-1)  # (n, dim): f0 f0 f1 f1 ...
        x = t.reshape(*t.shape[:-1], -1, 2)
        x1, x2 = x.unbind(dim=-1)
        rot = torch.stack((-x2, x1), dim=-1).reshape(t.shape)
        return t * ang.cos() + rot * ang.sin()


def _install_shims():
    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    from einops import rearrange

    def rearrange_many(tensors, pattern, **kw):
        return tuple(rearrange(t, pattern, **kw) for t in tensors)

    def check_shape(tensor, pattern, **kw):
        return rearrange(tensor, f"{pattern} -> {pattern}", **kw)

    class _Dummy:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return _Dummy()

        def __call__(self, *a, **k):
            return _Dummy()

    def try_real(name):
        try:
            importlib.import_module(name)
            return True
        except Exception:
            return False

    if not try_real("rotary_embedding_torch"):
        mod("rotary_embedding_torch", RotaryEmbedding=RotaryEmbedding)
    if not try_real("einops_exts"):
        mod("einops_exts", rearrange_many=rearrange_many, check_shape=check_shape)
    if not try_real("ema_pytorch"):
        mod("ema_pytorch", EMA=_Dummy)
    if not try_real("accelerate"):
        mod("accelerate", Accelerator=_Dummy, DistributedDataParallelKwargs=_Dummy)
    if not try_real("tensorboardX"):
        mod("tensorboardX", SummaryWriter=_Dummy)
    if not try_real("IPython"):
        mod("IPython", embed=lambda *a, **k: None)
    if not try_real("h5py"):
        mod("h5py", File=_Dummy)
    if not try_real("termcolor"):
        mod("termcolor", colored=lambda s, *a, **k: s)
    if not try_real("imageio"):
        mod("imageio")
    if not try_real("matplotlib"):
        m = mod("matplotlib", use=lambda *a, **k: None)
        mod("matplotlib.pyplot")
        m.pyplot = sys.modules["matplotlib.pyplot"]
        mod("matplotlib.pylab")
        mod("matplotlib.backends")
        mod("matplotlib.backends.backend_pdf", PdfPages=_Dummy)
    if not try_real("pywt"):
        mod("pywt", Wavelet=_Dummy, wavedec=_Dummy, waverec=_Dummy, dwt_max_level=lambda *a, **k: 1)
    if not try_real("pytorch_wavelets"):
        mod("pytorch_wavelets", DWTForward=_Dummy, DWTInverse=_Dummy, DWT1DForward=_Dummy, DWT1DInverse=_Dummy)
    if not try_real("ptwt"):
        mod("ptwt", wavedec3=_Dummy, waverec3=_Dummy)
    if not try_real("multiprocess"):
        mod("multiprocess", Process=_Dummy)


def install_wavelet_shims():
    """Give the absent wavelet packages WORKING, differentiable stand-ins (oracle/wavelets_torch.py) so the reference's
    own glue code (inference_2d.py guidance_fn / InferencePipeline, eval_ddpm_burgers.py helpers) runs unchanged."""
    from . import wavelets_torch as wt
    _install_shims()
    for name, attrs in (("pywt", dict(Wavelet=wt.Wavelet, dwt_max_level=wt.dwt_max_level)),
                        ("ptwt", dict(wavedec3=wt.wavedec3, waverec3=wt.waverec3)),
                        ("pytorch_wavelets", dict(DWTForward=wt.DWTForward, DWTInverse=wt.DWTInverse,
                                                  DWT1DForward=wt.DWT1DForward, DWT1DInverse=wt.DWT1DInverse))):
        m = sys.modules[name]
        if getattr(m, "__file__", None) is None:  # only our shim modules, never a real installation
            for k, v in attrs.items():
                setattr(m, k, v)


def _import_from(path, name, patch=None):
    """exec a reference source file as module `name` (optionally patching its text in memory)."""
    with open(path, "r") as f:
        src = f.read()
    if patch is not None:
        src = patch(src)
    m = types.ModuleType(name)
    m.__file__ = path
    sys.modules[name] = m
    exec(compile(src, path, "exec"), m.__dict__)
    return m


_cache = {}


def smoke():
    """-> namespace with Unet3D_with_Conv3D and GaussianDiffusion of the smoke experiment."""
    if "smoke" in _cache:
        return _cache["smoke"]
    assert available(), "reference tree not mounted"
    _install_shims()
    sroot = os.path.join(REF_ROOT, "smoke")
    if sroot not in sys.path:
        sys.path.insert(0, sroot)
    conv3d = importlib.import_module("video_diffusion_pytorch.video_diffusion_pytorch_conv3d")
    diff = importlib.import_module("ddpm.diffusion_2d")
    wave_utils = importlib.import_module("ddpm.wave_utils")
    ns = types.SimpleNamespace(Unet3D_with_Conv3D=conv3d.Unet3D_with_Conv3D, GaussianDiffusion=diff.GaussianDiffusion,
                               conv3d=conv3d, diffusion_2d=diff, wave_utils=wave_utils)
    try:
        ns.wave_trans_2d = _import_from(os.path.join(sroot, "wave_trans_2d.py"), "ref_wave_trans_2d",
                                        patch=lambda s: s.split("if __name__")[0])
    except Exception as e:  # pragma: no cover - informational
        ns.wave_trans_2d = None
        ns.wave_trans_2d_error = repr(e)
    _cache["smoke"] = ns
    return ns


def smoke_inference():
    """-> the reference's smoke/inference_2d.py (guidance_fn, InferencePipeline) as a module.  Its imports of the dataset
    classes and of the PhiFlow solver (`ddpm.data_2d`, `dataset.evaluate_solver`: never used by guidance_fn /
    run_model) are dropped from the text in memory; everything else is the file as it is."""
    if "smoke_inference" in _cache:
        return _cache["smoke_inference"]
    smoke()
    install_wavelet_shims()
    sroot = os.path.join(REF_ROOT, "smoke")

    def patch(src):
        src = src.split("if __name__")[0]
        src = src.replace("from ddpm.data_2d import Smoke, Smoke_wave", "")
        return src.replace("from dataset.evaluate_solver import *", "")
    m = _import_from(os.path.join(sroot, "inference_2d.py"), "ref_inference_2d", patch=patch)
    _cache["smoke_inference"] = m
    return m


def burgers():
    """-> namespace with Unet2D and GaussianDiffusion(1D) of the Burgers experiment.
    diffusion_1d.py:51 has a `device='cuda'` default argument that cannot be evaluated without a GPU: the
    text is patched in memory to device='cpu' (no other change)."""
    if "burgers" in _cache:
        return _cache["burgers"]
    assert available(), "reference tree not mounted"
    _install_shims()
    broot = os.path.join(REF_ROOT, "burgers")
    if broot not in sys.path:
        sys.path.insert(0, broot)
    # the smoke tree also has a top-level `ddpm`; burgers uses `ddpm_burgers` so there is no clash
    unet = importlib.import_module("ddpm_burgers.unet")
    model_utils = importlib.import_module("ddpm_burgers.model_utils")
    wavelet_utils = importlib.import_module("ddpm_burgers.wavelet_utils")
    wt = _import_from(os.path.join(broot, "wave_trans.py"), "wave_trans", patch=lambda s: s.split("if __name__")[0])
    d1 = _import_from(os.path.join(broot, "ddpm_burgers", "diffusion_1d.py"), "ddpm_burgers.diffusion_1d",
                      patch=lambda s: s.replace("device='cuda')", "device='cpu')") if not torch.cuda.is_available() else s)
    ns = types.SimpleNamespace(Unet2D=unet.Unet2D, Unet1D=unet.Unet1D, GaussianDiffusion=d1.GaussianDiffusion,
                               GaussianDiffusion1D=d1.GaussianDiffusion1D, unet=unet, diffusion_1d=d1,
                               model_utils=model_utils, wavelet_utils=wavelet_utils, wave_trans=wt)
    _cache["burgers"] = ns
    return ns


def run_reference_main(rel_path, cwd):
    """Execute a reference script's `__main__` block unchanged (runpy, run_name='__main__') with `cwd` as the working
    directory -- used for the offline coefficient builders (smoke/wave_trans_2d.py:61-189, burgers/wave_trans.py:66-127),
    whose inputs and outputs are relative paths.  The wavelet stand-ins must be installed first (install_wavelet_shims).
    Returns the exception that ended the script, if any (the smoke builder loops over 20 000 hard-coded simulation ids and
    stops at the first missing directory, after having saved the files of the simulations that exist)."""
    import runpy
    assert available(), "reference tree not mounted"
    install_wavelet_shims()
    path = os.path.join(REF_ROOT, rel_path)
    sdir = os.path.dirname(path)
    old, added = os.getcwd(), False
    if sdir not in sys.path:
        sys.path.insert(0, sdir)
        added = True
    os.chdir(cwd)
    try:
        runpy.run_path(path, run_name="__main__")
        return None
    except (FileNotFoundError, OSError) as e:
        return e
    finally:
        os.chdir(old)
        if added:
            sys.path.remove(sdir)
