"""The drop-in switch: `wdno_b200.install(reference_root)` makes the reference's own scripts
(smoke/inference_2d.py, smoke/train_2d.py, burgers/eval_ddpm_burgers.py, burgers/train_ddpm_burgers.py) import the
B200 engine for the hot path and the reference's files for everything else.

How: the mirror trees `wdno_b200/smoke/` and `wdno_b200/burgers/` go FIRST on sys.path, the reference's `smoke/` and
`burgers/` right behind them.  The mirror packages `ddpm`, `video_diffusion_pytorch`, `ddpm_burgers` append the
reference's directory of the same name to their `__path__`, so `ddpm.diffusion_2d`, `ddpm.wave_utils`,
`video_diffusion_pytorch.video_diffusion_pytorch_conv3d`, `ddpm_burgers.{diffusion_1d,unet,wavelet_utils}` and the top-level
`wave_trans_2d` / `wave_trans` resolve to the engine, while `ddpm.data_2d`, `ddpm.utils`, `ddpm.modules`,
`video_diffusion_pytorch.video_diffusion_pytorch`, `ddpm_burgers.{test_util,train_diffusion,model_utils,...}`,
`dataset.*` fall through to the reference unchanged.  Names the reference imports from a mirrored module but that are
outside the hot path (`Trainer`, `Unet`, ... from `ddpm.diffusion_2d`; smoke/ddpm/utils.py:11, smoke/train_2d.py:5) are
served lazily (PEP 562) from the reference's own file, loaded under a private module name.
"""
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_MIRRORED = ("ddpm", "video_diffusion_pytorch", "ddpm_burgers", "wave_trans_2d", "wave_trans")
_state = {"root": os.environ.get("WDNO_REFERENCE_ROOT")}


def reference_root():
    return _state["root"]


def extend_path(pkg_path, rel):
    """`__path__` of a mirror package + the reference's directory `rel` (if a reference tree is configured)."""
    root = _state["root"]
    out = list(pkg_path)
    if root:
        d = os.path.join(root, rel)
        if os.path.isdir(d) and d not in out:
            out.append(d)
    return out


def reference_attr(rel_file, name, private_name):
    """attribute `name` of the reference source file `rel_file`, loaded once under `private_name`."""
    mod = sys.modules.get(private_name)
    if mod is None:
        root = _state["root"]
        if not root:
            raise AttributeError(
                f"{name!r} lives in the reference's {rel_file}; call wdno_b200.install(<reference root>) "
                "(or set WDNO_REFERENCE_ROOT) to make it importable next to the engine classes")
        path = os.path.join(root, rel_file)
        spec = importlib.util.spec_from_file_location(private_name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[private_name] = mod
        try:
            spec.loader.exec_module(mod)
        except BaseException:
            sys.modules.pop(private_name, None)
            raise
    try:
        return getattr(mod, name)
    except AttributeError:
        raise AttributeError(f"neither the engine mirror nor the reference's {rel_file} defines {name!r}") from None


def install_wavelet_modules(force=True):
    """Register `pytorch_wavelets`, `ptwt`, `pywt` stand-ins backed by the DWT kernels (wdno_b200.wavelets).
    force=False keeps real installations of those packages if they are importable."""
    from . import wavelets as W

    def want(name):
        if force:
            return True
        try:
            return importlib.util.find_spec(name) is None
        except (ImportError, ValueError):
            return True

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__wdno_b200__ = True
        sys.modules[name] = m

    if want("pytorch_wavelets"):
        mod("pytorch_wavelets", DWTForward=W.DWTForward, DWTInverse=W.DWTInverse, DWT1DForward=W.DWT1DForward,
            DWT1DInverse=W.DWT1DInverse)
    if want("ptwt"):
        mod("ptwt", wavedec3=W.wavedec3, waverec3=W.waverec3)
    if want("pywt"):
        mod("pywt", Wavelet=W.Wavelet)


def install(reference_root=None, *, smoke=True, burgers=True, wavelets=True):
    """Route the reference scripts' hot-path imports to the engine.  Call before importing any reference module.

    reference_root: checkout of AI4Science-WestlakeU/wdno (default: $WDNO_REFERENCE_ROOT); without one only the
    engine classes are importable under the reference's module names.
    wavelets: True -> always use the DWT kernels for `pytorch_wavelets` / `ptwt` / `pywt`; "missing" -> only where
    those packages are not installed; False -> leave them alone."""
    if reference_root is not None:
        _state["root"] = os.path.abspath(reference_root)
        os.environ["WDNO_REFERENCE_ROOT"] = _state["root"]
    root = _state["root"]
    # forget modules of the same names that were imported from elsewhere (e.g. the reference's own ddpm package)
    for name in list(sys.modules):
        top = name.split(".", 1)[0]
        if top in _MIRRORED:
            f = getattr(sys.modules[name], "__file__", None) or ""
            if not os.path.abspath(f).startswith(_HERE):
                del sys.modules[name]
    front = []
    for on, sub in ((smoke, "smoke"), (burgers, "burgers")):
        if not on:
            continue
        front.append(os.path.join(_HERE, sub))
        if root and os.path.isdir(os.path.join(root, sub)):
            front.append(os.path.join(root, sub))
    for p in front:
        while p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = front
    # already-imported mirror packages: refresh their __path__ now that a root is known
    for pkg, rel in (("ddpm", "smoke/ddpm"), ("video_diffusion_pytorch", "smoke/video_diffusion_pytorch"),
                     ("ddpm_burgers", "burgers/ddpm_burgers")):
        m = sys.modules.get(pkg)
        if m is not None and hasattr(m, "__path__"):
            m.__path__ = extend_path([p for p in m.__path__ if os.path.abspath(p).startswith(_HERE)], rel)
    if wavelets:
        install_wavelet_modules(force=(wavelets is True))
    return front
