"""Run in a fresh interpreter by tests/test_dropin_cpu.py: install the drop-in, then execute the module bodies of the
reference's four scripts UNCHANGED (everything above `if __name__ == "__main__"`, so argparse does not run) and print
which module every hot-path name resolved to.  Third-party packages this image lacks (accelerate, ema_pytorch,
rotary_embedding_torch, einops_exts, h5py, matplotlib, IPython, tensorboardX) get the import shims of oracle/ref_loader.py;
the PhiFlow solver package `dataset.evaluate_solver` (TensorFlow 1) is stubbed.  Prints one JSON object."""
import json
import os
import sys
import types
import warnings

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = sys.argv[1]

from oracle import ref_loader  # noqa: E402

ref_loader._install_shims()
for name in ("pywt", "ptwt", "pytorch_wavelets"):   # ref_loader's dummies must not shadow the engine's stand-ins
    sys.modules.pop(name, None)
import wdno_b200  # noqa: E402

front = wdno_b200.install(REF)
pk = types.ModuleType("dataset")
pk.__path__ = []
ev = types.ModuleType("dataset.evaluate_solver")
ev.__all__ = []
sys.modules["dataset"], sys.modules["dataset.evaluate_solver"] = pk, ev


def body(rel, name):
    path = os.path.join(REF, rel)
    src = open(path).read().split("if __name__")[0]
    mod = types.ModuleType(name)
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def where(obj):
    return getattr(obj, "__module__", None)


out = {"front": front}
os.chdir(os.path.join(REF, "smoke"))
inf = body("smoke/inference_2d.py", "ref_inference_2d")
import ddpm.utils as U  # noqa: E402  (the reference's own file, found through the mirror package's __path__)

out["smoke_inference"] = {
    "utils_file": U.__file__,
    "GaussianDiffusion": where(U.GaussianDiffusion), "Unet3D_with_Conv3D": where(U.Unet3D_with_Conv3D),
    "Trainer": where(U.Trainer), "load_ddpm_base_model": where(inf.load_ddpm_base_model),
    "upsample_coef": where(inf.upsample_coef), "tensor_to_coef": where(inf.tensor_to_coef),
    "coef_to_tensor": where(inf.coef_to_tensor), "DWTForward": where(inf.DWTForward),
    "waverec3": where(inf.ptwt.waverec3), "Wavelet": where(inf.pywt.Wavelet), "Smoke_wave": where(inf.Smoke_wave),
    "InferencePipeline": where(inf.InferencePipeline)}
tr = body("smoke/train_2d.py", "ref_train_2d")
out["smoke_train"] = {"GaussianDiffusion": where(tr.GaussianDiffusion), "Unet3D_with_Conv3D": where(tr.Unet3D_with_Conv3D),
                      "Unet3D": where(tr.Unet3D), "Unet": where(tr.Unet), "Trainer": where(tr.Trainer)}
# the classes load_ddpm_base_model constructs (smoke/ddpm/utils.py:92-134) are the engine's, with the reference's arguments
import torch  # noqa: E402

m = U.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42)
gd = U.GaussianDiffusion(m, torch.ones(1, 1, 42, 1, 1), True, True, True, False, "bior1.3", "zero", [18, 34, 34],
                         [32, 64, 64], image_size=40, frames=24, timesteps=1000, sampling_timesteps=100,
                         ddim_sampling_eta=1.0, standard_fixed_ratio=0.01, coeff_ratio=0.1)
out["constructed"] = [type(m).__module__, type(gd).__module__, len(gd.state_dict())]

os.chdir(os.path.join(REF, "burgers"))   # the reference's own convention: test_util.py:4 appends ./ddpm_burgers/
evb = body("burgers/eval_ddpm_burgers.py", "ref_eval_burgers")
import ddpm_burgers.test_util as TU  # noqa: E402

out["burgers_eval"] = {"test_util_file": TU.__file__, "GaussianDiffusion": where(TU.GaussianDiffusion),
                       "Trainer": where(TU.Trainer), "get_wt_T": where(evb.get_wt_T),
                       "upsample_coef": where(evb.upsample_coef), "tensor_to_coef": where(evb.tensor_to_coef),
                       "DWTInverse": where(evb.DWTInverse), "load_2dconv_base_model": where(evb.load_2dconv_base_model)}
trb = body("burgers/train_ddpm_burgers.py", "ref_train_burgers")
out["burgers_train"] = {"Unet2D": where(trb.Unet2D), "GaussianDiffusion": where(trb.GaussianDiffusion),
                        "GaussianDiffusion1D": where(trb.GaussianDiffusion1D), "Trainer": where(trb.Trainer),
                        "get_wavelet_preprocess": where(trb.get_wavelet_preprocess)}
print("DROPIN_JSON " + json.dumps(out))
