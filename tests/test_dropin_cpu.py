"""the drop-in boundary.  `wdno_b200.install(reference_root)` must let the reference's OWN scripts
import unchanged -- smoke/inference_2d.py:19-23, smoke/ddpm/utils.py:10-11, smoke/train_2d.py:5-8,
burgers/eval_ddpm_burgers.py:7-14, burgers/train_ddpm_burgers.py:7-10 -- with the hot-path classes resolving to the engine
and everything else to the reference's files (round-1 verdict: the plain sys.path switch shadowed `ddpm.data_2d`,
`ddpm.utils`, `Trainer`)."""
import json
import os
import subprocess
import sys

import pytest

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def probe():
    r = subprocess.run([sys.executable, os.path.join(HERE, "helpers", "dropin_probe.py"), ref_loader.REF_ROOT],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("DROPIN_JSON ")][-1]
    return json.loads(line[len("DROPIN_JSON "):])


def test_smoke_inference_script_imports_over_the_engine(probe):
    s = probe["smoke_inference"]
    assert s["utils_file"].startswith(ref_loader.REF_ROOT)                  # ddpm.utils is the reference's own file
    assert s["GaussianDiffusion"] == "wdno_b200.diffusion_smoke"
    assert s["Unet3D_with_Conv3D"] == "wdno_b200.unet3d"
    assert s["Trainer"].startswith("_wdno_reference.")                      # served from the reference's diffusion_2d.py
    assert s["load_ddpm_base_model"] == "ddpm.utils"
    assert s["upsample_coef"] == s["tensor_to_coef"] == s["coef_to_tensor"] == "wdno_b200.packing"
    assert s["DWTForward"] == s["waverec3"] == s["Wavelet"] == "wdno_b200.wavelets"
    assert s["Smoke_wave"] == "ddpm.data_2d" and s["InferencePipeline"] == "ref_inference_2d"
    assert probe["constructed"][:2] == ["wdno_b200.unet3d", "wdno_b200.diffusion_smoke"]


def test_smoke_train_script_imports_over_the_engine(probe):
    s = probe["smoke_train"]
    assert s["GaussianDiffusion"] == "wdno_b200.diffusion_smoke" and s["Unet3D_with_Conv3D"] == "wdno_b200.unet3d"
    assert s["Unet3D"] == "video_diffusion_pytorch.video_diffusion_pytorch"  # the unused lucidrains class: reference file
    assert s["Unet"].startswith("_wdno_reference.") and s["Trainer"].startswith("_wdno_reference.")


def test_burgers_scripts_import_over_the_engine(probe):
    e, t = probe["burgers_eval"], probe["burgers_train"]
    assert e["test_util_file"].startswith(ref_loader.REF_ROOT)
    assert e["GaussianDiffusion"] == "wdno_b200.diffusion_burgers" and e["Trainer"] == "ddpm_burgers.train_diffusion"
    assert e["get_wt_T"] == e["upsample_coef"] == e["tensor_to_coef"] == "wdno_b200.packing"
    assert e["DWTInverse"] == "wdno_b200.wavelets" and e["load_2dconv_base_model"] == "ddpm_burgers.test_util"
    assert t["Unet2D"] == "wdno_b200.unet2d"
    assert t["GaussianDiffusion"] == t["GaussianDiffusion1D"] == "wdno_b200.diffusion_burgers"
    assert t["Trainer"] == "ddpm_burgers.train_diffusion" and t["get_wavelet_preprocess"] == "ddpm_burgers.data_burgers_1d"


def test_mirror_without_a_reference_tree_still_serves_the_engine_classes():
    code = ("import sys, os; os.environ.pop('WDNO_REFERENCE_ROOT', None); sys.path.insert(0, %r); import wdno_b200; "
            "wdno_b200.install(); from ddpm.diffusion_2d import GaussianDiffusion as G; "
            "from ddpm_burgers.unet import Unet2D as U; print(G.__module__, U.__module__)\n"
            "try:\n    from ddpm.diffusion_2d import Trainer\nexcept (ImportError, AttributeError) as e:\n    print('NOTRAINER')\n") % os.path.dirname(HERE)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "wdno_b200.diffusion_smoke wdno_b200.unet2d" in r.stdout and "NOTRAINER" in r.stdout
