"""Generate the committed golden vectors by running the REAL reference modules (imported from /root/reference).

    python tests/golden/make_golden.py        # build container only

Weights are NOT stored (95 MB): the reference module built under torch.manual_seed(0) is bit-reproduced by
wdno_b200's own module under the same seed (tests/test_oracle_vs_reference.py::test_engine_state_dict_keys_match_reference),
and each fixture records a checksum of the weights/inputs it was generated with so a drifted RNG is detected, not
silently compared.  Outputs are stored as strided sub-samples (stride 7) plus norms.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import ref_loader  # noqa: E402
from tests.test_oracle_vs_reference import NoiseTape, patched_randn  # noqa: E402

STRIDE = 7


def checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def sub(t):
    return t.reshape(-1)[::STRIDE].clone()


def smoke_fixture():
    s = ref_loader.smoke()
    torch.manual_seed(0)
    m = s.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42).eval()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 24, 42, 40, 40, generator=g)
    t = torch.tensor([321])
    with torch.no_grad():
        y = m(x, t)
    out = dict(weights_checksum=checksum(m.state_dict()), x_checksum=float(x.double().abs().sum()), t=t, y_sub=sub(y),
               y_norm=float(y.norm()), stride=STRIDE)
    torch.save(out, os.path.join(HERE, "smoke_unet3d_fwd.pt"))
    print("smoke_unet3d_fwd", out["weights_checksum"], out["y_norm"])
    # DDIM, 4 steps, eta = 1, base simulation config (C3 shape), injected noise
    gd = s.GaussianDiffusion(m, torch.ones(1), True, True, True, False, "bior1.3", "zero", [18, 34, 34], [32, 64, 64],
                             image_size=40, frames=24, timesteps=1000, sampling_timesteps=4, ddim_sampling_eta=1.0)
    init = torch.randn(1, 24, 40, 40, generator=g)
    control = torch.randn(1, 24, 16, 40, 40, generator=g)
    with patched_randn(NoiseTape(11)), torch.no_grad():
        smp = gd.sample(batch_size=1, init=init, control=control)
    out = dict(weights_checksum=checksum(m.state_dict()), init_checksum=float(init.double().abs().sum()),
               sample_sub=sub(smp), sample_norm=float(smp.norm()), stride=STRIDE, steps=4, eta=1.0, tape_seed=11)
    torch.save(out, os.path.join(HERE, "smoke_ddim4.pt"))
    print("smoke_ddim4", out["sample_norm"])


def burgers_fixture():
    b = ref_loader.burgers()
    torch.manual_seed(0)
    m = b.Unet2D(dim=128, dim_mults=[1, 2, 4, 8], channels=9, out_dim=9, resnet_block_groups=1).eval()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 9, 64, 64, generator=g)
    t = torch.tensor([17, 803])
    with torch.no_grad():
        y = m(x, t)
    out = dict(weights_checksum=checksum(m.state_dict()), x_checksum=float(x.double().abs().sum()), t=t, y=y.clone(),
               y_norm=float(y.norm()))
    # config C1: DDIM (4 steps here), eta = 0.5 (eta = 1 gives sqrt of a rounding-negative number at t=999 in the
    # reference itself), u0 + f conditioning, padded 41x60 coefficients
    gd = b.GaussianDiffusion(m, seq_length=(64, 64), is_wavelet=True, pad_mode="periodization", wave_type="bior2.4",
                             padded_shape=[41, 60], ori_shape=[81, 120], timesteps=1000, sampling_timesteps=4,
                             ddim_sampling_eta=0.5, is_condition_u0=True, is_condition_f=True)
    u0 = torch.randn(2, 32, 64, generator=g)
    f = torch.randn(2, 4, 64, 64, generator=g)
    with patched_randn(NoiseTape(12)), torch.no_grad():
        smp = gd.sample(batch_size=2, u_init=u0, f=f)
    out.update(u0_checksum=float(u0.double().abs().sum()), sample=smp.clone(), sample_norm=float(smp.norm()), steps=4,
               eta=0.5, tape_seed=12)
    torch.save(out, os.path.join(HERE, "burgers_unet2d_ddim4.pt"))
    print("burgers", out["weights_checksum"], out["y_norm"], out["sample_norm"])


if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] == "smoke":
        smoke_fixture()
    if len(sys.argv) < 2 or sys.argv[1] == "burgers":
        burgers_fixture()
