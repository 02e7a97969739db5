"""drop-in for /root/reference/smoke/ddpm/diffusion_2d.py: GaussianDiffusion and the beta schedules are the engine's;
every other name of that file (`Trainer`, `Unet`, ... -- smoke/ddpm/utils.py:11, smoke/train_2d.py:5) is served from the
reference's own source, loaded lazily under a private module name (wdno_b200._dropin)."""
from wdno_b200._dropin import reference_attr
from wdno_b200.diffusion_smoke import (GaussianDiffusion, cosine_beta_schedule, linear_beta_schedule,  # noqa: F401
                                       sigmoid_beta_schedule)


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return reference_attr("smoke/ddpm/diffusion_2d.py", name, "_wdno_reference.smoke.ddpm.diffusion_2d")
