"""drop-in for /root/reference/burgers/ddpm_burgers/diffusion_1d.py: the two diffusion classes are the engine's; other
names of that file are served from the reference's own source (wdno_b200._dropin)."""
from wdno_b200._dropin import reference_attr
from wdno_b200.diffusion_burgers import GaussianDiffusion, GaussianDiffusion1D  # noqa: F401


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return reference_attr("burgers/ddpm_burgers/diffusion_1d.py", name, "_wdno_reference.burgers.ddpm_burgers.diffusion_1d")
