"""drop-in for /root/reference/burgers/ddpm_burgers/diffusion_1d.py"""
from wdno_b200.diffusion_burgers import GaussianDiffusion, GaussianDiffusion1D  # noqa: F401
