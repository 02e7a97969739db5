"""drop-in for /root/reference/smoke/ddpm/diffusion_2d.py (GaussianDiffusion only; Trainer is out of scope)"""
from wdno_b200.diffusion_smoke import (GaussianDiffusion, cosine_beta_schedule, linear_beta_schedule,  # noqa: F401
                                       sigmoid_beta_schedule)
