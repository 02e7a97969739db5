"""drop-in for /root/reference/smoke/ddpm/wave_utils.py"""
from wdno_b200.packing import smoke_upsample_coef as upsample_coef  # noqa: F401
