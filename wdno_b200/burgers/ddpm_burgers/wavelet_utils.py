"""drop-in for /root/reference/burgers/ddpm_burgers/wavelet_utils.py"""
from wdno_b200.packing import burgers_get_wt_T as get_wt_T  # noqa: F401
from wdno_b200.packing import burgers_upsample_coef as upsample_coef  # noqa: F401
