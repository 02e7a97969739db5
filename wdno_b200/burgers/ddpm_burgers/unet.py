"""drop-in for /root/reference/burgers/ddpm_burgers/unet.py (Unet2D; Unet1D is never constructed by WDNO)"""
from wdno_b200.unet2d import Unet1D, Unet2D  # noqa: F401
