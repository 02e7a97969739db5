"""drop-in for /root/reference/smoke/video_diffusion_pytorch/video_diffusion_pytorch_conv3d.py (Unet3D_with_Conv3D)"""
from wdno_b200.unet3d import Unet3D_with_Conv3D  # noqa: F401

Unet3D = Unet3D_with_Conv3D  # north-star name
