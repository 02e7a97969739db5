"""Mirror of the reference's `smoke/` import root: put this directory on sys.path (instead of the reference's
smoke/) and `from ddpm.diffusion_2d import GaussianDiffusion`, `from video_diffusion_pytorch.video_diffusion_pytorch_conv3d
import Unet3D_with_Conv3D`, `from wave_trans_2d import ...`, `from ddpm.wave_utils import upsample_coef` resolve to the
B200 engine."""
