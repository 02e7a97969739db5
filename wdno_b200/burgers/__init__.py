"""Mirror of the reference's `burgers/` import root: with this directory on sys.path (instead of the reference's
burgers/), `from ddpm_burgers.diffusion_1d import GaussianDiffusion`, `from ddpm_burgers.unet import Unet2D`,
`from ddpm_burgers.wavelet_utils import upsample_coef, get_wt_T` and `from wave_trans import ...` resolve to the
B200 engine."""
