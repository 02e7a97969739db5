"""mirror of the reference package `burgers/ddpm_burgers`: `diffusion_1d`, `unet`, `wavelet_utils` are the engine's, every
other submodule (`test_util`, `train_diffusion`, `model_utils`, `data_burgers_1d`, `generate_burgers`, `result_io`) falls
through to the reference tree configured by wdno_b200.install()."""
from wdno_b200._dropin import extend_path

__path__ = extend_path(__path__, "burgers/ddpm_burgers")
