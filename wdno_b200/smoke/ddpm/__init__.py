"""mirror of the reference package `smoke/ddpm`: `diffusion_2d` and `wave_utils` are the engine's, every other submodule
(`data_2d`, `utils`, `modules`) falls through to the reference tree configured by wdno_b200.install()."""
from wdno_b200._dropin import extend_path

__path__ = extend_path(__path__, "smoke/ddpm")
