"""wdno_b200 -- B200-native engine for the WDNO denoising loop + DWT/IDWT hot path."""
__version__ = "0.2.0"

from ._dropin import install, install_wavelet_modules, reference_root  # noqa: E402,F401
