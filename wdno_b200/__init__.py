"""wdno_b200 -- B200-native engine for the WDNO denoising loop + DWT/IDWT hot path."""
__version__ = "0.1.0"
