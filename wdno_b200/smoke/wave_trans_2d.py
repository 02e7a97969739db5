"""drop-in for /root/reference/smoke/wave_trans_2d.py: the packing functions and, as `python -m wdno_b200.smoke.wave_trans_2d`,
the offline job of its __main__ block (wave_trans_2d.py:61-189) on the DWT kernels"""
from wdno_b200.coef_builders import build_smoke_coef_files, smoke_sims_to_coef  # noqa: F401
from wdno_b200.packing import smoke_coef_to_tensor as coef_to_tensor  # noqa: F401
from wdno_b200.packing import smoke_tensor_to_coef as tensor_to_coef  # noqa: F401
from wdno_b200.wavelets import DWT1DForward, DWT1DInverse, DWTForward, DWTInverse, wavedec3, waverec3  # noqa: F401

if __name__ == "__main__":
    print("Max", build_smoke_coef_files("./data/2d/", "train/", range(20000)))
