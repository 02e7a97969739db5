"""drop-in for /root/reference/burgers/wave_trans.py: the packing functions and, as `python -m wdno_b200.burgers.wave_trans`,
the offline job of its __main__ block (wave_trans.py:66-127) on the DWT kernels"""
from wdno_b200.coef_builders import build_burgers_coef_file, burgers_data_to_coef  # noqa: F401
from wdno_b200.packing import burgers_coef_to_tensor as coef_to_tensor  # noqa: F401
from wdno_b200.packing import burgers_tensor_to_coef as tensor_to_coef  # noqa: F401
from wdno_b200.packing import burgers_tensor_to_coef_super as tensor_to_coef_super  # noqa: F401
from wdno_b200.packing import burgers_upsample_coef as upsample_coef  # noqa: F401
from wdno_b200.wavelets import DWT1DForward, DWT1DInverse, DWTForward, DWTInverse  # noqa: F401

if __name__ == "__main__":
    print("Save", build_burgers_coef_file("data/1d/train"))
