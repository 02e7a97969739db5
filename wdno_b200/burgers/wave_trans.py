"""drop-in for the functions of /root/reference/burgers/wave_trans.py (the offline __main__ job is out of scope)"""
from wdno_b200.packing import burgers_coef_to_tensor as coef_to_tensor  # noqa: F401
from wdno_b200.packing import burgers_tensor_to_coef as tensor_to_coef  # noqa: F401
from wdno_b200.packing import burgers_tensor_to_coef_super as tensor_to_coef_super  # noqa: F401
from wdno_b200.packing import burgers_upsample_coef as upsample_coef  # noqa: F401
from wdno_b200.wavelets import DWT1DForward, DWT1DInverse, DWTForward, DWTInverse  # noqa: F401
