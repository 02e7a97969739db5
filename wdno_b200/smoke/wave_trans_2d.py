"""drop-in for the functions of /root/reference/smoke/wave_trans_2d.py (the offline __main__ job is out of scope)"""
from wdno_b200.packing import smoke_coef_to_tensor as coef_to_tensor  # noqa: F401
from wdno_b200.packing import smoke_tensor_to_coef as tensor_to_coef  # noqa: F401
from wdno_b200.wavelets import DWT1DForward, DWT1DInverse, DWTForward, DWTInverse, wavedec3, waverec3  # noqa: F401
