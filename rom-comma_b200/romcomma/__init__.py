"""rom-comma's dense-GP hot path (GPR fit/predict, closed Sobol indices) on hand-written sm_100a kernels.

Same package / module / class names as the reference ``romcomma`` so callers switch by changing ``sys.path``; TensorFlow and
GPflow are not imported - the FLOPs run in ``csrc/librc_b200.so`` through ``romcomma._capi`` (ctypes).
"""
from . import base, data, gpf, gpr, gsa, user  # noqa: F401
