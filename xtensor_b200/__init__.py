"""xtensor_b200 -- B200 (sm_100a) evaluation backend for xtensor's hot path.

The product is `lib/libxtb200.so` (hand-written CUDA behind the C ABI in
include/xtb200.h) plus the header-only C++ boundary in include/xtb200/.  The
Python modules here are thin test / benchmark plumbing over that C ABI:

    from xtensor_b200 import expr as xt
    c = xt.DeviceArray.empty((1024, 1024, 64), xt.F32)
    xt.noalias(c).assign(xt.sin(a) * b + np.float32(2.0) * d)
"""
from . import capi  # noqa: F401

__all__ = ["capi", "expr"]
