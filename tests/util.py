"""Helpers shared by the parity tests."""
import os
from contextlib import contextmanager

import numpy as np


def ulp_distance(got: np.ndarray, want: np.ndarray) -> int:
    """Max distance in units in the last place between two float arrays (NaN == NaN)."""
    assert got.dtype == want.dtype and got.shape == want.shape, (got.dtype, want.dtype, got.shape, want.shape)
    if got.size == 0:
        return 0
    it = np.int32 if got.dtype == np.float32 else np.int64
    a = got.view(it).astype(np.int64)
    b = want.view(it).astype(np.int64)
    # map the sign-magnitude float ordering onto a monotone integer line
    sign = np.int64(np.iinfo(it).min)
    a = np.where(a < 0, sign - a, a)
    b = np.where(b < 0, sign - b, b)
    d = np.abs(a - b)
    both_nan = np.isnan(got) & np.isnan(want)
    d = np.where(both_nan, 0, d)
    return int(d.max())


def run_both(xt, f, *arrays, dtype=None):
    """Evaluate f(leaves...) once on the GPU (libxtb200) and once on the CPU oracle."""
    dev = [xt.DeviceArray.from_numpy(a) if isinstance(a, np.ndarray) else a for a in arrays]
    host = [xt.HostArray.from_numpy(a) if isinstance(a, np.ndarray) else a for a in arrays]
    got = xt.evaluate(f(*dev), dtype).numpy()
    want = xt.evaluate(f(*host), dtype).numpy()
    return got, want


def assert_bit_exact(got, want):
    assert got.dtype == want.dtype, (got.dtype, want.dtype)
    assert got.shape == want.shape, (got.shape, want.shape)
    if got.dtype.kind == "f":
        assert ulp_distance(got, want) == 0, f"max ulp distance {ulp_distance(got, want)}"
    else:
        assert np.array_equal(got, want)


@contextmanager
def interpreter_only():
    """Force the run-time interpreter kernels (no compile-time program match)."""
    from xtensor_b200 import capi
    capi.check(capi.lib().xtb_set_option(b"no_static", 1))
    try:
        yield
    finally:
        capi.check(capi.lib().xtb_set_option(b"no_static", 0))


def last_kernel():
    from xtensor_b200 import capi
    return capi.lib().xtb_last_kernel().decode()
