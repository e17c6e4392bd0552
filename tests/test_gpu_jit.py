"""GPU: run-time specialisation (NVRTC) of the kernel templates for programs that have no
ahead-of-time instantiation.  XTB_JIT_MIN_ELEMS=0 forces it for small inputs; results must equal
the interpreter's bit for bit (same functor code) and match the oracle."""
import os

import numpy as np
import pytest

from util import assert_bit_exact, interpreter_only, last_kernel, run_both, ulp_distance

pytestmark = pytest.mark.gpu
F32, F64 = np.float32, np.float64


def _set_min(v):
    from xtensor_b200 import capi
    capi.check(capi.lib().xtb_set_option(b"jit_min_elems", v))


@pytest.fixture(autouse=True)
def force_jit():
    _set_min(0)
    yield
    _set_min(1 << 20)


def rnd(shape, dtype=F32, lo=-2.0, hi=2.0, seed=0):
    return np.random.default_rng(seed).uniform(lo, hi, shape).astype(dtype)


CASES = {
    "chain_f32": (lambda xt, A, B: xt.sqrt(xt.abs(A)) * xt.cos(B) - A / (B * B + F32(1.0)), F32, 1),
    "mixed_f64": (lambda xt, A, B: xt.where(A > B, xt.exp(A * 0.5), xt.log1p(xt.abs(B))) + 2.0, F64, 2),
    "clip_fma": (lambda xt, A, B: xt.clip(A, F32(-1), F32(1)) + xt.fma(A, B, A) * xt.tanh(B), F32, 2),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("shape", [(64, 96), (5, 33, 7)])
def test_elementwise_jit(xt, gpu, name, shape):
    f, dt, bar = CASES[name]
    a, b = rnd(shape, dt, seed=1), rnd(shape, dt, seed=2)
    got, want = run_both(xt, lambda A, B: f(xt, A, B), a, b)
    assert "jit" in last_kernel(), last_kernel()
    # a few 1-2 ulp libm calls composed, then subtracted: bound the error by the operand scale
    tol = 2e-6 if dt == F32 else 4e-15
    assert np.all(np.abs(got.astype(F64) - want.astype(F64)) <= tol * (1.0 + np.abs(a).astype(F64) + np.abs(b)) * bar * 4)
    with interpreter_only():
        got_i, _ = run_both(xt, lambda A, B: f(xt, A, B), a, b)
    assert "interp" in last_kernel()
    assert_bit_exact(got, got_i)                   # same functor code either way


def test_jit_is_cached_and_handles_views(xt, gpu):
    a, b = rnd((40, 64), F32, seed=3), rnd((64, 40), F32, seed=4)
    f = lambda A, B: A * F32(3) - xt.transpose(B) * A + xt.sin(A)
    got, want = run_both(xt, f, a, b)
    assert "jit" in last_kernel()
    assert np.allclose(got, want, rtol=2e-6, atol=2e-6)
    k0 = last_kernel()
    got2, _ = run_both(xt, f, a, b)                 # second evaluation: cached kernel
    assert last_kernel() == k0
    assert_bit_exact(got, got2)
    i = np.random.default_rng(5).integers(-50, 50, (33, 65)).astype(np.int32)
    g, w = run_both(xt, lambda I: (I * 3 + (I >> 2)) % 7 - (I & 12), i)
    assert "jit" in last_kernel()
    assert_bit_exact(g, w)


def test_fused_map_reduce_jit(xt, gpu):
    a = np.round(rnd((48, 80), F32, -8, 8, seed=6))
    m = np.round(rnd((80,), F32, -2, 2, seed=7))
    for axes in ([0], [1], [0, 1]):
        d = xt.evaluate(xt.sum(xt.abs(xt.DeviceArray.from_numpy(a) - xt.DeviceArray.from_numpy(m)) * F32(2), axes)).numpy()
        assert "jit" in last_kernel() or "copy" in last_kernel(), last_kernel()
        assert_bit_exact(d, (np.abs(a - m) * 2).sum(axis=tuple(axes)).astype(F32))
    g = xt.evaluate(xt.amax(xt.DeviceArray.from_numpy(a) * xt.DeviceArray.from_numpy(a), [1])).numpy()
    assert_bit_exact(g, (a * a).max(axis=1))


def test_large_expression_uses_jit_by_default(xt, gpu):
    _set_min(1 << 20)
    try:
        a = rnd((1 << 21,), F32, seed=8)
        got, want = run_both(xt, lambda A: xt.sqrt(A * A + F32(1.0)) - xt.abs(A), a)
        assert "jit" in last_kernel()
        assert ulp_distance(got, want) <= 2
        small = rnd((100,), F32, seed=9)
        run_both(xt, lambda A: xt.sqrt(A * A + F32(1.0)) - xt.abs(A), small)
        assert "jit" not in last_kernel()           # below the threshold: interpreter kernel
    finally:
        _set_min(0)
