"""GPU parity: xtb_reduce (sum / prod / amax / amin / mean / variance) vs the CPU oracle.

Integer-valued data makes every summation order exact, so the axis / stride / merge logic
is checked bit-exactly; random data is checked against the reference's
sequential-order result within the north-star tolerance (1e-6 fp32 / 1e-12 fp64, relative
to the magnitude of the summands) and reported against an fp64 sum as well.
Mirrors test/test_xreducer.cpp: fixture ones({3,2,4,6,5}) with axes {1,3} (:55-84, 230-283),
sum_all == 732 (:475-481), lazy == immediate sweep (:704-771), keep_dims / initial /
ones_first / empty_axes / zero_shape (:964-1079), axis errors (:213-220, 1081-1088),
uint8 no-overflow (:471-472), mean (:542-577), amin/amax (test/test_xmath.cpp:147-168).
"""
import itertools

import numpy as np
import pytest

from util import assert_bit_exact, interpreter_only, last_kernel

pytestmark = pytest.mark.gpu
F32, F64 = np.float32, np.float64


def ints(shape, dtype=F32, lo=-8, hi=8, seed=0):
    return np.random.default_rng(seed).integers(lo, hi + 1, shape).astype(dtype)


def both(xt, build, a, mode=0):
    g = xt.evaluate(build(xt.DeviceArray.from_numpy(a))).numpy()
    r = build(xt.HostArray.from_numpy(a))
    w = (xt._run_reducer(r, xt.HostArray, mode=mode) if isinstance(r, xt.Reducer) else xt.evaluate(r)).numpy()
    return g, w


def test_reducer_fixture_12_24_732(xt, gpu):
    a = np.ones((3, 2, 4, 6, 5), dtype=F64)
    a[1, :, 1, :, 1] = 2        # m_a(1, i, 1, j, 1) = 2 (test_xreducer.cpp:77-83)
    for kind in (xt.DeviceArray, xt.HostArray):
        A = kind.from_numpy(a)
        r = xt.evaluate(xt.sum(A, [1, 3])).numpy()
        assert r.shape == (3, 4, 5)
        assert r[0, 0, 0] == 12 and r[1, 1, 1] == 24
        assert np.array_equal(r, a.sum(axis=(1, 3)))
        assert float(xt.evaluate(xt.sum(A)).numpy()) == 732


AXES_SWEEP = [[0], [1], [2], [3], [0, 1], [1, 2], [2, 3], [0, 2], [0, 3], [1, 3], [0, 1, 2], [1, 2, 3], [0, 2, 3],
              [0, 1, 2, 3], [-1], [-3, -1]]


@pytest.mark.parametrize("axes", AXES_SWEEP)
@pytest.mark.parametrize("red", ["sum", "amax", "amin", "prod"])
def test_axis_sweep_bit_exact(xt, gpu, axes, red):
    """lazy == immediate == device on integer-valued data, every axis subset of a 4-D array."""
    lo, hi = (-8, 8) if red != "prod" else (1, 2)
    a = ints((6, 5, 7, 9), F32, lo, hi, seed=len(axes))
    f = lambda A: getattr(xt, red)(A, axes)
    g, w0 = both(xt, f, a, mode=0)
    _, w1 = both(xt, f, a, mode=1)
    assert_bit_exact(w0, w1)          # lazy == immediate (test_xreducer.cpp:704-771)
    assert_bit_exact(g, w0)
    npf = {"sum": np.sum, "amax": np.max, "amin": np.min, "prod": np.prod}[red]
    assert_bit_exact(g, npf(a, axis=tuple(axes)).astype(F32))


@pytest.mark.parametrize("shape,axes", [((64, 48, 16), [0]), ((64, 48, 16), [2]), ((64, 48, 16), [1]),
                                        ((4096, 24), [0]), ((24, 4096), [1]), ((100000, 10), [0]),
                                        ((10, 100000), [1]), ((10, 100000), [0, 1]), ((300001,), [0]),
                                        ((7, 33, 5), [0, 2]), ((33, 1, 65), [1]), ((1, 1, 9), [0, 1])])
@pytest.mark.parametrize("dtype", [F32, F64, np.int32])
def test_shapes_bit_exact(xt, gpu, shape, axes, dtype):
    """cfg3-shaped (strided axis 0 / contiguous axis 2), benchmark_reducer shapes, splits, tails."""
    a = ints(shape, dtype, -4, 4, seed=3)
    for red in ("sum", "amax"):
        f = lambda A: getattr(xt, red)(A, axes)
        g, w = both(xt, f, a)
        assert_bit_exact(g, w)
        with interpreter_only():
            g2, _ = both(xt, f, a)
        assert "interp" in last_kernel()
        assert_bit_exact(g2, w)


@pytest.mark.parametrize("shape,axes", [((512, 300), [0]), ((300, 512), [1]), ((64, 64, 16), [0]), ((64, 64, 16), [2]),
                                        ((200000,), [0])])
@pytest.mark.parametrize("dtype,tol", [(F32, 1e-6), (F64, 1e-12)])
def test_random_data_tolerance(xt, gpu, shape, axes, dtype, tol):
    """Reorder-sensitive sums: |gpu - ref| <= tol * sum|x| against the sequential reference order."""
    a = np.random.default_rng(6).uniform(-1, 1, shape).astype(dtype)
    g, w = both(xt, lambda A: xt.sum(A, axes), a)
    scale = np.abs(a).sum(axis=tuple(axes), dtype=F64)
    err = np.abs(g.astype(F64) - w.astype(F64))
    assert np.all(err <= tol * np.maximum(scale, 1.0)), float((err / np.maximum(scale, 1)).max())
    exact = a.astype(F64).sum(axis=tuple(axes))
    # the device result must not be further from the fp64 truth than the sequential reference
    assert np.abs(g - exact).max() <= np.abs(w - exact).max() * 1.5 + np.finfo(dtype).eps * scale.max()
    gm, wm = both(xt, lambda A: xt.amax(A, axes), a)
    assert_bit_exact(gm, wm)


def test_keep_dims_initial_ones_first_empty_axes(xt, gpu):
    a = ints((3, 4, 5), F64)
    g, w = both(xt, lambda A: xt.sum(A, [1], keep_dims=True), a)
    assert g.shape == (3, 1, 5)
    assert_bit_exact(g, w)
    g, w = both(xt, lambda A: xt.sum(A, [0, 2], keep_dims=True, initial=5), a)
    assert g.shape == (1, 4, 1)
    assert_bit_exact(g, w)
    assert_bit_exact(g, a.sum(axis=(0, 2), keepdims=True) + 5)
    g, w = both(xt, lambda A: xt.amax(A, [2], initial=3), a)
    assert_bit_exact(g, np.maximum(a.max(axis=2), 3))
    b = ints((1, 1, 4, 5), F32)                   # size-1 leading dims (test_xreducer.cpp:1020-1031)
    for axes in ([0], [1], [0, 1], [2], [0, 2], [3]):
        g, w = both(xt, lambda A: xt.sum(A, axes), b)
        assert_bit_exact(g, w)
        assert_bit_exact(g, b.sum(axis=tuple(axes)))
    g, w = both(xt, lambda A: xt.sum(A, []), a)   # empty axes: identity map reduce(init, x)
    assert_bit_exact(g, a)


def test_zero_size(xt, gpu):
    """zero_shape / empty_array (test_xreducer.cpp:1044-1079): reduced extent 0 yields init."""
    a = np.zeros((0, 4), F32)
    g, w = both(xt, lambda A: xt.sum(A, [0]), a)
    assert g.shape == (4,) and np.array_equal(g, np.zeros(4, F32))
    assert_bit_exact(g, w)
    g, w = both(xt, lambda A: xt.prod(A, [0]), a)
    assert np.array_equal(g, np.ones(4, F32))
    g, w = both(xt, lambda A: xt.sum(A, [1]), a)
    assert g.shape == (0,)


def test_axis_errors(xt, gpu):
    a = xt.DeviceArray.from_numpy(ints((3, 4, 5)))
    with pytest.raises(RuntimeError, match="sorted"):
        xt.evaluate(xt.sum(a, [1, 0]))
    with pytest.raises(RuntimeError, match="duplicates"):
        xt.evaluate(xt.sum(a, [1, 1]))
    with pytest.raises(RuntimeError, match="out of bounds"):
        xt.evaluate(xt.sum(a, [3]))


def test_uint8_no_overflow_and_promotions(xt, gpu):
    u = np.full((1000,), 255, np.uint8)            # sum is int, no overflow (test_xreducer.cpp:471-472)
    g, w = both(xt, lambda A: xt.sum(A), u)
    assert g.dtype == np.int32 and int(g) == 255000
    s = ints((40, 50), np.int16, -300, 300)
    g, w = both(xt, lambda A: xt.sum(A, [0]), s)
    assert g.dtype == np.int32
    assert_bit_exact(g, w)
    f = ints((40, 50), F32)
    g, w = both(xt, lambda A: xt.sum(A, [1], dtype=xt.F64), f)   # sum<double>(float)
    assert g.dtype == F64
    assert_bit_exact(g, w)
    g, w = both(xt, lambda A: xt.amin(A, [1]), ints((9, 300), np.uint8, 0, 255))
    assert g.dtype == np.uint8
    assert_bit_exact(g, w)


def test_mean_and_variance(xt, gpu):
    """mean of fp32 has value_type double (test_xmath_result_type.cpp:237-238); two-pass variance."""
    a = ints((60, 70), F32, -8, 8, seed=8)
    g, w = both(xt, lambda A: xt.mean(A, [0]), a)
    assert g.dtype == F64
    assert_bit_exact(g, w)
    assert_bit_exact(g, a.sum(axis=0).astype(F64) / 60.0)
    g, w = both(xt, lambda A: xt.mean(A, [1], dtype=xt.F32), a)
    assert g.dtype == F32
    assert_bit_exact(g, w)
    r = np.random.default_rng(1).uniform(-1, 1, (4, 5, 6, 7))
    g, w = both(xt, lambda A: xt.variance(A, [0, 2]), r)         # axes of test_extended_xmath_reducers
    assert np.allclose(g, w, rtol=1e-12, atol=0)
    assert np.allclose(g, r.var(axis=(0, 2)), rtol=1e-5, atol=1e-8)   # xt::allclose defaults
    g, w = both(xt, lambda A: xt.variance(A, [0, 2], ddof=1), r)
    assert np.allclose(g, r.var(axis=(0, 2), ddof=1), rtol=1e-5, atol=1e-8)
    g, w = both(xt, lambda A: xt.stddev(A, [1]), r)
    assert np.allclose(g, r.std(axis=1), rtol=1e-5, atol=1e-8)


def test_fused_map_reduce_and_views(xt, gpu):
    """Reducers consume any expression (chaining_reducers, test_xreducer.cpp:773-781, 1102-1124)."""
    a, m = ints((50, 64), F32), ints((64,), F32, seed=2)
    g, w = both(xt, lambda A: xt.sum(xt.square(A - A[0]), [0]), a)
    assert_bit_exact(g, w)
    assert_bit_exact(g, ((a - a[0]) ** 2).sum(axis=0))
    g, w = both(xt, lambda A: xt.sum(xt.transpose(A), [0]), a)   # transposed operand
    assert_bit_exact(g, a.T.sum(axis=0))
    g, w = both(xt, lambda A: xt.sum(A[::2, 1:63], [1]), a)      # strided, misaligned view
    assert_bit_exact(g, a[::2, 1:63].sum(axis=1))
    g, w = both(xt, lambda A: xt.sum(xt.sum(A, [0]) * F32(2)), a)  # reducer inside an expression
    assert float(g) == float(a.sum() * 2)


# ---- reductions rewritten into two passes (xtb_reduce.cu: reduce_decomposed) and scalar-access kernels ----------
DECOMP_CASES = [
    # narrow: innermost kept, < 1024 outputs
    ((1 << 15, 64), [0]), ((40000, 33), [0]), ((1 << 20, 3), [0]), ((100003, 16), [0]), ((512, 512, 16), [0, 1]),
    ((300, 70, 50, 6), [0, 2]), ((70001, 20), [0]),
    # mixed: innermost reduced together with an outer axis, kept dim between
    ((512, 256, 16), [0, 2]), ((64, 32, 40, 16), [0, 3]), ((64, 32, 40, 16), [1, 3]), ((300, 17, 257), [0, 2]),
]


@pytest.mark.parametrize("shape,axes", DECOMP_CASES)
@pytest.mark.parametrize("dtype", [F32, np.int32, F64])
def test_decomposed_reductions_bit_exact(xt, gpu, shape, axes, dtype):
    """Two-pass formulation == single-pass kernels == numpy on integer-valued data (sum, amax, keep_dims, initial, mean)."""
    a = ints(shape, dtype, -4, 4, seed=11)
    A = xt.DeviceArray.from_numpy(a)
    for red, npf in (("sum", np.sum), ("amax", np.max)):
        got = xt.evaluate(getattr(xt, red)(A, axes)).numpy()
        assert_bit_exact(got, npf(a, axis=tuple(axes)).astype(dtype))
        gpu.xtb_set_option(b"no_decompose", 1)
        try:
            single = xt.evaluate(getattr(xt, red)(A, axes)).numpy()
        finally:
            gpu.xtb_set_option(b"no_decompose", 0)
        assert_bit_exact(got, single)
    kd = xt.evaluate(xt.sum(A, axes, keep_dims=True)).numpy()
    assert_bit_exact(kd, a.sum(axis=tuple(axes), keepdims=True).astype(dtype))
    ini = xt.evaluate(xt.sum(A, axes, initial=dtype(100))).numpy()
    assert_bit_exact(ini, (a.sum(axis=tuple(axes)) + 100).astype(dtype))
    if dtype != np.int32:
        mean = xt.evaluate(xt.mean(A, axes)).numpy()
        n = np.prod([shape[x] for x in axes])
        assert mean.dtype == F64
        assert_bit_exact(mean, a.sum(axis=tuple(axes)).astype(F64) / F64(n))


def test_decomposed_fused_map_with_broadcast_leaf(xt, gpu):
    """sum(square(a - m), {0}) with m broadcast along the split axis (variance second pass on a narrow matrix)."""
    a = ints((1 << 16, 48), F32, -4, 4, seed=5)
    m = ints((48,), F32, -2, 2, seed=6)
    got = xt.evaluate(xt.sum(xt.square(xt.DeviceArray.from_numpy(a) - xt.DeviceArray.from_numpy(m)), [0])).numpy()
    assert_bit_exact(got, np.square(a - m).sum(axis=0).astype(F32))
    v = xt.evaluate(xt.variance(xt.DeviceArray.from_numpy(a), [0])).numpy()
    want = np.square(a - a.mean(axis=0, dtype=F64)).mean(axis=0)
    assert np.allclose(v, want, rtol=1e-5)


@pytest.mark.parametrize("shape,axes", [((2047, 1022), [0]), ((2047, 1022), [1]), ((513, 255, 15), [0]), ((513, 255, 15), [2]),
                                        ((1025, 1023), [0, 1]), ((300001, 7), [0])])
@pytest.mark.parametrize("dtype", [F32, F64, np.int32])
def test_unaligned_shapes_use_compiled_kernels(xt, gpu, shape, axes, dtype):
    """Odd extents (row pitch not a multiple of 16 bytes) run the scalar-access instantiation of the same kernels,
    not the interpreter, and match numpy bit for bit."""
    a = ints(shape, dtype, -4, 4, seed=12)
    A = xt.DeviceArray.from_numpy(a)
    for red, npf in (("sum", np.sum), ("amax", np.max)):
        got = xt.evaluate(getattr(xt, red)(A, axes)).numpy()
        assert "interp" not in last_kernel(), last_kernel()
        assert_bit_exact(got, npf(a, axis=tuple(axes)).astype(dtype))
    # an offset view of an aligned container: base misaligned by one element
    b = ints((1024, 1024), dtype, -4, 4, seed=13)
    B = xt.DeviceArray.from_numpy(b)
    got = xt.evaluate(xt.sum(xt.view(B, slice(None), slice(1, None)), [0])).numpy()
    assert_bit_exact(got, b[:, 1:].sum(axis=0).astype(dtype))
