"""GPU parity: xtb_assign (fused broadcast elementwise) vs the CPU oracle.

Bar: bit-exact for integer dtypes and for + - * / (device built with -fmad=false,
oracle with -ffp-contract=off); <= 2 ulp for transcendentals (CUDA libm vs glibc).
Shapes mirror the reference's own tests: the 3x2x4 layout fixtures of
test/test_common.hpp:136-200, the broadcast shapes of
test/test_extended_broadcast_view.cpp:126-810, the strided store of
test/test_strided_assign.cpp:178-196 -- plus small versions of BASELINE cfg1/2/4/5.
"""
import numpy as np
import pytest

from util import ulp_bar, assert_bit_exact, interpreter_only, last_kernel, run_both, ulp_distance

pytestmark = pytest.mark.gpu

F32, F64 = np.float32, np.float64


def rnd(shape, dtype=F32, lo=-3.0, hi=3.0, seed=0):
    rng = np.random.default_rng(seed)
    if np.dtype(dtype).kind == "f":
        return rng.uniform(lo, hi, shape).astype(dtype)
    return rng.integers(int(lo), int(hi) + 1, shape).astype(dtype)


@pytest.mark.parametrize("n", [1, 3, 4, 255, 256, 1000, 4096 + 7, 1 << 20])
@pytest.mark.parametrize("dtype", [F32, F64])
def test_cfg1_contiguous_add(xt, gpu, n, dtype):
    """noalias(c) = a + b on 1-D containers (benchmark_assign style, linear_assigner path)."""
    a, b = rnd((n,), dtype, seed=1), rnd((n,), dtype, seed=2)
    got, want = run_both(xt, lambda A, B: A + B, a, b)
    assert_bit_exact(got, want)
    if n > 1:
        assert "add_f" in last_kernel()      # compile-time program; n == 1 has nothing to vectorise
    with interpreter_only():
        got2, _ = run_both(xt, lambda A, B: A + B, a, b)
    assert "interp" in last_kernel()
    assert_bit_exact(got2, want)


@pytest.mark.parametrize("shape", [(4, 6, 8), (3, 5, 7), (16, 32, 64), (2, 1024, 64)])
def test_cfg2_fused_broadcast(xt, gpu, shape):
    """c = sin(a) * b(1,N,1) + 2.0f * d: value type stays float; <= 2 ulp (sin)."""
    a, d = rnd(shape, seed=3), rnd(shape, seed=5)
    b = rnd((1, shape[1], 1), lo=0.5, hi=1.5, seed=4)
    f = lambda A, B, D: xt.sin(A) * B + F32(2.0) * D
    got, want = run_both(xt, f, a, b, d)
    assert got.dtype == F32
    # sin is <= 2 ulp; the final add can cancel, so the bound is 2 ulp of the larger term (+ its own rounding)
    gs, ws = run_both(xt, lambda A: xt.sin(A), a)
    assert ulp_distance(gs, ws) <= 2
    # propagate it: 2 ulp(sin) * |b|, plus one rounding each for the product and the final sum
    sn = np.sin(a)
    bound = (2 * np.spacing(np.abs(sn)) * np.abs(b) + np.spacing(np.abs(sn * b))
             + np.spacing(np.maximum(np.abs(want), np.abs(got)))).astype(F64)
    assert np.all(np.abs(got.astype(F64) - want.astype(F64)) <= bound)
    with interpreter_only():
        got2, _ = run_both(xt, f, a, b, d)
    assert_bit_exact(got2, got)  # interpreter and compile-time program share the functor code


@pytest.mark.parametrize("n", [4, 33, 64, 130])
def test_cfg4_transpose_plus_strided_view(xt, gpu, n):
    """out = transpose(a) + view(b, range(0, _, 2), all()) on fp64."""
    a, b = rnd((n, n), F64, -1, 1, 7), rnd((2 * n, n), F64, -1, 1, 8)
    got, want = run_both(xt, lambda A, B: xt.transpose(A) + xt.view(B, slice(0, None, 2), slice(None)), a, b)
    assert_bit_exact(got, want)
    assert_bit_exact(got, a.T + b[::2])


@pytest.mark.parametrize("ni,nj", [(64, 32), (64, 64), (130, 98), (1000, 776), (515, 1026), (2048, 4096)])
@pytest.mark.parametrize("dtype", [F64, F32, np.int32])
def test_transposed_leaf_tma_tile(xt, gpu, ni, nj, dtype):
    """The all-TMA tile kernel (xtb_ew_tma.cuh): transposed leaf in 128-byte-swizzled boxes, direct leaf and result
    through their own boxes; edge tiles are zero-filled / clipped by the TMA unit.  Bit-exact against numpy (one add /
    multiply-add per element), identical to the register-staged tile kernel (no_tma) and to the interpreter."""
    rng = np.random.default_rng(ni * 7 + nj)
    if np.dtype(dtype).kind == "f":
        a, b = rng.uniform(-1, 1, (nj, ni)).astype(dtype), rng.uniform(-1, 1, (2 * ni, nj)).astype(dtype)
    else:
        a, b = rng.integers(-1000, 1000, (nj, ni)).astype(dtype), rng.integers(-1000, 1000, (2 * ni, nj)).astype(dtype)
    f = lambda A, B: xt.transpose(A) + xt.view(B, slice(0, None, 2), slice(None))
    got = xt.evaluate(f(xt.DeviceArray.from_numpy(a), xt.DeviceArray.from_numpy(b))).numpy()
    k = last_kernel()
    assert_bit_exact(got, a.T + b[::2])
    aligned = (ni * np.dtype(dtype).itemsize) % 16 == 0 and (nj * np.dtype(dtype).itemsize) % 16 == 0
    assert ("k_ew_tile_tma" in k) == aligned, k            # 16-byte global strides are what TMA needs
    capi_lib = gpu
    capi_lib.xtb_set_option(b"no_tma", 1)
    try:
        got2 = xt.evaluate(f(xt.DeviceArray.from_numpy(a), xt.DeviceArray.from_numpy(b))).numpy()
        assert "k_ew_tile_tma" not in last_kernel()
    finally:
        capi_lib.xtb_set_option(b"no_tma", 0)
    assert_bit_exact(got2, got)
    # a transposed leaf alone (out = transpose(a)) and two transposed leaves
    A = xt.DeviceArray.from_numpy(a)
    assert_bit_exact(xt.evaluate(xt.transpose(A) - xt.transpose(A) * xt.transpose(A)).numpy(), a.T - a.T * a.T)
    # a view of the transposed leaf that starts mid-buffer and a destination view
    if ni >= 130:
        o = xt.DeviceArray.from_numpy(np.zeros((ni, nj + 8), dtype))
        xt.noalias(o[:, 4:nj + 4]).assign(f(xt.DeviceArray.from_numpy(a), xt.DeviceArray.from_numpy(b)))
        exp = np.zeros((ni, nj + 8), dtype)
        exp[:, 4:nj + 4] = a.T + b[::2]
        assert_bit_exact(o.numpy(), exp)


@pytest.mark.parametrize("rows,cols", [(8, 16), (64, 8192 // 8), (5, 37)])
def test_cfg5_exp_minus_mean(xt, gpu, rows, cols):
    a, m = rnd((rows, cols), lo=-1, hi=1, seed=9), rnd((cols,), lo=-0.1, hi=0.1, seed=10)
    got, want = run_both(xt, lambda A, M: xt.exp(A - M), a, m)
    assert ulp_distance(got, want) <= 2


def test_benchmark_assign_axmby(xt, gpu):
    """noalias(res) = 3.0 * x - 2.0 * y (benchmark/benchmark_assign.cpp:80-90)."""
    x, y = rnd((64, 48), F64, seed=11), rnd((64, 48), F64, seed=12)
    got, want = run_both(xt, lambda X, Y: 3.0 * X - 2.0 * Y, x, y)
    assert_bit_exact(got, want)
    assert_bit_exact(got, 3.0 * x - 2.0 * y)


# -- the reference's layout fixtures: 3x2x4 tensor in several layouts (test/test_common.hpp:136-200)
@pytest.mark.parametrize("op", ["+", "-", "*", "/"])
@pytest.mark.parametrize("layout", ["rm", "cm", "ctm"])
def test_layout_mixing_operation_tester(xt, gpu, op, layout):
    """operation_tester (test/test_xsemantic.hpp:27-87): a op x with x in another layout."""
    vals = np.array([-1] + list(range(1, 24)), dtype=np.int32).reshape(3, 2, 4)
    perm = {"rm": (0, 1, 2), "cm": (2, 1, 0), "ctm": (0, 2, 1)}[layout]
    # store x so that its logical (3,2,4) view has the requested strides
    stored = np.ascontiguousarray(vals.transpose(perm))
    inv = np.argsort(perm)

    def f(A, X):
        Xv = X.transpose(list(inv))
        assert Xv.shape == (3, 2, 4)
        return {"+": A + Xv, "-": A - Xv, "*": A * Xv, "/": A / Xv}[op]

    a = (vals * 2 + 1).astype(np.int32)
    got, want = run_both(xt, f, a, stored)
    assert_bit_exact(got, want)
    ref = {"+": a + vals, "-": a - vals, "*": a * vals, "/": (a / vals).astype(np.int64)}[op]
    assert np.array_equal(got, np.trunc(ref).astype(np.int32))


def test_unit_shape_stride0(xt, gpu):
    """3x1x4 operand with strides {4,0,1} against 3x2x4 (test/test_common.hpp:175-200)."""
    a = rnd((3, 2, 4), np.int32, -20, 20)
    u = rnd((3, 1, 4), np.int32, -20, 20, seed=1)
    got, want = run_both(xt, lambda A, U: A + U, a, u)
    assert_bit_exact(got, want)
    assert np.array_equal(got, a + u)


@pytest.mark.parametrize("shapes", [((5, 1, 7), (1, 5, 1, 7)), ((7,), (5, 1, 7)), ((5, 1, 7), (1, 1, 1, 7)),
                                    ((1, 5, 1, 7), (2, 5, 4, 7))])
def test_extended_broadcast_shapes(xt, gpu, shapes):
    """numpy-generated broadcasting cases of test/test_extended_broadcast_view.cpp:126-810."""
    a, b = rnd(shapes[0], F64, seed=42), rnd(shapes[1], F64, seed=43)
    got, want = run_both(xt, lambda A, B: A + B, a, b)
    assert_bit_exact(got, want)
    assert_bit_exact(got, a + b)


def test_incompatible_shapes_throw(xt, gpu):
    a, b = xt.DeviceArray.from_numpy(rnd((3, 4))), xt.DeviceArray.from_numpy(rnd((5,)))
    with pytest.raises(xt.BroadcastError):
        xt.evaluate(a + b)
    out = xt.DeviceArray.empty((3, 5), xt.F32)
    with pytest.raises(xt.BroadcastError):
        xt.assign(out, a)


def test_strided_store_into_view(xt, gpu):
    """Assign into a strided view: expected buffer of test/test_strided_assign.cpp:178-196 style."""
    buf = np.full((4, 6), -1, dtype=np.int32)
    src = np.arange(1, 9, dtype=np.int32).reshape(2, 4)
    for kind in (xt.DeviceArray, xt.HostArray):
        o = kind.from_numpy(buf)
        xt.noalias(o[1:3, 2:6]).assign(kind.from_numpy(src))
        res = o.numpy()
        exp = buf.copy()
        exp[1:3, 2:6] = src
        assert np.array_equal(res, exp)


UNARY = ["abs", "exp", "exp2", "expm1", "log", "log10", "log2", "log1p", "sqrt", "cbrt", "sin", "cos", "tan", "asin",
         "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh", "erf", "erfc", "tgamma", "lgamma", "ceil",
         "floor", "trunc", "round", "nearbyint", "rint", "sign", "deg2rad", "rad2deg", "square", "cube"]
DOMAIN = {"log": (0.01, 50), "log10": (0.01, 50), "log2": (0.01, 50), "log1p": (-0.9, 50), "sqrt": (0, 50),
          "asin": (-1, 1), "acos": (-1, 1), "acosh": (1, 50), "atanh": (-0.99, 0.99), "tgamma": (0.1, 20),
          "lgamma": (0.1, 50), "exp": (-20, 20), "exp2": (-20, 20), "expm1": (-5, 5), "sinh": (-10, 10),
          "cosh": (-10, 10)}
# measured CUDA-libm vs glibc distance; the north-star bar is 2 ulp
EXACT = {"abs", "sqrt", "ceil", "floor", "trunc", "round", "nearbyint", "rint", "sign", "square", "cube", "deg2rad",
         "rad2deg"}


@pytest.mark.parametrize("name", UNARY)
@pytest.mark.parametrize("dtype", [F32, F64])
def test_unary_functors(xt, gpu, name, dtype):
    lo, hi = DOMAIN.get(name, (-6.0, 6.0))
    a = rnd((257, 33), dtype, lo, hi, seed=hash(name) % 1000)
    got, want = run_both(xt, lambda A: getattr(xt, name)(A), a)
    d = ulp_distance(got, want)
    bar = ulp_bar(name, "f32" if dtype == F32 else "f64", EXACT)    # 2 ulp; the five documented deviations: tests/util.py
    assert d <= bar, f"{name}/{np.dtype(dtype).name}: {d} ulp"


@pytest.mark.parametrize("name,dtype", [("erfc", F32), ("tgamma", F32), ("erfc", F64), ("tgamma", F64), ("lgamma", F64)])
def test_documented_ulp_deviations(xt, gpu, name, dtype):
    """The five functor x dtype pairs that sit more than 2 ulp from glibc (tests/util.py: ULP_DEVIATIONS).  Measured
    against the TRUE value (mpmath, 50 digits): the device must be at least as accurate as the reference's own libm
    (within half an ulp of slack), and its distance to glibc must stay within the documented bar."""
    import mpmath
    from util import ULP_DEVIATIONS
    mpmath.mp.dps = 50
    tag = "f32" if dtype == F32 else "f64"
    lo, hi = {"erfc": (-3.0, 9.0), "tgamma": (0.05, 25.0), "lgamma": (0.05, 60.0)}[name]
    a = rnd((4096,), dtype, lo, hi, seed=17)
    got, want = run_both(xt, lambda A: getattr(xt, name)(A), a)
    fn = {"erfc": mpmath.erfc, "tgamma": mpmath.gamma, "lgamma": mpmath.loggamma}[name]
    err_dev = err_ref = 0.0
    for x, g, w in zip(a, got, want):
        t = fn(mpmath.mpf(float(x)))
        ulp = mpmath.mpf(float(np.spacing(dtype(float(t)))))          # one ulp of the correctly rounded result
        if ulp == 0 or not np.isfinite(float(t)):
            continue
        err_dev = max(err_dev, float(abs(mpmath.mpf(float(g)) - t) / ulp))
        err_ref = max(err_ref, float(abs(mpmath.mpf(float(w)) - t) / ulp))
    d = ulp_distance(got, want)
    print(f"{name}/{tag}: device {err_dev:.2f} ulp from the true value, glibc {err_ref:.2f} ulp, device vs glibc {d} ulp")
    assert d <= ULP_DEVIATIONS[(name, tag)]
    assert err_dev <= max(2.0, err_ref + 0.5), (err_dev, err_ref)
    if dtype == F32:
        assert err_dev <= 1.0          # evaluated in fp64 and rounded once


@pytest.mark.parametrize("name", ["fmod", "remainder", "fmax", "fmin", "fdim", "pow", "hypot", "atan2", "maximum",
                                  "minimum"])
def test_binary_functors(xt, gpu, name):
    a, b = rnd((129, 65), F32, 0.1, 9, seed=1), rnd((129, 65), F32, 0.1, 4, seed=2)
    got, want = run_both(xt, lambda A, B: getattr(xt, name)(A, B), a, b)
    bar = 2 if name in ("pow", "hypot", "atan2") else 0
    assert ulp_distance(got, want) <= bar


@pytest.mark.parametrize("dtype", [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64])
def test_integer_arithmetic_bit_exact(xt, gpu, dtype):
    info = np.iinfo(dtype)
    a = rnd((63, 17), dtype, max(info.min, -100), min(info.max, 100), seed=3)
    b = rnd((63, 17), dtype, 1, min(info.max, 50), seed=4)
    f = lambda A, B: (A + B) * A - (A / B) + (A % B) + (A & B) - (A | B) + (A ^ B)
    got, want = run_both(xt, f, a, b)
    assert_bit_exact(got, want)
    # C++ promotion: narrow types compute as int
    if np.dtype(dtype).itemsize < 4:
        assert got.dtype == np.int32


def test_mixed_type_promotion(xt, gpu):
    """2.0 * float_array is double; int + float is float (test/test_xmath_result_type.cpp:227-256)."""
    a = rnd((40, 9), F32, seed=5)
    i = rnd((40, 9), np.int32, -9, 9, seed=6)
    got, want = run_both(xt, lambda A: 2.0 * A, a)
    assert got.dtype == F64
    assert_bit_exact(got, want)
    assert_bit_exact(got, 2.0 * a.astype(F64))
    got, want = run_both(xt, lambda A, I: A + I, a, i)
    assert got.dtype == F32
    assert_bit_exact(got, want)
    got, want = run_both(xt, lambda I: xt.sqrt(xt.abs(I)), i)   # std::sqrt(int) -> double
    assert got.dtype == F64
    assert_bit_exact(got, want)


def test_comparison_where_cast(xt, gpu):
    a, b = rnd((33, 45), F32, seed=7), rnd((33, 45), F32, seed=8)
    got, want = run_both(xt, lambda A, B: xt.where(A > B, A, B * F32(0.5)), a, b)
    assert_bit_exact(got, want)
    assert_bit_exact(got, np.where(a > b, a, b * F32(0.5)))
    got, want = run_both(xt, lambda A, B: (A < B), a, b)
    assert got.dtype == np.bool_
    assert np.array_equal(got, a < b)
    got, want = run_both(xt, lambda A: xt.cast(A * F32(40), xt.I8), a)
    assert got.dtype == np.int8
    assert_bit_exact(got, want)
    got, want = run_both(xt, lambda A, B: xt.clip(A, F32(-1), F32(1)) + xt.fma(A, B, A), a, b)
    assert ulp_distance(got, want) == 0


def test_store_cast_into_other_dtype(xt, gpu):
    """has_assign_conversion: container dtype differs from the expression's (xassign.hpp:613-642)."""
    a = rnd((50, 3), F64, -100, 100)
    got, want = run_both(xt, lambda A: A * 1.5, a, dtype=xt.I32)
    assert got.dtype == np.int32
    assert_bit_exact(got, want)
    got, want = run_both(xt, lambda A: A * 1.5, a, dtype=xt.F32)
    assert_bit_exact(got, want)


def test_views_offsets_negative_steps_and_misalignment(xt, gpu):
    a = rnd((40, 50), F32, seed=9)
    f = lambda A: A[1:, 3:48] * F32(2) + A[:-1, 5:50]       # misaligned rows (offset 3 and 5 floats)
    got, want = run_both(xt, f, a)
    assert_bit_exact(got, want)
    assert_bit_exact(got, a[1:, 3:48] * F32(2) + a[:-1, 5:50])
    g = lambda A: A[::-1, ::-2] + A[:, ::2]                  # negative strides
    got, want = run_both(xt, g, a)
    assert_bit_exact(got, want)
    assert_bit_exact(got, a[::-1, ::-2] + a[:, ::2])
    h = lambda A: A[3, None, :] - A[:, 7, None]              # integer index, newaxis
    got, want = run_both(xt, h, a)
    assert_bit_exact(got, a[3, None, :] - a[:, 7, None])


def test_explicit_broadcast_and_scalar_only(xt, gpu):
    b = rnd((7,), F32)
    got, want = run_both(xt, lambda B: xt.broadcast(B, (3, 5, 7)) * F32(3), b)
    assert_bit_exact(got, want)
    assert got.shape == (3, 5, 7)
    out = xt.DeviceArray.empty((4, 5), xt.F64)
    xt.noalias(out).assign(2.5)                               # xscalar broadcast fill
    assert np.array_equal(out.numpy(), np.full((4, 5), 2.5))


def test_empty_and_zero_dim(xt, gpu):
    e = np.zeros((0, 5), F32)
    got, want = run_both(xt, lambda A: A + F32(1), e)
    assert got.shape == (0, 5)
    s = np.array(3.0)
    got, want = run_both(xt, lambda A: A * 2.0, s)
    assert got.shape == () and float(got) == 6.0


def test_high_rank_generic_kernel(xt, gpu):
    a = rnd((3, 4, 2, 5, 3, 2), F32)
    b = rnd((4, 1, 5, 1, 2), F32, seed=2)
    f = lambda A, B: A.transpose([5, 1, 2, 3, 4, 0]).transpose([5, 1, 2, 3, 4, 0])[:, :, ::-1] * B
    got, want = run_both(xt, f, a, b)
    assert_bit_exact(got, want)
    assert_bit_exact(got, a[:, :, ::-1] * b)


def test_compound_assign(xt, gpu):
    a, b = rnd((20, 30), F64, seed=1), rnd((30,), F64, seed=2)
    d = xt.DeviceArray.from_numpy(a)
    xt.noalias(d).plus_assign(xt.DeviceArray.from_numpy(b))
    assert_bit_exact(d.numpy(), a + b)
    xt.noalias(d).multiplies_assign(3.0)
    assert_bit_exact(d.numpy(), (a + b) * 3.0)


def test_assign_host_streamed(xt, gpu):
    """xtb_assign_host: host-resident operands streamed through the device in chunks."""
    import ctypes as C
    from xtensor_b200 import capi
    shape = (37, 50, 64)
    a, d = rnd(shape, seed=3), rnd(shape, seed=5)
    b = rnd((1, 50, 1), lo=0.5, hi=1.5, seed=4)
    H = xt.HostArray.from_numpy
    ha, hb, hd = H(a), H(b), H(d)
    expr = xt.sin(ha) * hb + F32(2.0) * hd
    lw = xt.lower(expr)
    out = xt.HostArray.empty(shape, xt.F32)
    prog, ops, oop = lw.program(), lw.operands(), out.operand()
    for chunk in (0, 40000, 1 << 30):              # default, many small chunks, a single chunk
        out.owner[:] = 0
        capi.check(gpu.xtb_assign_host(C.byref(prog), C.byref(oop), ops, chunk))
        dev = xt.evaluate(xt.sin(xt.DeviceArray.from_numpy(a)) * xt.DeviceArray.from_numpy(b)
                          + F32(2.0) * xt.DeviceArray.from_numpy(d)).numpy()
        assert_bit_exact(out.numpy(), dev)


def test_sin_cos_f32_own_reduction(xt, gpu):
    """fp32 sin / cos use the backend's own Cody-Waite fast path (one range check per vector) and the
    library routine beyond |x| > 105615: <= 2 ulp of glibc everywhere, specials preserved, and the three
    evaluators (ahead-of-time, interpreter, generic strided) agree bit for bit."""
    rng = np.random.default_rng(21)
    k = np.arange(-4000, 4000, dtype=np.float64)
    near = np.concatenate([(k * np.pi / 2).astype(np.float32), np.nextafter((k * np.pi / 2).astype(np.float32), np.float32(np.inf)),
                           (k * np.pi / 4).astype(np.float32)])
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 105615.0, -105615.0, 105615.01, 1e-30, -1e-30, 1e-45, 3e38, -3e38,
                        1e6, 1e9, 7e4, -9.9e4], np.float32)
    x = np.concatenate([rng.uniform(-np.pi, np.pi, 1 << 18), rng.uniform(-100, 100, 1 << 18), rng.uniform(-105615, 105615, 1 << 18),
                        rng.uniform(-1e7, 1e7, 1 << 16), near, special]).astype(np.float32)
    x = np.resize(x, (x.size + 3) // 4 * 4)          # mixed small / huge / non-finite values inside one 4-vector
    for fn in (xt.sin, xt.cos):
        with np.errstate(invalid="ignore"):
            got, want = run_both(xt, lambda A: fn(A), x)
        assert ulp_distance(got, want) <= 2, fn.__name__
        assert np.array_equal(np.signbit(got[want == 0]), np.signbit(want[want == 0]))
        with interpreter_only():
            interp = xt.evaluate(fn(xt.DeviceArray.from_numpy(x))).numpy()
        assert_bit_exact(interp, got)
        strided = xt.evaluate(fn(xt.DeviceArray.from_numpy(np.repeat(x, 2))[::2])).numpy()      # gather path, scalar evaluation
        assert_bit_exact(strided, got)
