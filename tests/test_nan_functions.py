"""nan-aware reducers, counts and nan_to_num (core/xmath.hpp:2307-2860; the reference's own tests:
test/test_xnan_functions.cpp) against tests/golden/ref_vectors_nan.npz = outputs of the REAL reference
(generator: tests/golden/make_golden_nan.py).  The inputs are small integers with NaN / +-inf sprinkled
in (and all-NaN lanes), so sums and products are exact in any order: nansum / nanprod / nanmin / nanmax /
counts / nan_to_num must be bit-exact on the oracle AND on the device; nanmean / nanvar / nanstd divide
and take roots of exact sums and are held to the reduction tolerance of the north star (1e-6 relative
for fp32 results, 1e-12 for fp64).

CPU (not gpu): the oracle evaluates the lowered programs.  GPU: the same programs through libxtb200."""
import os

import numpy as np
import pytest

from util import assert_bit_exact

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors_nan.npz"))
AXES = [[0], [1], [2], [0, 1], [1, 2], [0, 2], [0, 1, 2]]
EXACT = ["nansum", "nanprod", "nanmin", "nanmax", "count_nonzero", "count_nonnan"]
CLOSE = ["nanmean", "nanvar", "nanstd", "nanmean_t", "nanvar_t"]


def _call(xt, name, a, ax):
    if name.endswith("_t"):
        return getattr(xt, name[:-2])(a, ax, dtype=a.dtype)
    return getattr(xt, name)(a, ax)


def _check(xt, make, name, tag, ax):
    a = make(G[f"in_{tag}"])
    want = G[f"{name}_{tag}_ax{''.join(map(str, ax))}"]
    got = xt.evaluate(_call(xt, name, a, ax)).numpy()
    if name in EXACT:
        assert_bit_exact(got, want)
    else:
        assert got.dtype == want.dtype and got.shape == want.shape
        rtol = 1e-6 if want.dtype == np.float32 else 1e-12
        assert np.array_equal(np.isnan(got), np.isnan(want))         # all-NaN lanes: 0 / 0
        assert np.allclose(got, want, rtol=rtol, atol=0, equal_nan=True)


def _check_misc(xt, make):
    for tag in ("f32", "f64"):
        assert_bit_exact(xt.evaluate(xt.nan_to_num(make(G[f"n2n_{tag}_in"]))).numpy(), G[f"n2n_{tag}_out"])
    for ax in AXES:
        got = xt.evaluate(xt.count_nonzero(make(G["cnz_i32_in"]), ax)).numpy()
        assert_bit_exact(got, G[f"cnz_i32_ax{''.join(map(str, ax))}"])
    # whole-array forms (no axes argument) and keep_dims
    a = G["in_f64"]
    assert_bit_exact(xt.evaluate(xt.nansum(make(a))).numpy(), G["nansum_f64_ax012"])
    assert_bit_exact(xt.evaluate(xt.count_nonnan(make(a))).numpy(), G["count_nonnan_f64_ax012"])
    kd = xt.evaluate(xt.nansum(make(a), [1], keep_dims=True)).numpy()
    assert kd.shape == (6, 1, 7)
    assert_bit_exact(kd.reshape(6, 7), G["nansum_f64_ax1"])


def _check_reference_kats(xt, make):
    """Literal expectations of the reference's own tests (test/test_xnan_functions.cpp:38-120, 130-215)."""
    nan, inf = np.nan, np.inf
    a = np.array([[0, 1, 2, 3], [nan, nan, nan, nan], [3, nan, 1, nan]])
    assert int(xt.evaluate(xt.count_nonnan(make(a))).numpy()) == 6                        # :60-64
    assert np.array_equal(xt.evaluate(xt.count_nonnan(make(a), [0])).numpy(), np.array([2, 1, 2, 1], np.uint64))
    assert np.array_equal(xt.evaluate(xt.count_nonnan(make(a), [1])).numpy(), np.array([4, 0, 2], np.uint64))
    b = np.array([[nan, nan, 123], [0.5123, -inf, inf]])
    fi = np.finfo(np.float64)
    assert np.array_equal(xt.evaluate(xt.nan_to_num(make(b))).numpy(), np.array([[0, 0, 123], [0.5123, fi.min, fi.max]]))   # :76-99
    aN = np.array([[nan, nan, 123, 3], [1, 2, nan, 3], [1, 1, nan, 3]])
    aR = np.where(np.isnan(aN), 0.0, aN)
    aP = np.where(np.isnan(aN), 1.0, aN)
    aI = np.where(np.isnan(aN), fi.max, aN)
    aA = np.where(np.isnan(aN), fi.tiny, aN)
    for ax in (None, [0], [1]):
        axis = None if ax is None else ax[0]
        assert np.array_equal(xt.evaluate(xt.nansum(make(aN), ax)).numpy(), aR.sum(axis=axis))      # nansum(aN) == sum(aR)
        assert np.array_equal(xt.evaluate(xt.nanprod(make(aN), ax)).numpy(), aP.prod(axis=axis))    # nanprod(aN) == prod(aP)
        assert np.array_equal(xt.evaluate(xt.nanmin(make(aN), ax)).numpy(), aI.min(axis=axis))      # nanmin(aN) == amin(aI)
        assert np.array_equal(xt.evaluate(xt.nanmax(make(aN), ax)).numpy(), aA.max(axis=axis))      # nanmax(aN) == amax(aA)
    xN = np.array([[[nan, nan], [1, 2]], [[3, nan], [nan, 5]]])
    for i in range(3):   # NAN_SENSITIVE_EQ: NaN exactly where no value exists along the axis
        got = xt.evaluate(xt.nanmin(make(xN), [i])).numpy()
        with np.errstate(all="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                want = np.nanmin(xN, axis=i)
        assert np.array_equal(got, want, equal_nan=True)


def _check_average(xt, make):
    """xt::average (core/xmath.hpp:1925-2010) against the real reference; sums of small integers are exact, the one
    division is the same IEEE operation."""
    a = G["avg_in"]
    for axis in range(3):
        got = xt.evaluate(xt.average(make(a), make(G[f"avg_w1_ax{axis}"]), [axis])).numpy()
        assert_bit_exact(got, G[f"avg_out1_ax{axis}"])
    for ax in ([0], [1, 2], [0, 1, 2]):
        got = xt.evaluate(xt.average(make(a), make(G["avg_wfull"]), ax)).numpy()
        assert_bit_exact(got, G[f"avg_outfull_ax{''.join(map(str, ax))}"])
    assert_bit_exact(xt.evaluate(xt.average(make(a), make(G["avg_wfull"]))).numpy(), G["avg_outfull_all"])
    assert_bit_exact(xt.evaluate(xt.average(make(a))).numpy(), np.asarray(a.mean()))
    with pytest.raises(RuntimeError, match="same shape as expression at axes"):
        xt.average(make(a), make(np.ones(3)), [1])
    with pytest.raises(RuntimeError, match="same shape as expression"):
        xt.average(make(a), make(np.ones((4, 6, 4))), [0])


@pytest.fixture(scope="module")
def H(xt):
    return xt.HostArray.from_numpy


@pytest.fixture(scope="module")
def D(xt, gpu):
    return xt.DeviceArray.from_numpy


@pytest.mark.parametrize("ax", AXES, ids=lambda a: "ax" + "".join(map(str, a)))
@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("name", EXACT + CLOSE)
def test_oracle_nan_reducers(xt, H, name, tag, ax):
    _check(xt, H, name, tag, ax)


def test_oracle_nan_misc(xt, H):
    _check_misc(xt, H)


def test_oracle_reference_kats(xt, H):
    _check_reference_kats(xt, H)


def test_oracle_average(xt, H):
    _check_average(xt, H)


@pytest.mark.gpu
@pytest.mark.parametrize("ax", AXES, ids=lambda a: "ax" + "".join(map(str, a)))
@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("name", EXACT + CLOSE)
def test_gpu_nan_reducers(xt, D, name, tag, ax):
    _check(xt, D, name, tag, ax)


@pytest.mark.gpu
def test_gpu_nan_misc(xt, D):
    _check_misc(xt, D)


@pytest.mark.gpu
def test_gpu_reference_kats(xt, D):
    _check_reference_kats(xt, D)


@pytest.mark.gpu
def test_gpu_average(xt, D):
    _check_average(xt, D)
