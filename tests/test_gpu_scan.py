"""GPU parity: xtb_scan (cumsum / cumprod) vs the CPU oracle, the reference's golden vectors
and the literal expectations of test/test_xaccumulator.cpp:23-211."""
import json
import os

import numpy as np
import pytest

from util import assert_bit_exact

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "ref_vectors.npz"))
KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))


def both(xt, a, axis, fn="cumsum", dtype=None):
    g = getattr(xt, fn)(xt.DeviceArray.from_numpy(a), axis, dtype).numpy()
    w = getattr(xt, fn)(xt.HostArray.from_numpy(a), axis, dtype).numpy()
    return g, w


def test_kats(xt, gpu):
    k = KATS["accumulator_one_d"]
    r = xt.cumsum(xt.DeviceArray.from_numpy(np.array(k["input_int16"], np.int16))).numpy()
    assert r.dtype == np.int32 and r.tolist() == k["expected_int32"]           # short -> int promotion
    k = KATS["accumulator_four_d"]
    a = xt.DeviceArray.from_numpy(np.arange(36, dtype=np.float64).reshape(k["shape"]))
    assert xt.cumsum(a).numpy().tolist() == k["flat"]
    assert xt.cumsum(a, 0).numpy().reshape(-1).tolist() == k["axis0"]
    assert xt.cumsum(a, 1).numpy().reshape(-1).tolist() == k["axis1"]
    one = np.array([[5.0, 6.0, 7.0]])
    assert np.array_equal(xt.cumsum(xt.DeviceArray.from_numpy(one), 0).numpy(), one)
    assert np.array_equal(xt.cumsum(xt.transpose(xt.DeviceArray.from_numpy(one)), 1).numpy(), one.T)


@pytest.mark.parametrize("tag", ["f32", "f64", "i32", "i16"])
@pytest.mark.parametrize("axis", [None, 0, 1, 2])
def test_golden(xt, gpu, tag, axis):
    got = xt.cumsum(xt.DeviceArray.from_numpy(G[f"cumsum_{tag}_in"]), axis).numpy()
    want = G[f"cumsum_{tag}_{'flat' if axis is None else axis}"]
    if tag[0] == "i" or axis in (0, 1):
        assert_bit_exact(got, want)          # strided-axis scans walk the axis in the reference's order
    else:
        assert np.allclose(got, want, rtol=1e-6 if tag == "f32" else 1e-12, atol=1e-6 if tag == "f32" else 1e-12)


@pytest.mark.parametrize("shape,axis", [((1,), 0), ((5,), None), ((2048,), 0), ((2049,), 0), ((100000,), None),
                                        ((3000000,), 0), ((37, 5000), 1), ((5000, 37), 0), ((5000, 37), 1),
                                        ((13, 40, 50), 1), ((13, 40, 50), None), ((4, 3, 2, 5, 6), 2)])
@pytest.mark.parametrize("dtype", [np.int32, np.int64, np.float32, np.float64, np.uint8])
def test_shapes_integer_valued(xt, gpu, shape, axis, dtype):
    """Integer-valued data: every association is exact, so the tile / look-back logic is bit-exact."""
    lo, hi = (0, 3) if dtype == np.uint8 else (-3, 3)
    a = np.random.default_rng(5).integers(lo, hi + 1, shape).astype(dtype)
    g, w = both(xt, a, axis)
    assert_bit_exact(g, w)
    ref = np.cumsum(a, axis=axis, dtype=g.dtype)
    assert_bit_exact(g, ref.astype(g.dtype))


def test_random_fp_tolerance_and_determinism(xt, gpu):
    a = np.random.default_rng(6).uniform(-1, 1, 1 << 21).astype(np.float32)
    g1, w = both(xt, a, 0)
    g2, _ = both(xt, a, 0)
    assert_bit_exact(g1, g2)                                       # run-to-run deterministic
    scale = np.cumsum(np.abs(a), dtype=np.float64)
    assert np.all(np.abs(g1.astype(np.float64) - w.astype(np.float64)) <= 1e-6 * np.maximum(scale, 1.0))
    exact = np.cumsum(a, dtype=np.float64)
    assert np.abs(g1 - exact).max() <= np.abs(w - exact).max() * 1.5 + 1e-3


def test_cumprod_views_and_errors(xt, gpu):
    a = np.random.default_rng(7).integers(1, 3, (6, 7)).astype(np.int64)
    g, w = both(xt, a, 1, "cumprod")
    assert_bit_exact(g, w)
    assert_bit_exact(g, np.cumprod(a, axis=1))
    g = xt.cumsum(xt.transpose(xt.DeviceArray.from_numpy(a)), 0).numpy()   # strided (transposed) input
    assert_bit_exact(g, np.cumsum(a.T, axis=0))
    g = xt.cumsum(xt.DeviceArray.from_numpy(a)[::2, 1:6], 1).numpy()
    assert_bit_exact(g, np.cumsum(a[::2, 1:6], axis=1))
    g = xt.cumsum(xt.DeviceArray.from_numpy(a.astype(np.float32)), 1, dtype=xt.F64).numpy()  # cumsum<double>(float)
    assert g.dtype == np.float64
    with pytest.raises(RuntimeError, match="Axis larger"):
        xt.cumsum(xt.DeviceArray.from_numpy(a), 2)
    e = xt.cumsum(xt.DeviceArray.from_numpy(np.zeros((0, 3), np.float32)), 0).numpy()
    assert e.shape == (0, 3)
