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


# ---- the tiled kernels: staged super-tiles (long rows), packed short rows, column tiles -------------
BIG = [
    # contiguous axis: several super-tiles per row -> two-level look-back (block totals from 256 tiles on)
    ((16384 * 5 + 4,), None), ((16384 * 300 + 8,), None), ((16384 * 300 + 5,), None), ((3, 16384 * 40), 1), ((5, 100003), 1),
    # rows of one super-tile, ragged rows, short rows packed per warp / one warp per row
    ((300, 8192), 1), ((300, 5000), 1), ((4096, 64), 1), ((4096, 4), 1), ((1000, 24), 1), ((513, 512), 1), ((77, 300), 1),
    # strided axis: column tiles (wide, narrow with row groups, ragged, tree of 2 and 3 levels, outer dims)
    ((3000, 512), 0), ((3000, 300), 0), ((20000, 64), 0), ((20000, 3), 0), ((70000, 8), 0), ((2, 700, 260), 1),
    ((5, 130, 7, 9), 1), ((128, 1000), 0),
]


@pytest.mark.parametrize("shape,axis", BIG)
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64, np.int16])
def test_tiled_kernels_integer_valued(xt, gpu, shape, axis, dtype):
    a = np.random.default_rng(11).integers(-3, 4, shape).astype(dtype)
    got = xt.cumsum(xt.DeviceArray.from_numpy(a), axis).numpy()
    want = np.cumsum(a, axis=axis, dtype=got.dtype).reshape(got.shape)
    assert_bit_exact(got, want)


def test_tiled_kernels_views_and_conversions(xt, gpu):
    """Inputs the bulk-copy path cannot take (strided, unaligned, other dtype) go through registers."""
    rng = np.random.default_rng(12)
    a = rng.integers(-3, 4, (600, 1030)).astype(np.float32)
    d = xt.DeviceArray.from_numpy(a)
    assert_bit_exact(xt.cumsum(d[:, 1:1027], 0).numpy(), np.cumsum(a[:, 1:1027], axis=0))       # unaligned columns
    assert_bit_exact(xt.cumsum(d[::2, :], 0).numpy(), np.cumsum(a[::2, :], axis=0))              # strided rows
    assert_bit_exact(xt.cumsum(xt.transpose(d), 1).numpy(), np.cumsum(a.T, axis=1))              # scan axis strided
    assert_bit_exact(xt.cumsum(xt.transpose(d), 0).numpy(), np.cumsum(a.T, axis=0))
    assert_bit_exact(xt.cumsum(d, 0, dtype=xt.F64).numpy(), np.cumsum(a, axis=0, dtype=np.float64))
    b = rng.integers(-3, 4, (3, 70001)).astype(np.int16)
    assert_bit_exact(xt.cumsum(xt.DeviceArray.from_numpy(b), 1).numpy(), np.cumsum(b, axis=1, dtype=np.int32))
    assert_bit_exact(xt.cumsum(xt.DeviceArray.from_numpy(b)[:, 3:], 1).numpy(), np.cumsum(b[:, 3:], axis=1, dtype=np.int32))
    c = rng.integers(1, 3, (40, 300)).astype(np.float64) * 0.5 + 0.5                              # {1, 1.5}: cumprod exact for a while
    c[:, 20:] = 1.0
    assert_bit_exact(xt.cumprod(xt.DeviceArray.from_numpy(c), 0).numpy(), np.cumprod(c, axis=0))


def test_tiled_kernels_fp_tolerance_and_determinism(xt, gpu):
    rng = np.random.default_rng(13)
    for shape, axis in [((1 << 22,), None), ((4096, 1024), 0), ((64, 65536), 1)]:
        a = rng.uniform(-1, 1, shape).astype(np.float32)
        d = xt.DeviceArray.from_numpy(a)
        g1 = xt.cumsum(d, axis).numpy()
        g2 = xt.cumsum(d, axis).numpy()
        assert_bit_exact(g1, g2)                                   # fixed look-back tree: no timing dependence
        ref = np.cumsum(a.astype(np.float64), axis=axis).reshape(g1.shape)
        scale = np.cumsum(np.abs(a).astype(np.float64), axis=axis).reshape(g1.shape)
        assert np.all(np.abs(g1 - ref) <= 1e-6 * np.maximum(scale, 1.0))


# ---- round 2 kernels: reduce-ahead scan (k_scan_ahead) and column walkers (k_scan_colwalk) --------------
AHEAD = [
    ((16384 * 4 + 1,), None), ((4096 * 33 + 7,), None), ((4096 * 1030 + 12,), None),     # 1-, 2- and 3-level aggregate trees
    ((1 << 24,), None), ((3, 4096 * 70 + 4), 1), ((2, 4096 * 1100), 1), ((5, 16385 * 4), 1),
]


@pytest.mark.parametrize("shape,axis", AHEAD)
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
def test_reduce_ahead_scan(xt, gpu, shape, axis, dtype):
    from util import last_kernel
    a = np.random.default_rng(21).integers(-3, 4, shape).astype(dtype)
    got = xt.cumsum(xt.DeviceArray.from_numpy(a), axis).numpy()
    if np.prod(shape[-1:]) * a.itemsize > 64 * 1024:
        assert "k_scan_ahead" in last_kernel(), last_kernel()
    assert_bit_exact(got, np.cumsum(a, axis=axis, dtype=got.dtype).reshape(got.shape))


def test_reduce_ahead_scan_views_and_prod(xt, gpu):
    rng = np.random.default_rng(22)
    a = rng.integers(-3, 4, (3, 200003)).astype(np.int16)                       # converted input, unaligned rows
    assert_bit_exact(xt.cumsum(xt.DeviceArray.from_numpy(a), 1).numpy(), np.cumsum(a, axis=1, dtype=np.int32))
    b = rng.integers(-3, 4, (2, 300000)).astype(np.float32)
    d = xt.DeviceArray.from_numpy(b)
    assert_bit_exact(xt.cumsum(d[:, 5:], 1).numpy(), np.cumsum(b[:, 5:], axis=1))                # unaligned view
    assert_bit_exact(xt.cumsum(d[:, ::2], 1).numpy(), np.cumsum(b[:, ::2], axis=1))              # strided view
    c = np.ones(1 << 20, np.float64)
    c[::4099] = 2.0
    c[1::4099] = 0.5
    assert_bit_exact(xt.cumprod(xt.DeviceArray.from_numpy(c)).numpy(), np.cumprod(c))


WALK = [((300, 4736), 0), ((1000, 9472 + 4), 0), ((2, 257, 8192), 1), ((513, 148 * 128), 0), ((4097, 5000), 0)]


@pytest.mark.parametrize("shape,axis", WALK)
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
def test_column_walkers(xt, gpu, shape, axis, dtype):
    from util import last_kernel
    rng = np.random.default_rng(23)
    a = rng.integers(-3, 4, shape).astype(dtype)
    got = xt.cumsum(xt.DeviceArray.from_numpy(a), axis).numpy()
    assert "k_scan_colwalk" in last_kernel(), last_kernel()
    assert_bit_exact(got, np.cumsum(a, axis=axis, dtype=got.dtype))
    if dtype in (np.float32, np.float64):
        # walkers accumulate every column in the reference's order: random fp data is bit-exact against the oracle too
        b = rng.uniform(-1, 1, (shape[0] if len(shape) == 2 else shape[0], *shape[1:])).astype(dtype)
        g, w = both(xt, b, axis)
        assert_bit_exact(g, w)
        b[0] = -0.0                                            # out[0] = in[0]: a leading -0.0 survives
        g, w = both(xt, b, axis)
        assert_bit_exact(g, w)


@pytest.mark.parametrize("shape,axis", [((300, 4001), 0), ((129, 5003), 0), ((1000, 3999), 0), ((2, 257, 3001), 1),
                                        ((131, 37, 111), 0), ((4097, 4095), 0)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
def test_unaligned_strided_axis_walkers(xt, gpu, shape, axis, dtype):
    """Odd row pitch (no tensor map possible): the cp.async walkers accumulate every column in the reference's order,
    so random floating-point data is bit-exact against the oracle, cumprod included."""
    from util import last_kernel
    rng = np.random.default_rng(9)
    a = (rng.uniform(-1, 1, shape) if np.issubdtype(dtype, np.floating) else rng.integers(-3, 4, shape)).astype(dtype)
    g, w = both(xt, a, axis)
    assert "colwalk_plain" in last_kernel(), last_kernel()
    assert_bit_exact(g, w)
    b = (rng.uniform(0.9, 1.1, shape) if np.issubdtype(dtype, np.floating) else rng.integers(1, 2, shape)).astype(dtype)
    g, w = both(xt, b, axis, fn="cumprod")
    assert_bit_exact(g, w)
    # a view whose base is off by one element (aligned pitch, unaligned base)
    if len(shape) == 2:
        c = (rng.uniform(-1, 1, (shape[0], 4100)) if np.issubdtype(dtype, np.floating) else rng.integers(-3, 4, (shape[0], 4100))).astype(dtype)
        gv = xt.cumsum(xt.view(xt.DeviceArray.from_numpy(c), slice(None), slice(1, None)), 0).numpy()
        wv = xt.cumsum(xt.view(xt.HostArray.from_numpy(c), slice(None), slice(1, None)), 0).numpy()
        assert_bit_exact(gv, wv)
