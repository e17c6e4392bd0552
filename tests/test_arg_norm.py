"""Index-carrying and lambda-functor reducers against outputs of the REAL reference
(tests/golden/ref_vectors_arg.npz, generator tests/golden/make_golden_arg.py):
  * argmin / argmax  (misc/xsort.hpp:1150-1300; the reference's own cases test/test_xsort.cpp:217-281) over every
    axis and flattened, every dtype, inputs full of ties (first index wins) and NaNs (a NaN never wins unless it
    is element 0 of its lane) -- bit-exact;
  * minmax           (core/xmath.hpp:2195-2228) -- bit-exact;
  * the norms        (reducers/xnorm.hpp:369-620): l0 / l1 / sq / linf on small integers are exact in any order
    -> bit-exact incl. the result dtype; l2 / lp_to_p / lp within the reduction tolerance (1e-12, fp64 results).

CPU (not gpu): the oracle's sequential restatement.  GPU: libxtb200 (xtb_argreduce = two fused map-reduce launches;
nan_min / nan_max as native merges of the reduction kernels)."""
import os

import numpy as np
import pytest

from util import assert_bit_exact

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors_arg.npz"))
ARG_KEYS = ["i8", "u8", "i16", "i32", "i64", "u64", "f32", "f64", "f32_nan", "f64_nan"]
NORM_TAGS = ["i32", "u16", "f32", "f64"]
NORMS = [("l0", 0.0), ("l1", 0.0), ("sq", 0.0), ("l2", 0.0), ("linf", 0.0), ("lp_to_p", 3.0), ("lp", 1.5)]
EXACT_NORMS = {"l0", "l1", "sq", "linf"}
WIDTH_DT = {("l0", 8): np.uint64}


def _check_arg(xt, make, key):
    a = make(G[f"arg_in_{key}"])
    for fn in ("argmin", "argmax"):
        f = getattr(xt, fn)
        got = f(a).numpy()
        assert got.dtype == np.uint64 and got.shape == ()
        assert_bit_exact(got, G[f"{fn}_{key}_flat"])
        for ax in range(3):
            assert_bit_exact(f(a, ax).numpy(), G[f"{fn}_{key}_ax{ax}"])
        assert_bit_exact(f(a, -1).numpy(), G[f"{fn}_{key}_ax2"])            # normalize_axis
    # a strided view: flattened traversal of the view (eval() gives a dense copy), and along an axis
    v = a[1:6:2, :, 0:8:3]
    hv = G[f"arg_in_{key}"][1:6:2, :, 0:8:3]
    ref = _seq_arg(hv.reshape(-1), True)
    assert int(xt.argmin(v).numpy()) == ref
    assert np.array_equal(xt.argmax(v, 1).numpy(), np.apply_along_axis(lambda l: _seq_arg(l, False), 1, hv).astype(np.uint64))


def _seq_arg(lane, is_min):
    """The reference's loop (misc/xsort.hpp:1193-1207): strict comparison, first index, NaN never replaces."""
    best, bi = lane[0], 0
    for i in range(1, len(lane)):
        if (lane[i] < best) if is_min else (lane[i] > best):
            best, bi = lane[i], i
    return bi


def _check_kats(xt, make):
    a = make(G["kat_a"])                                                     # test/test_xsort.cpp:219-240, 248-256
    assert int(xt.argmin(a).numpy()) == 2 and int(xt.argmax(a).numpy()) == 0
    assert np.array_equal(xt.argmin(a, 0).numpy(), [1, 0, 0]) and np.array_equal(xt.argmin(a, 1).numpy(), [2, 0])
    assert np.array_equal(xt.argmax(a, 0).numpy(), [0, 1, 1]) and np.array_equal(xt.argmax(a, 1).numpy(), [0, 0])
    for fn in ("argmin", "argmax"):
        assert_bit_exact(getattr(xt, fn)(a).numpy(), G[f"kat_{fn}_flat"])
        assert_bit_exact(getattr(xt, fn)(a, 0).numpy(), G[f"kat_{fn}_ax0"])
        assert_bit_exact(getattr(xt, fn)(a, 1).numpy(), G[f"kat_{fn}_ax1"])
    b = make(np.array([1, 3, 4, -100], np.float64))                          # :220, 227-228
    assert int(xt.argmin(b).numpy()) == 3 and int(xt.argmin(b, 0).numpy()) == 3
    c = make(np.array([[[1, 2, 3, 4]], [[4, 3, 2, 1]]], np.int32))           # :262-270
    assert np.array_equal(xt.argmax(c, 2).numpy(), [[3], [0]])
    assert np.array_equal(xt.argmax(c, 0).numpy(), [[1, 1, 0, 0]])
    assert np.array_equal(xt.argmax(c, 1).numpy(), [[0, 0, 0, 0], [0, 0, 0, 0]])
    ya = make(np.array([1, 0, 3, 2, 2], np.float64))                         # :241-243, 272-274
    assert int(xt.argmin(ya).numpy()) == 1 and int(xt.argmax(ya, 0).numpy()) == 2
    assert int(xt.argmax(make(np.array([0, 1, 0], np.float64))).numpy()) == 1    # xtensor#2568
    with pytest.raises(RuntimeError):
        xt.argmin(a, 2)


def _check_minmax(xt, make):
    for tag in ("i16", "i32", "f32", "f64", "f32_nan"):
        got = xt.minmax(make(G[f"minmax_in_{tag}"])).numpy()
        assert_bit_exact(got, G[f"minmax_out_{tag}"])
    x = G["minmax_in_f64"]
    got = xt.minmax(make(x)[3:30:2, 5:]).numpy()
    assert np.array_equal(got, [x[3:30:2, 5:].min(), x[3:30:2, 5:].max()])


def _check_norms(xt, make, tag):
    a = make(G[f"norm_in_{tag}"])
    for name, p in NORMS:
        f = getattr(xt, "norm_" + name)
        for ax in range(3):
            raw, meta = G[f"norm_{name}_{tag}_ax{ax}"], G[f"norm_{name}_{tag}_ax{ax}_meta"]
            width, shp = int(meta[0]), tuple(int(s) for s in meta[1:])
            got = xt.evaluate(f(a, p, [ax]) if name.startswith("lp") else f(a, [ax])).numpy()
            assert got.shape == shp and got.dtype.itemsize == width, (name, tag, got.dtype, width)
            want = raw.view(got.dtype).reshape(shp)
            if name in EXACT_NORMS:
                assert_bit_exact(got, want)
            else:
                # fp32 operands of the lp norms go through powf (real_promote_type_t<float> = float): every term is
                # within 2 ulp(fp32) of glibc's, the fp64 sum of them within 3e-7; everything else computes in fp64
                rtol = 3e-7 if (tag == "f32" and name.startswith("lp")) else 1e-12
                assert got.dtype == np.float64 and np.allclose(got, want, rtol=rtol, atol=0)
    # every axis at once == the whole-array form
    assert_bit_exact(xt.evaluate(xt.norm_l1(a)).numpy(), xt.evaluate(xt.norm_l1(a, [0, 1, 2])).numpy())


# ---- CPU: the oracle ---------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def H(xt):
    return xt.HostArray.from_numpy


@pytest.mark.parametrize("key", ARG_KEYS)
def test_oracle_arg(xt, H, key): _check_arg(xt, H, key)
def test_oracle_arg_kats(xt, H): _check_kats(xt, H)
def test_oracle_minmax(xt, H): _check_minmax(xt, H)
@pytest.mark.parametrize("tag", NORM_TAGS)
def test_oracle_norms(xt, H, tag): _check_norms(xt, H, tag)


# ---- GPU ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def D(xt, gpu):
    return xt.DeviceArray.from_numpy


@pytest.mark.gpu
@pytest.mark.parametrize("key", ARG_KEYS)
def test_gpu_arg(xt, D, key): _check_arg(xt, D, key)


@pytest.mark.gpu
def test_gpu_arg_kats(xt, D): _check_kats(xt, D)


@pytest.mark.gpu
def test_gpu_minmax(xt, D): _check_minmax(xt, D)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", NORM_TAGS)
def test_gpu_norms(xt, D, tag): _check_norms(xt, D, tag)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,axis", [((1 << 22,), None), ((4096, 1031), 0), ((4096, 1031), 1), ((37, 211, 129), 1),
                                        ((3, 1 << 20), 1), ((1 << 20, 3), 0)])
@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32, np.uint8])
def test_gpu_arg_large(xt, D, shape, axis, dt):
    """Sizes that take the split / merged reduction kernels: numpy's argmin / argmax have the reference's semantics
    (first extreme) on NaN-free data."""
    rng = np.random.default_rng(5)
    a = rng.integers(0, 200, shape).astype(dt)
    da = D(a)
    assert np.array_equal(xt.argmin(da, axis).numpy(), np.argmin(a, axis=axis).astype(np.uint64))
    assert np.array_equal(xt.argmax(da, axis).numpy(), np.argmax(a, axis=axis).astype(np.uint64))


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["i8", "u8", "i16", "i32", "f32", "f32_nan"])
def test_gpu_arg_both_formulations(xt, D, gpu, key):
    """<= 32-bit element types take the one-pass packed-key reduction; the two-pass formulation (what 64-bit types use)
    must give the same indices."""
    gpu.xtb_set_option(b"arg_two_pass", 1)
    try:
        _check_arg(xt, D, key)
    finally:
        gpu.xtb_set_option(b"arg_two_pass", 0)


@pytest.mark.gpu
def test_gpu_arg_nan_large(xt, D):
    """NaN rules at sizes that split: a NaN in front of a lane pins index 0, other NaNs never win, all-NaN lanes give 0."""
    rng = np.random.default_rng(8)
    a = rng.integers(-50, 50, (3000, 700)).astype(np.float32)
    a[rng.random(a.shape) < 0.2] = np.nan
    a[0, ::3] = np.nan
    a[:, 5] = np.nan
    da = D(a)
    for is_min, f in ((True, xt.argmin), (False, xt.argmax)):
        want0 = np.array([_seq_arg(a[:, j], is_min) for j in range(a.shape[1])], np.uint64)
        assert np.array_equal(f(da, 0).numpy(), want0)
        want1 = np.array([_seq_arg(a[i, :], is_min) for i in range(0, a.shape[0], 7)], np.uint64)
        assert np.array_equal(f(da, 1).numpy()[::7], want1)


@pytest.mark.gpu
def test_gpu_nanmin_large(xt, D):
    """nan_min / nan_max as native merges through the split + merge kernels (outer) and the warp / block kernels (inner)."""
    rng = np.random.default_rng(6)
    a = rng.uniform(-1, 1, (8192, 1536)).astype(np.float32)
    a[rng.random(a.shape) < 0.1] = np.nan
    a[:, 7] = np.nan
    a[11, :] = np.nan
    da = D(a)
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for ax in (0, 1):
                assert_bit_exact(xt.evaluate(xt.nanmin(da, [ax])).numpy(), np.nanmin(a, axis=ax))
                assert_bit_exact(xt.evaluate(xt.nanmax(da, [ax])).numpy(), np.nanmax(a, axis=ax))
            assert_bit_exact(xt.evaluate(xt.nanmax(da)).numpy(), np.float32(np.nanmax(a)))
