"""ctypes loader of the REAL reference compiled under oracle/_ref (see oracle/ref/Makefile).

TEST INFRASTRUCTURE ONLY (tests/, bench.py cpu_baseline / --impl reference).  The
libraries are prebuilt in the build container (where /root/reference exists) and travel
to the GPU box; nothing here reads /root/reference at run time."""
import ctypes as C
import os
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def available(fast: bool = False) -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libxtref_fast.so" if fast else "libxtref.so"))


def lib(fast: bool = False):
    key = bool(fast)
    if key not in _LIBS:
        path = os.path.join(_HERE, "_ref", "libxtref_fast.so" if fast else "libxtref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = C.CDLL(path)
        L.xtref_last_error.restype = C.c_char_p
        _LIBS[key] = L
    return _LIBS[key]


def _p(a):
    return C.c_void_p(a.ctypes.data)


def _i64(seq):
    return (C.c_int64 * max(len(seq), 1))(*seq)


def _i32(seq):
    return (C.c_int32 * max(len(seq), 1))(*seq)


def cfg1(a, b, fast=False):
    c = np.empty_like(a)
    lib(fast).xtref_cfg1_add_f64(_p(a), _p(b), _p(c), C.c_int64(a.size))
    return c


def cfg2(a, b, d, fast=False):
    c = np.empty_like(a)
    lib(fast).xtref_cfg2_f32(_p(a), _p(b), _p(d), _p(c), *(C.c_int64(s) for s in a.shape))
    return c


def cfg4(a, b, fast=False):
    out = np.empty_like(a)
    lib(fast).xtref_cfg4_f64(_p(a), _p(b), _p(out), C.c_int64(a.shape[0]))
    return out


def cfg5_map(a, m, fast=False):
    out = np.empty_like(a)
    lib(fast).xtref_cfg5_exp_sub_f32(_p(a), _p(m), _p(out), C.c_int64(a.shape[0]), C.c_int64(a.shape[1]))
    return out


def axmby(x, y, fast=False):
    r = np.empty_like(x)
    lib(fast).xtref_axmby_f64(_p(x), _p(y), _p(r), C.c_int64(x.shape[0]), C.c_int64(x.shape[1]))
    return r


def bcast_add(a, b):
    shp = np.broadcast_shapes(a.shape, b.shape)
    out = np.empty(shp, np.float64)
    r = lib().xtref_bcast_add_f64(_p(a), a.ndim, _i64(a.shape), _p(b), b.ndim, _i64(b.shape), _p(out), out.ndim, _i64(shp))
    if r != 0:
        raise RuntimeError(lib().xtref_last_error().decode())
    return out


def unary(name, a):
    out = np.empty_like(a)
    r = lib().xtref_unary(name.encode(), int(a.dtype == np.float64), _p(a), _p(out), C.c_int64(a.size))
    if r != 0:
        raise RuntimeError(lib().xtref_last_error().decode())
    return out


def binary(name, a, b):
    out = np.empty_like(a)
    r = lib().xtref_binary(name.encode(), int(a.dtype == np.float64), _p(a), _p(b), _p(out), C.c_int64(a.size))
    if r != 0:
        raise RuntimeError(lib().xtref_last_error().decode())
    return out


_DT = {np.dtype(np.int8): 1, np.dtype(np.uint8): 2, np.dtype(np.int16): 3, np.dtype(np.uint16): 4, np.dtype(np.int32): 5,
       np.dtype(np.uint32): 6, np.dtype(np.int64): 7, np.dtype(np.uint64): 8, np.dtype(np.float32): 9,
       np.dtype(np.float64): 10}


def int_expr(a, b):
    rt = a.dtype if a.dtype.itemsize >= 4 else np.dtype(np.int32)
    out = np.empty(a.shape, rt)
    r = lib().xtref_int_expr(_DT[a.dtype], _p(a), _p(b), _p(out), C.c_int64(a.size))
    assert r == 0
    return out


def reduce(op, a, axes, keep_dims=False, mode=0, fast=False):
    """op: 0 sum, 1 prod, 2 amax, 3 amin; mode 0 lazy, 1 immediate."""
    axes = list(axes)
    shp = [1 if d in axes else s for d, s in enumerate(a.shape)] if keep_dims else [s for d, s in enumerate(a.shape) if d not in axes]
    out = np.empty(shp, a.dtype)
    r = lib(fast).xtref_reduce(op, _DT[a.dtype], _p(a), a.ndim, _i64(a.shape), len(axes), _i32(axes), int(keep_dims), mode, _p(out))
    if r < 0:
        raise RuntimeError(lib(fast).xtref_last_error().decode())
    return out


def mean_f32(a, axes, as_f32=False):
    shp = [s for d, s in enumerate(a.shape) if d not in axes]
    if as_f32:
        out = np.empty(shp, np.float32)
        lib().xtref_mean_f32_f32(_p(a), a.ndim, _i64(a.shape), len(axes), _i32(axes), _p(out))
    else:
        out = np.empty(shp, np.float64)
        lib().xtref_mean_f32(_p(a), a.ndim, _i64(a.shape), len(axes), _i32(axes), _p(out))
    return out


def variance(a, axes, ddof=0):
    shp = [s for d, s in enumerate(a.shape) if d not in axes]
    out = np.empty(shp, a.dtype)
    if a.dtype == np.float64:
        lib().xtref_variance_f64(_p(a), a.ndim, _i64(a.shape), len(axes), _i32(axes), ddof, _p(out))
    else:
        assert ddof == 0
        lib().xtref_variance_f32_f32(_p(a), a.ndim, _i64(a.shape), len(axes), _i32(axes), _p(out))
    return out


def cumsum(a, axis=None):
    rt = np.dtype(np.int32) if a.dtype == np.int16 else a.dtype
    out = np.empty(a.shape if axis is not None else (a.size,), rt)
    r = lib().xtref_cumsum(_DT[a.dtype], _p(a), a.ndim, _i64(a.shape), -1 if axis is None else axis, _p(out))
    if r < 0:
        raise RuntimeError(lib().xtref_last_error().decode())
    return out


def run_cfg2(sample_rows: int):
    """Timed CPU baseline for bench.py: the reference's own evaluation of cfg2 on the first
    `sample_rows` leading rows.  cfg2 selects xtensor's single-threaded stepper_assigner
    (SURVEY.md Appendix A), so OpenMP does not engage and cores = 1."""
    if not available(fast=True):
        return None
    shape = (sample_rows, 1024, 64)
    a = np.random.default_rng(3).uniform(-np.pi, np.pi, shape).astype(np.float32)
    b = np.random.default_rng(4).uniform(0.5, 1.5, (1, 1024, 1)).astype(np.float32)
    d = np.random.default_rng(5).uniform(-np.pi, np.pi, shape).astype(np.float32)
    cfg2(a[:1], b, d[:1], fast=True)
    t0 = time.perf_counter()
    cfg2(a, b, d, fast=True)
    dt = time.perf_counter() - t0
    nbytes = 3 * a.size * 4 + 1024 * 4
    return {"gbs": nbytes / dt / 1e9, "seconds": dt, "cores": 1,
            "how": "real xtensor 0.27.1 headers (oracle/_ref/libxtref_fast.so: -O3 -march=x86-64-v3 -fopenmp "
                   "-DXTENSOR_USE_OPENMP, xtl stand-in, no xsimd); this expression runs xtensor's single-threaded stepper_assigner"}


def run_context():
    """Timed context numbers for bench.py (reported next to the device numbers, not a target): xtensor's own CPU
    evaluation of the other BASELINE configs on bounded samples, libxtref_fast.so (OpenMP build: the linear and
    strided-loop assigners use all host threads, the stepper assigner and reduce_immediate are single-threaded)."""
    if not available(fast=True):
        return None
    threads = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    rng = np.random.default_rng(11)
    res = {"threads": threads, "library": "oracle/_ref/libxtref_fast.so (real xtensor 0.27.1, -O3 -march=x86-64-v3 -fopenmp -DXTENSOR_USE_OPENMP, no xsimd)"}

    def best(fn, reps=3):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return min(ts)

    n = 1 << 24
    a, b = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    res["cfg1_add_f64_2^24"] = {"GBs": round(3 * n * 8 / best(lambda: cfg1(a, b, fast=True)) / 1e9, 2), "path": "linear_assigner (OpenMP)"}
    x = rng.uniform(-1, 1, (256, 4096, 16)).astype(np.float32)
    for axis in (0, 2):
        dt = best(lambda: reduce(0, x, [axis], mode=1, fast=True), reps=2)
        res[f"cfg3_sum_axis{axis}_sample_256x4096x16"] = {"GBs": round(x.nbytes / dt / 1e9, 2), "path": "reduce_immediate (1 thread)"}
    m = 2048
    a4, b4 = rng.uniform(-1, 1, (m, m)), rng.uniform(-1, 1, (2 * m, m))
    res["cfg4_transpose_view_f64_sample_2048"] = {"GBs": round(3 * m * m * 8 / best(lambda: cfg4(a4, b4, fast=True), reps=2) / 1e9, 2),
                                                   "path": "stepper_assigner (1 thread)"}
    return res


def nanfn(name, a, axes):
    """xt::nansum / nanprod / nanmin / nanmax / nanmean / nanvar / nanstd / count_nonzero / count_nonnan over `axes`
    ("nanmean_t" / "nanvar_t": result type = input type)."""
    a = np.ascontiguousarray(a)
    dt = {np.dtype(np.float32): 9, np.dtype(np.float64): 10}[a.dtype]
    out_shape = tuple(s for d, s in enumerate(a.shape) if d not in axes)
    if name.startswith("count"):
        odt = np.uint64
    elif name in ("nanmean", "nanvar", "nanstd"):
        odt = np.float64
    else:
        odt = a.dtype
    out = np.empty(out_shape, odt)
    r = lib().xtref_nanfn(name.encode(), dt, _p(a), a.ndim, _i64(a.shape), len(axes), _i32(axes), _p(out))
    if r < 0:
        raise RuntimeError(lib().xtref_last_error().decode())
    assert r == out.size
    return out


def count_nonzero_i32(a, axes):
    a = np.ascontiguousarray(a, np.int32)
    out = np.empty(tuple(s for d, s in enumerate(a.shape) if d not in axes), np.uint64)
    r = lib().xtref_count_nonzero_i32(_p(a), a.ndim, _i64(a.shape), len(axes), _i32(axes), _p(out))
    assert r == out.size
    return out


def nan_to_num(a):
    a = np.ascontiguousarray(a)
    out = np.empty_like(a)
    lib().xtref_nan_to_num(int(a.dtype == np.float64), _p(a), _p(out), C.c_int64(a.size))
    return out


def average(a, w, axes):
    """xt::average(a, w[, axes]) on fp64; axes == [] is the whole-array form."""
    a, w = np.ascontiguousarray(a, np.float64), np.ascontiguousarray(w, np.float64)
    out = np.empty(tuple(s for d, s in enumerate(a.shape) if d not in axes) if axes else (), np.float64)
    r = lib().xtref_average_f64(_p(a), a.ndim, _i64(a.shape), _p(w), w.ndim, _i64(w.shape), len(axes), _i32(axes), _p(out))
    if r < 0:
        raise RuntimeError(lib().xtref_last_error().decode())
    assert r == out.size
    return out


_DT = {np.dtype(np.int8): 1, np.dtype(np.uint8): 2, np.dtype(np.int16): 3, np.dtype(np.uint16): 4, np.dtype(np.int32): 5,
       np.dtype(np.int64): 7, np.dtype(np.uint64): 8, np.dtype(np.float32): 9, np.dtype(np.float64): 10}


def argfn(name, a, axis=None):
    """xt::argmin / xt::argmax (misc/xsort.hpp:1237-1295) of the real reference; axis None = flattened."""
    a = np.ascontiguousarray(a)
    shp = () if axis is None else tuple(s for d, s in enumerate(a.shape) if d != (axis % a.ndim))
    out = np.empty(shp, np.uint64)
    r = lib().xtref_argfn(int(name == "argmax"), _DT[a.dtype], _p(a), a.ndim, _i64(a.shape), -100 if axis is None else int(axis), _p(out))
    if r < 0:
        raise RuntimeError(lib().xtref_last_error().decode())
    assert r == max(out.size, 1)
    return out


def minmax(a):
    a = np.ascontiguousarray(a)
    out = np.empty(2, a.dtype)
    r = lib().xtref_minmax(_DT[a.dtype], _p(a), a.ndim, _i64(a.shape), _p(out))
    if r < 0:
        raise RuntimeError(lib().xtref_last_error().decode())
    return out


def normfn(name, a, axis, p=0.0):
    """xt::norm_<name>(a, {axis}[, p]) (reducers/xnorm.hpp); the result dtype is whatever the reference computes."""
    a = np.ascontiguousarray(a)
    shp = tuple(s for d, s in enumerate(a.shape) if d != axis)
    raw = np.empty(int(np.prod(shp, dtype=np.int64)) * 8, np.uint8)
    r = lib().xtref_normfn(name.encode(), _DT[a.dtype], _p(a), a.ndim, _i64(a.shape), int(axis), C.c_double(p), _p(raw))
    if r < 0:
        raise RuntimeError(lib().xtref_last_error().decode())
    n, width = r // 16, r % 16
    assert n == int(np.prod(shp, dtype=np.int64))
    return raw[: n * width], width, shp
