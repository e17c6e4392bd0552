"""Time xtb_reduce on a set of shapes / axes / dtypes (CUDA events) and check each against numpy.
usage: python tools/reduce_bench.py [filter ...]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xtensor_b200 import capi  # noqa: E402
from xtensor_b200 import expr as xt  # noqa: E402

lib = capi.lib()
capi.check(lib.xtb_init(0))


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = C.c_void_p(), C.c_void_p()
    capi.check(lib.xtb_event_create(C.byref(e0)))
    capi.check(lib.xtb_event_create(C.byref(e1)))
    capi.check(lib.xtb_sync())
    capi.check(lib.xtb_event_record(e0))
    for _ in range(iters):
        fn()
    capi.check(lib.xtb_event_record(e1))
    ms = C.c_float()
    capi.check(lib.xtb_event_elapsed_ms(e0, e1, C.byref(ms)))
    return ms.value / iters


CASES = [
    ((8192, 8192), [0], np.float32, "sum"), ((8192, 8192), [1], np.float32, "sum"), ((8192, 8192), [0, 1], np.float32, "sum"),
    ((8192, 8192), [0], np.float64, "sum"), ((8192, 8192), [1], np.float64, "sum"),
    ((1 << 26,), [0], np.float32, "sum"), ((1 << 26,), [0], np.float32, "amax"),
    ((4096, 4096, 16), [1], np.float32, "sum"), ((4096, 4096, 16), [0, 2], np.float32, "sum"),
    ((4096, 4096, 16), [1, 2], np.float32, "sum"), ((4096, 4096, 16), [0, 1], np.float32, "sum"),
    ((64, 1 << 20), [0], np.float32, "sum"), ((64, 1 << 20), [1], np.float32, "sum"),
    ((1 << 20, 64), [0], np.float32, "sum"), ((1 << 20, 64), [1], np.float32, "sum"),
    ((1 << 18, 256), [0], np.float32, "sum"), ((1 << 16, 1024), [0], np.float32, "amax"), ((1 << 14, 4096), [0], np.float32, "mean"),
    ((8192, 8192), [0], np.uint8, "sum"), ((8192, 8192), [1], np.uint8, "sum"), ((8192, 8192), [0], np.int32, "amax"),
    ((8192, 8192), [1], np.int64, "sum"), ((8192, 8192), [0], np.int64, "sum"), ((8192, 8192), [0], np.int16, "sum"),
    ((8191, 8190), [0], np.float32, "sum"), ((8191, 8190), [1], np.float32, "sum"), ((4095, 4097, 15), [0], np.float32, "sum"),
]
only = sys.argv[1:]
for shape, axes, dt, op in CASES:
    tag = f"{'x'.join(map(str, shape))}:{''.join(map(str, axes))}:{np.dtype(dt).name}:{op}"
    if only and not any(o in tag for o in only):
        continue
    rng = np.random.default_rng(1)
    n = int(np.prod(shape))
    a = (rng.integers(-3, 4, n) if dt != np.uint8 else rng.integers(0, 4, n)).astype(dt).reshape(shape)
    d = xt.DeviceArray.from_numpy(a)
    r = getattr(xt, op)(d, axes)
    res = xt.evaluate(r)
    if op == "mean":
        f = lambda: xt.assign(res, getattr(xt, op)(d, axes))
    else:
        f = lambda: xt._run_reducer(r, xt.DeviceArray, out=res)
    ms = timed(f)
    got = res.numpy()
    want = {"sum": np.sum, "amax": np.max, "mean": np.mean}[op](a.astype(np.float64) if op == "mean" else a, axis=tuple(axes),
                                                               **({"dtype": got.dtype} if op == "sum" else {}))
    ok = np.array_equal(got, np.asarray(want).astype(got.dtype).reshape(got.shape))
    nbytes = n * a.itemsize + got.size * got.itemsize
    print(f"{tag:36s} {ms:8.4f} ms {nbytes / ms / 1e6:8.1f} GB/s  ok={ok}  {lib.xtb_last_kernel().decode()[:80]}", flush=True)
    del d
