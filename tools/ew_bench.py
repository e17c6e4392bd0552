"""Time xtb_assign on aligned and odd-shaped broadcast expressions (CUDA events), checked against numpy.
usage: python tools/ew_bench.py"""
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


rng = np.random.default_rng(0)
D = xt.DeviceArray.from_numpy
CASES = {
    "add_dense_odd_8191x8190": lambda: (lambda a, b: (a + b, [a, b]))(rng.uniform(-1, 1, (8191, 8190)).astype(np.float32), rng.uniform(-1, 1, (8191, 8190)).astype(np.float32)),
    "bcast_row_odd_8191x8190": lambda: (lambda a, b: (a * b, [a, b]))(rng.uniform(-1, 1, (8191, 8190)).astype(np.float32), rng.uniform(-1, 1, (8191, 1)).astype(np.float32)),
    "bcast_col_odd_8191x8190": lambda: (lambda a, b: (a - b, [a, b]))(rng.uniform(-1, 1, (8191, 8190)).astype(np.float32), rng.uniform(-1, 1, (8190,)).astype(np.float32)),
    "cfg2_odd_1023x1023x63": lambda: (lambda a, b, d: (a * b + np.float32(2) * d, [a, b, d]))(rng.uniform(-1, 1, (1023, 1023, 63)).astype(np.float32), rng.uniform(-1, 1, (1, 1023, 1)).astype(np.float32), rng.uniform(-1, 1, (1023, 1023, 63)).astype(np.float32)),
    "bcast_col_f64_odd_4095x4097": lambda: (lambda a, b: (a - b, [a, b]))(rng.uniform(-1, 1, (4095, 4097)), rng.uniform(-1, 1, (4097,))),
    "view_offset_8192x8191": lambda: None,
}
for name, mk in CASES.items():
    if name == "view_offset_8192x8191":
        a = rng.uniform(-1, 1, (8192, 8192)).astype(np.float32)
        b = rng.uniform(-1, 1, (8191,)).astype(np.float32)
        A, B = D(a), D(b)
        e = xt.view(A, slice(None), slice(1, None)) + B
        want = a[:, 1:] + b
        inputs = [a, b]
    else:
        want, inputs = mk()
        devs = [D(x) for x in inputs]
        if name.startswith("add") or name.startswith("bcast_row"):
            e = devs[0] + devs[1] if name.startswith("add") else devs[0] * devs[1]
        elif name.startswith("bcast_col"):
            e = devs[0] - devs[1]
        else:
            e = devs[0] * devs[1] + np.float32(2) * devs[2]
    out = xt.evaluate(e)
    ms = timed(lambda: xt.assign(out, e))
    got = out.numpy()
    ok = np.array_equal(got, want.astype(got.dtype))
    nbytes = sum(x.nbytes for x in inputs) + got.nbytes
    print(f"{name:32s} {ms:8.4f} ms {nbytes / ms / 1e6:8.1f} GB/s ok={ok} {lib.xtb_last_kernel().decode()[:70]}", flush=True)
