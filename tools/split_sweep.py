"""Sweep the row-split count of k_reduce_outer (option reduce_split) and the PDL switch on the strided-axis
reductions of cfg3 / cfg5 (CUDA events, each case checked against the default's result).
usage: python tools/split_sweep.py [cfg3] [cfg5s] [cfg5]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xtensor_b200 import capi  # noqa: E402
from xtensor_b200 import expr as xt  # noqa: E402

lib = capi.lib()
capi.check(lib.xtb_init(0))


def timed(fn, iters=20, warm=3):
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


def red(r, o):
    return xt._run_reducer(r, xt.DeviceArray, out=o)


which = sys.argv[1:] or ["cfg3", "cfg5:16384", "cfg5:32768", "cfg5:65536", "cfg5:131072", "cfg5:262144"]
WAVES = (1, 2, 3, 4, 6, 8, 12)
rng = np.random.default_rng(0)
for w in which:
    if w == "cfg3":
        x = xt.DeviceArray.from_numpy(rng.integers(-8, 9, (4096, 4096, 16)).astype(np.float32))
        cases = {"sum": lambda o: red(xt.sum(x, [0]), o), "amax": lambda o: red(xt.amax(x, [0]), o)}
        gx = 64
        nbytes = 4096 * 4096 * 16 * 4
    else:
        rows = int(w.split(":")[1])
        x = xt.DeviceArray.from_numpy(rng.integers(-8, 9, (rows, 8192)).astype(np.float32))
        m = xt.DeviceArray.from_numpy(rng.integers(-2, 3, (8192,)).astype(np.float32))
        cases = {"sum": lambda o: red(xt.sum(x, [0]), o), "var": lambda o: red(xt.sum(xt.square(x - m), [0]), o)}
        gx = 8
        nbytes = rows * 8192 * 4
    splits = [0] + [444 * k // gx for k in WAVES]
    for name, f in cases.items():
        ref = None
        for sp in splits:
            capi.check(lib.xtb_set_option(b"reduce_split", sp))
            o = f(None)
            ms = timed(lambda: f(o))
            r = o.numpy()
            if ref is None:
                ref = r
            ok = np.array_equal(r, ref)
            print(f"{w:12s} {name:5s} split={sp:4d} {ms:8.4f} ms {nbytes / ms / 1e6:8.1f} GB/s ok={ok} {lib.xtb_last_kernel().decode()[:60]}", flush=True)
    capi.check(lib.xtb_set_option(b"reduce_split", 0))
    del x
