"""Opt-in invariant-leaf variant of k_reduce_outer (XTB_REDUCE_INV=1): same bits as the default kernel, timing of
the cfg5 variance pass sum(square(a - mean), {0}).  usage: python tools/inv_check.py [rows]"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xtensor_b200 import capi  # noqa: E402
from xtensor_b200 import expr as xt  # noqa: E402

lib = capi.lib()
capi.check(lib.xtb_init(0))
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cols = 8192
blk = np.random.default_rng(0).uniform(-1, 1, (1024, cols)).astype(np.float32)
a = xt.DeviceArray.empty((rows, cols), xt.F32)
for r0 in range(0, rows, 1024):
    capi.check(lib.xtb_memcpy(C.c_void_p(a.owner.ptr + r0 * cols * 4), C.c_void_p(blk.ctypes.data), blk.nbytes, capi.H2D))
m = xt.DeviceArray.from_numpy(np.random.default_rng(1).uniform(-0.1, 0.1, cols).astype(np.float32))
out = xt.DeviceArray.empty((cols,), xt.F32)
e = xt.sum(xt.square(a - m), [0])


def run():
    xt._run_reducer(e, xt.DeviceArray, out=out)


def timed(iters=10):
    for _ in range(3):
        run()
    e0, e1 = C.c_void_p(), C.c_void_p()
    capi.check(lib.xtb_event_create(C.byref(e0)))
    capi.check(lib.xtb_event_create(C.byref(e1)))
    capi.check(lib.xtb_sync())
    capi.check(lib.xtb_event_record(e0))
    for _ in range(iters):
        run()
    capi.check(lib.xtb_event_record(e1))
    capi.check(lib.xtb_sync())
    ms = C.c_float()
    capi.check(lib.xtb_event_elapsed_ms(e0, e1, C.byref(ms)))
    return ms.value / iters


res = {}
outs = {}
for on in (False, True, False, True):
    if on:
        os.environ["XTB_REDUCE_INV"] = "1"
    else:
        os.environ.pop("XTB_REDUCE_INV", None)
    ms = timed()
    k = lib.xtb_last_kernel().decode()
    outs[on] = out.numpy().copy()
    res.setdefault(k, []).append(round(ms, 4))
nbytes = rows * cols * 4
print(json.dumps({"rows": rows, "ms": res, "GBs": {k: round(nbytes / min(v) / 1e6, 1) for k, v in res.items()},
                  "same_bits": bool(np.array_equal(outs[False].view(np.int32), outs[True].view(np.int32)))}), flush=True)
