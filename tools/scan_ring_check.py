"""Correctness + timing of the opt-in persistent ring scan (XTB_SCAN_RING=1, k_scan_ring) against numpy and
against the default kernel.  usage: python tools/scan_ring_check.py   (prints one JSON line per case)"""
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


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    e0, e1 = C.c_void_p(), C.c_void_p()
    capi.check(lib.xtb_event_create(C.byref(e0)))
    capi.check(lib.xtb_event_create(C.byref(e1)))
    capi.check(lib.xtb_sync())
    capi.check(lib.xtb_event_record(e0))
    for _ in range(iters):
        fn()
    capi.check(lib.xtb_event_record(e1))
    capi.check(lib.xtb_sync())
    ms = C.c_float()
    capi.check(lib.xtb_event_elapsed_ms(e0, e1, C.byref(ms)))
    return ms.value / iters


def run(x, axis, ring, prod=False):
    if ring:
        os.environ["XTB_SCAN_RING"] = os.environ.get("RING_VARIANT", "1")
    else:
        os.environ.pop("XTB_SCAN_RING", None)
    d = xt.DeviceArray.from_numpy(x)
    f = xt.cumprod if prod else xt.cumsum
    out = f(d, axis)
    xt.sync()
    return out.numpy(), lib.xtb_last_kernel().decode()


CASES = [  # shape, axis, dtype, prod
    ((1 << 20,), None, np.float32, False), (((1 << 20) + 4 * 777,), None, np.float32, False), ((1 << 22,), None, np.int32, False),
    ((1 << 21,), None, np.float64, False), ((1 << 20,), None, np.int64, False), ((3, (1 << 19) + 8), 1, np.float32, False),
    ((5, 1 << 18), 1, np.float64, False), ((1 << 20,), None, np.float64, True), ((1 << 26,), None, np.float32, False),
]
ok_all = True
for shape, axis, dt, prod in CASES:
    rng = np.random.default_rng(3)
    if prod:
        x = np.where(rng.random(shape) < 0.5, 1.0, -1.0).astype(dt)     # exact products
    elif np.prod(shape) > (1 << 24) and np.dtype(dt).kind == "f":
        x = (rng.random(shape) < 0.2).astype(dt)                          # sums stay below 2^24: exact in fp32
    else:
        x = rng.integers(-4, 5, shape).astype(dt)
    got, kern = run(x, axis, True, prod)
    ref = (np.cumprod if prod else np.cumsum)(x.astype(np.float64) if np.dtype(dt).kind == "f" else x, axis=axis).astype(dt)
    ok = bool(np.array_equal(got.reshape(ref.shape), ref)) and kern.startswith("k_scan_ring")
    ok_all = ok_all and ok
    rec = {"shape": list(shape), "axis": axis, "dtype": np.dtype(dt).name, "prod": prod, "kernel": kern, "ok": ok}
    if not ok:
        bad = np.flatnonzero(got.reshape(-1) != ref.reshape(-1))
        rec["first_bad"] = int(bad[0]) if bad.size else -1
        rec["n_bad"] = int(bad.size)
    print(json.dumps(rec), flush=True)

# timing: flat 2^26 fp32, ring variants (XTB_SCAN_RING=1..4: look-back warps / skew; 5..7: reduce ahead by 8 / 16 / 32 tiles) vs default
x = np.random.default_rng(2).uniform(-1, 1, 1 << 26).astype(np.float32)
d = xt.DeviceArray.from_numpy(x)
y = xt.DeviceArray.empty((1 << 26,), xt.F32)
xi = np.random.default_rng(4).integers(-4, 5, (1 << 22) + 12).astype(np.int32)
di = xt.DeviceArray.from_numpy(xi)
nbytes = 2 * (1 << 26) * 4
res = {}
for variant in ("1", "3", "5", "6", "7", None, "6"):
    if variant:
        os.environ["XTB_SCAN_RING"] = variant
    else:
        os.environ.pop("XTB_SCAN_RING", None)
    good = bool(np.array_equal(xt.cumsum(di).numpy(), np.cumsum(xi)))
    ok_all = ok_all and good
    ms = timed(lambda: xt.cumsum(d, out=y))
    res[f"{variant}:{lib.xtb_last_kernel().decode()}"] = {"ms": round(ms, 4), "GBs": round(nbytes / ms / 1e6, 1), "int32_exact": good}
os.environ.pop("XTB_SCAN_RING", None)
print(json.dumps({"timing_flat_f32_2^26": res, "all_ok": ok_all}), flush=True)
