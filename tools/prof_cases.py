"""Run each BASELINE config a few times (for ncu captures).  usage: python tools/prof_cases.py [case ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xtensor_b200 import capi  # noqa: E402
from xtensor_b200 import expr as xt  # noqa: E402

capi.check(capi.lib().xtb_init(0))
cases = sys.argv[1:] or ["cfg1", "cfg2", "cfg3a0", "cfg3a2", "cfg4", "cfg5map", "cfg5sum", "cfg5var", "cumsum_flat", "cumsum_ax1", "cumsum_ax0"]
rng = np.random.default_rng(0)
REPS = int(os.environ.get("REPS", "3"))
for c in cases:
    if c == "cfg1":
        a, b = (xt.DeviceArray.from_numpy(rng.uniform(-1, 1, 1 << 24)) for _ in range(2))
        o = xt.DeviceArray.empty((1 << 24,), xt.F64)
        f = lambda: xt.assign(o, a + b)
    elif c == "cfg2":
        a, d = (xt.DeviceArray.from_numpy(rng.uniform(-3, 3, (1024, 1024, 64)).astype(np.float32)) for _ in range(2))
        b = xt.DeviceArray.from_numpy(rng.uniform(0.5, 1.5, (1, 1024, 1)).astype(np.float32))
        o = xt.DeviceArray.empty((1024, 1024, 64), xt.F32)
        f = lambda: xt.assign(o, xt.sin(a) * b + np.float32(2.0) * d)
    elif c in ("cfg3a0", "cfg3a2", "cfg3max0", "cfg3max2"):
        x = xt.DeviceArray.from_numpy(rng.uniform(-1, 1, (4096, 4096, 16)).astype(np.float32))
        ax = [0] if c.endswith("0") else [2]
        red = xt.amax if "max" in c else xt.sum
        f = lambda: xt.evaluate(red(x, ax))
    elif c == "cfg4":
        a = xt.DeviceArray.from_numpy(rng.uniform(-1, 1, (8192, 8192)))
        b = xt.DeviceArray.from_numpy(rng.uniform(-1, 1, (16384, 8192)))
        o = xt.DeviceArray.empty((8192, 8192), xt.F64)
        f = lambda: xt.assign(o, xt.transpose(a) + xt.view(b, slice(0, None, 2), slice(None)))
    elif c in ("cfg5map", "cfg5sum", "cfg5var"):
        rows = int(os.environ.get("ROWS", "65536"))
        a = xt.DeviceArray.from_numpy(rng.uniform(-1, 1, (rows, 8192)).astype(np.float32))
        m = xt.DeviceArray.from_numpy(rng.uniform(-0.1, 0.1, (8192,)).astype(np.float32))
        o = xt.DeviceArray.empty((rows, 8192), xt.F32)
        if c == "cfg5map":
            f = lambda: xt.assign(o, xt.exp(a - m))
        elif c == "cfg5sum":
            f = lambda: xt.evaluate(xt.sum(a, [0]))
        else:
            f = lambda: xt.evaluate(xt.sum(xt.square(a - m), [0]))
    elif c in ("cumsum_flat", "cumsum_ax1", "cumsum_ax0"):
        x = xt.DeviceArray.from_numpy(rng.uniform(-1, 1, 1 << 26).astype(np.float32))
        x2 = x.reshape_view((8192, 8192))
        f = {"cumsum_flat": lambda: xt.cumsum(x), "cumsum_ax1": lambda: xt.cumsum(x2, 1), "cumsum_ax0": lambda: xt.cumsum(x2, 0)}[c]
    else:
        raise SystemExit(f"unknown case {c}")
    for _ in range(REPS):
        f()
    xt.sync()
    print(c, capi.lib().xtb_last_kernel().decode())
