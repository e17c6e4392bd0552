import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
from xtensor_b200 import capi
from xtensor_b200 import expr as xt
capi.check(capi.lib().xtb_init(0))
for shape, axis, dt in [((8192*14,), None, np.int32), ((8192*200,), None, np.int32), ((8192*257,), None, np.int32), ((8192*300+5,), None, np.int32), ((8192*600,), None, np.int32),
                        ((64, 8192*2), 1, np.int32), ((8192, 8192), 1, np.float64), ((3, 8192*300), 1, np.float32)]:
    n = int(np.prod(shape))
    a = np.ones(shape, dt)
    for rep in range(2):
        r = xt.cumsum(xt.DeviceArray.from_numpy(a), axis).numpy()
        want = np.cumsum(a, axis=axis, dtype=r.dtype).reshape(r.shape)
        bad = np.flatnonzero(r.reshape(-1) != want.reshape(-1))
        msg = "ok" if bad.size == 0 else f"{bad.size} bad, first at {bad[0]} (tile {bad[0] % (shape[-1]) // 8192}) got {r.reshape(-1)[bad[0]]} want {want.reshape(-1)[bad[0]]}; last bad {bad[-1]}"
        print(shape, axis, np.dtype(dt).name, rep, msg, flush=True)
