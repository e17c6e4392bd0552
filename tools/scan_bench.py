"""Time xtb_scan on a set of shapes (CUDA events) and check each against numpy.  usage: python tools/scan_bench.py"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xtensor_b200 import capi  # noqa: E402
from xtensor_b200 import expr as xt  # noqa: E402

lib = capi.lib()
capi.check(lib.xtb_init(0))


def timed(fn, iters=int(os.environ.get("ITERS", "10"))):
    for _ in range(int(os.environ.get("WARM", "3"))):
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
    ((1 << 26,), None, np.float32), ((1 << 26,), None, np.int32), ((1 << 25,), None, np.float64),
    ((8192, 8192), 1, np.float32), ((8192, 8192), 0, np.float32), ((8192, 8192), 1, np.float64), ((8192, 8192), 0, np.float64),
    ((1 << 20, 64), 1, np.float32), ((1 << 17, 512), 1, np.float32), ((64, 1 << 20), 1, np.float32), ((64, 1 << 20), 0, np.float32),
    ((1 << 20, 64), 0, np.float32), ((256, 512, 512), 1, np.float32), ((1 << 26,), None, np.uint8),
    # odd extents: row pitch not a multiple of 16 bytes
    (((1 << 26) + 3,), None, np.float32), ((8191, 8190), 1, np.float32), ((8191, 8190), 0, np.float32),
    ((4095, 4097, 15), 1, np.float32), ((4095, 4097), 0, np.float64), ((4095, 4097), 1, np.float64), ((255, 513, 511), 2, np.float32),
]
only = [a for a in sys.argv[1:] if not a.startswith("--")]
variants = [int(a.split("=")[1]) for a in sys.argv[1:] if a.startswith("--variant=")] or [0]
for a in sys.argv[1:]:
    if a.startswith("--cv="):
        capi.check(lib.xtb_set_option(b"tile_variant", int(a.split("=")[1])))
    if a.startswith("--nv="):
        capi.check(lib.xtb_set_option(b"scan_nv", int(a.split("=")[1])))
for variant in variants:
  capi.check(lib.xtb_set_option(b"scan_variant", variant))
  print(f"# scan_variant={variant} (>0: look-ahead MB of k_scan_ahead, <0: ring stages of k_scan_colwalk)", flush=True)
  for shape, axis, dt in CASES:
      tag = f"{'x'.join(map(str, shape))}:{axis}:{np.dtype(dt).name}"
      if only and not any(o in tag for o in only):
          continue
      rng = np.random.default_rng(1)
      n = int(np.prod(shape))
      a = rng.integers(-3, 4, n).astype(dt).reshape(shape) if dt != np.uint8 else rng.integers(0, 2, n).astype(dt).reshape(shape)
      d = xt.DeviceArray.from_numpy(a)
      res = xt.cumsum(d, axis)

      def f():
          xt.cumsum(d, axis, out=res)

      ms = timed(f)
      r = res.numpy()
      want = np.cumsum(a, axis=axis, dtype=r.dtype)
      ok = np.array_equal(r, want.reshape(r.shape))
      nbytes = n * (a.itemsize + r.itemsize)
      print(f"{tag:32s} {ms:8.4f} ms {nbytes / ms / 1e6:8.1f} GB/s  ok={ok}  {lib.xtb_last_kernel().decode()}", flush=True)
