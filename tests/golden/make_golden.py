"""Generate tests/golden/ref_vectors.npz from the REAL reference (oracle/_ref/libxtref.so =
xtensor 0.27.1 headers from /root/reference compiled through the xtl stand-in, -O2
-ffp-contract=off).  Run in the build container (where /root/reference exists):

    make -C oracle/ref && python tests/golden/make_golden.py

The .npz holds inputs and the reference's outputs for small instances of every BASELINE
config and of every functor / reducer / accumulator on the path; tests/test_oracle_golden.py
checks the CPU oracle against it, tests/test_gpu_golden.py checks the device."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
from oracle import refbin  # noqa: E402

F32, F64 = np.float32, np.float64
UNARY = ["abs", "exp", "exp2", "expm1", "log", "log10", "log2", "log1p", "sqrt", "cbrt", "sin", "cos", "tan", "asin",
         "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh", "erf", "erfc", "tgamma", "lgamma", "ceil",
         "floor", "trunc", "round", "nearbyint", "rint", "sign", "deg2rad", "rad2deg", "square", "cube", "neg"]
DOMAIN = {"log": (0.01, 50), "log10": (0.01, 50), "log2": (0.01, 50), "log1p": (-0.9, 50), "sqrt": (0, 50),
          "asin": (-1, 1), "acos": (-1, 1), "acosh": (1, 50), "atanh": (-0.99, 0.99), "tgamma": (0.1, 20),
          "lgamma": (0.1, 50), "exp": (-20, 20), "exp2": (-20, 20), "expm1": (-5, 5), "sinh": (-10, 10), "cosh": (-10, 10)}
BINARY = ["add", "sub", "mul", "div", "fmod", "remainder", "fmax", "fmin", "fdim", "pow", "hypot", "atan2", "maximum",
          "minimum", "where_gt", "clip_fma"]
AXES = [[0], [1], [2], [3], [0, 1], [1, 2], [2, 3], [0, 2], [1, 3], [0, 1, 2], [1, 2, 3], [0, 1, 2, 3]]


def main():
    g = {}
    rng = np.random.default_rng(42)   # same seed as the reference's test/files/preprocess.py:108
    # cfg1 / cfg2 / cfg4 / cfg5 / benchmark_assign, small instances
    a, b = rng.uniform(-1, 1, 1003), rng.uniform(-1, 1, 1003)
    g["cfg1_a"], g["cfg1_b"], g["cfg1_out"] = a, b, refbin.cfg1(a, b)
    a = rng.uniform(-np.pi, np.pi, (6, 10, 12)).astype(F32)
    b = rng.uniform(0.5, 1.5, (1, 10, 1)).astype(F32)
    d = rng.uniform(-np.pi, np.pi, (6, 10, 12)).astype(F32)
    g["cfg2_a"], g["cfg2_b"], g["cfg2_d"], g["cfg2_out"] = a, b, d, refbin.cfg2(a, b, d)
    a, b = rng.uniform(-1, 1, (18, 18)), rng.uniform(-1, 1, (36, 18))
    g["cfg4_a"], g["cfg4_b"], g["cfg4_out"] = a, b, refbin.cfg4(a, b)
    a, m = rng.uniform(-1, 1, (20, 24)).astype(F32), rng.uniform(-0.1, 0.1, 24).astype(F32)
    g["cfg5_a"], g["cfg5_m"], g["cfg5_out"] = a, m, refbin.cfg5_map(a, m)
    x, y = rng.uniform(-3, 3, (16, 12)), rng.uniform(-3, 3, (16, 12))
    g["axmby_x"], g["axmby_y"], g["axmby_out"] = x, y, refbin.axmby(x, y)
    # broadcasting shapes of test/test_extended_broadcast_view.cpp:126-810
    for i, (sa, sb) in enumerate([((5, 1, 7), (1, 5, 1, 7)), ((7,), (5, 1, 7)), ((5, 1, 7), (1, 1, 1, 7)),
                                  ((1, 5, 1, 7), (2, 5, 4, 7))]):
        a, b = rng.uniform(-1, 1, sa), rng.uniform(-1, 1, sb)
        g[f"bcast{i}_a"], g[f"bcast{i}_b"], g[f"bcast{i}_out"] = a, b, refbin.bcast_add(a, b)
    # functors
    for dt, tag in ((F32, "f32"), (F64, "f64")):
        for name in UNARY:
            lo, hi = DOMAIN.get(name, (-6.0, 6.0))
            a = rng.uniform(lo, hi, 96).astype(dt)
            g[f"un_{name}_{tag}_in"], g[f"un_{name}_{tag}_out"] = a, refbin.unary(name, a)
        for name in BINARY:
            a, b = rng.uniform(0.1, 9, 96).astype(dt), rng.uniform(0.1, 4, 96).astype(dt)
            g[f"bin_{name}_{tag}_a"], g[f"bin_{name}_{tag}_b"], g[f"bin_{name}_{tag}_out"] = a, b, refbin.binary(name, a, b)
    for dt in (np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64):
        info = np.iinfo(dt)
        a = rng.integers(max(info.min, -100), min(info.max, 100) + 1, 80).astype(dt)
        b = rng.integers(1, min(info.max, 50) + 1, 80).astype(dt)
        n = np.dtype(dt).name
        g[f"int_{n}_a"], g[f"int_{n}_b"], g[f"int_{n}_out"] = a, b, refbin.int_expr(a, b)
    # reducers: random fp32/fp64 (order-sensitive) and int32, lazy and immediate
    for dt, tag in ((F32, "f32"), (F64, "f64"), (np.int32, "i32")):
        a = (rng.uniform(-1, 1, (4, 5, 6, 7)) if tag != "i32" else rng.integers(-9, 10, (4, 5, 6, 7))).astype(dt)
        g[f"red_{tag}_in"] = a
        for op, oname in enumerate(["sum", "prod", "amax", "amin"]):
            for ax in AXES:
                for mode in (0, 1):
                    key = f"red_{tag}_{oname}_{''.join(map(str, ax))}_{'imm' if mode else 'lazy'}"
                    g[key] = refbin.reduce(op, a, ax, mode=mode)
        g[f"red_{tag}_sum_13_keep"] = refbin.reduce(0, a, [1, 3], keep_dims=True)
    a = rng.uniform(-1, 1, (4, 5, 6, 7)).astype(F32)
    g["mean_in"] = a
    g["mean_0_f64"] = refbin.mean_f32(a, [0])
    g["mean_13_f64"] = refbin.mean_f32(a, [1, 3])
    g["mean_2_f32"] = refbin.mean_f32(a, [2], as_f32=True)
    g["var_02_f32"] = refbin.variance(a, [0, 2])
    ad = rng.uniform(-1, 1, (4, 5, 6, 7))       # var/std over axes (0,2) as test_extended_xmath_reducers.cppy:27-80
    g["var_in_f64"] = ad
    g["var_02_f64"] = refbin.variance(ad, [0, 2])
    g["var_02_ddof1_f64"] = refbin.variance(ad, [0, 2], ddof=1)
    # accumulators
    for dt, tag in ((F32, "f32"), (F64, "f64"), (np.int32, "i32"), (np.int16, "i16")):
        a = (rng.uniform(-1, 1, (3, 4, 5)) if tag[0] == "f" else rng.integers(-9, 10, (3, 4, 5))).astype(dt)
        g[f"cumsum_{tag}_in"] = a
        for axis in (None, 0, 1, 2):
            g[f"cumsum_{tag}_{'flat' if axis is None else axis}"] = refbin.cumsum(a, axis)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.npz")
    np.savez_compressed(out, **g)
    print(f"wrote {out}: {len(g)} arrays, {os.path.getsize(out)} bytes")


if __name__ == "__main__":
    main()
