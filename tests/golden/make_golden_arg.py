"""Generate tests/golden/ref_vectors_arg.npz from the REAL reference (oracle/_ref/libxtref.so): argmin / argmax
(misc/xsort.hpp:1237-1295) over every axis and flattened, on inputs full of ties and -- for floats -- NaNs
(including a NaN in front and all-NaN lanes); minmax (core/xmath.hpp:2195-2228); the norms over one axis
(reducers/xnorm.hpp).  Run in the build container:

    make -C oracle/ref && python tests/golden/make_golden_arg.py
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
from oracle import refbin  # noqa: E402

ARG_DTYPES = {"i8": np.int8, "u8": np.uint8, "i16": np.int16, "i32": np.int32, "i64": np.int64, "u64": np.uint64,
              "f32": np.float32, "f64": np.float64}
NORM_DTYPES = {"i32": np.int32, "u16": np.uint16, "f32": np.float32, "f64": np.float64}
NORMS = [("l0", 0.0), ("l1", 0.0), ("sq", 0.0), ("l2", 0.0), ("linf", 0.0), ("lp_to_p", 3.0), ("lp", 1.5)]


def arg_input(rng, dt, nan):
    lo = 0 if np.dtype(dt).kind == "u" else -4
    a = rng.integers(lo, 5, (6, 5, 8)).astype(dt)              # few distinct values: many ties
    if nan:
        a[rng.random(a.shape) < 0.15] = np.nan
        a[0, 0, 0] = np.nan                                    # a NaN in front sticks: index 0
        a[3, :, 2] = np.nan                                    # all-NaN lanes along every axis
        a[:, 1, 5] = np.nan
        a[4, 2, :] = np.nan
    return a


def main():
    g = {}
    rng = np.random.default_rng(11)
    for tag, dt in ARG_DTYPES.items():
        for nan in ([False, True] if np.dtype(dt).kind == "f" else [False]):
            a = arg_input(rng, dt, nan)
            key = f"{tag}{'_nan' if nan else ''}"
            g[f"arg_in_{key}"] = a
            for fn in ("argmin", "argmax"):
                g[f"{fn}_{key}_flat"] = refbin.argfn(fn, a)
                for ax in range(3):
                    g[f"{fn}_{key}_ax{ax}"] = refbin.argfn(fn, a, ax)
    # the reference's own cases, test/test_xsort.cpp:217-281
    a = np.array([[5, 3, 1], [4, 4, 4]], np.float64)
    g["kat_a"] = a
    for fn in ("argmin", "argmax"):
        g[f"kat_{fn}_flat"] = refbin.argfn(fn, a)
        g[f"kat_{fn}_ax0"], g[f"kat_{fn}_ax1"] = refbin.argfn(fn, a, 0), refbin.argfn(fn, a, 1)
    for tag, dt in (("i16", np.int16), ("i32", np.int32), ("f32", np.float32), ("f64", np.float64)):
        x = rng.integers(-3000, 3000, (40, 33)).astype(dt)
        g[f"minmax_in_{tag}"], g[f"minmax_out_{tag}"] = x, refbin.minmax(x)
    xn = rng.uniform(-5, 5, (9, 11)).astype(np.float32)
    xn[rng.random(xn.shape) < 0.3] = np.nan
    g["minmax_in_f32_nan"], g["minmax_out_f32_nan"] = xn, refbin.minmax(xn)
    for tag, dt in NORM_DTYPES.items():
        lo = 0 if np.dtype(dt).kind == "u" else -6
        a = rng.integers(lo, 7, (5, 8, 6)).astype(dt)
        g[f"norm_in_{tag}"] = a
        for name, p in NORMS:
            for ax in range(3):
                raw, width, shp = refbin.normfn(name, a, ax, p)
                g[f"norm_{name}_{tag}_ax{ax}"] = raw
                g[f"norm_{name}_{tag}_ax{ax}_meta"] = np.array([width] + list(shp), np.int64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors_arg.npz")
    np.savez_compressed(path, **g)
    print(f"wrote {path}: {len(g)} arrays")


if __name__ == "__main__":
    main()
