"""Generate tests/golden/ref_vectors_nan.npz from the REAL reference (oracle/_ref/libxtref.so, see
make_golden.py): nan-aware reducers, counts and nan_to_num (core/xmath.hpp:2307-2860) on inputs with
NaN / +-inf sprinkled in, including all-NaN lanes.  Run in the build container:

    make -C oracle/ref && python tests/golden/make_golden_nan.py
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
from oracle import refbin  # noqa: E402

NAMES = ["nansum", "nanprod", "nanmin", "nanmax", "nanmean", "nanvar", "nanstd", "nanmean_t", "nanvar_t",
         "count_nonzero", "count_nonnan"]
AXES = [[0], [1], [2], [0, 1], [1, 2], [0, 2], [0, 1, 2]]


def make_input(rng, dt):
    # small integers: every sum / product is exact in any order, so the device may reorder and stay bit-exact
    a = rng.integers(-3, 4, (6, 5, 7)).astype(dt)
    a[rng.random(a.shape) < 0.2] = np.nan
    a[2, :, 3] = np.nan          # an all-NaN lane along axis 1
    a[:, 4, 6] = np.nan          # ... along axis 0
    a[5, 1, :] = np.nan          # ... along axis 2
    return a


def main():
    g = {}
    rng = np.random.default_rng(7)
    for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
        a = make_input(rng, dt)
        g[f"in_{tag}"] = a
        for name in NAMES:
            for ax in AXES:
                g[f"{name}_{tag}_ax{''.join(map(str, ax))}"] = refbin.nanfn(name, a, ax)
        x = rng.uniform(-5, 5, 64).astype(dt)
        x[::7] = np.nan
        x[3::11] = np.inf
        x[5::13] = -np.inf
        g[f"n2n_{tag}_in"], g[f"n2n_{tag}_out"] = x, refbin.nan_to_num(x)
    i = rng.integers(-2, 3, (6, 5, 7)).astype(np.int32)
    g["cnz_i32_in"] = i
    for ax in AXES:
        g[f"cnz_i32_ax{''.join(map(str, ax))}"] = refbin.count_nonzero_i32(i, ax)
    # xt::average: weights along one axis / of the full shape / whole array (integer-valued: exact in any order)
    a = rng.integers(-5, 6, (4, 6, 5)).astype(np.float64)
    g["avg_in"] = a
    for axis in range(3):
        w = rng.integers(1, 5, a.shape[axis]).astype(np.float64)
        g[f"avg_w1_ax{axis}"], g[f"avg_out1_ax{axis}"] = w, refbin.average(a, w, [axis])
    wf = rng.integers(1, 5, a.shape).astype(np.float64)
    g["avg_wfull"] = wf
    for ax in ([0], [1, 2], [0, 1, 2]):
        g[f"avg_outfull_ax{''.join(map(str, ax))}"] = refbin.average(a, wf, ax)
    g["avg_outfull_all"] = refbin.average(a, wf, [])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors_nan.npz")
    np.savez_compressed(path, **g)
    print(f"wrote {path}: {len(g)} arrays")


if __name__ == "__main__":
    main()
