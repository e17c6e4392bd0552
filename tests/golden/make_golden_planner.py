"""Generate tests/golden/ref_vectors_planner.npz from the REAL reference (oracle/_ref/libxtref.so): reductions over the
shapes the device planner rewrites or serves with scalar-access kernels (xtb_reduce.cu: reduce_decomposed, V = 1) --
narrow (< 1024 outputs), mixed (outer + innermost axis reduced, kept dim between), odd row pitches -- on RANDOM
floating-point data, so that the device's re-ordered summation is held against the reference's own results.
The inputs are several MB each (the planner only acts from 2^20 elements), so only their seeds are stored: the tests
regenerate them with numpy's PCG64 `default_rng(seed).uniform(-1, 1, shape).astype(dtype)`.  Run in the build container:

    make -C oracle/ref && python tests/golden/make_golden_planner.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
from oracle import refbin  # noqa: E402

CASES = [
    ("narrow_rem", (70001, 20), [0], "f32"), ("narrow_3", (1 << 20, 3), [0], "f32"), ("narrow_f64", (100003, 16), [0], "f64"),
    ("narrow_2axes", (512, 512, 16), [0, 1], "f32"), ("mixed_f64", (300, 17, 257), [0, 2], "f64"),
    ("mixed_f32", (512, 256, 16), [0, 2], "f32"), ("odd_ax0", (2047, 1022), [0], "f32"), ("odd_ax1", (2047, 1022), [1], "f32"),
]
NP = {"f32": np.float32, "f64": np.float64}


def bits_checksum(a):
    """Order-independent: the sum of the elements' bit patterns modulo 2^64."""
    u = a.view(np.uint32 if a.dtype.itemsize == 4 else np.uint64).astype(np.uint64)
    return int(u.sum(dtype=np.uint64))


def case_input(seed, shape, tag):
    return np.random.default_rng(seed).uniform(-1, 1, shape).astype(NP[tag])


def main():
    g, meta = {}, []
    for i, (name, shape, axes, tag) in enumerate(CASES):
        seed = 1000 + i
        a = case_input(seed, shape, tag)
        meta.append({"name": name, "shape": list(shape), "axes": axes, "dtype": tag, "seed": seed,
                     "input_checksum": bits_checksum(a)})
        g[f"{name}_sum_lazy"] = refbin.reduce(0, a, axes, mode=0)
        g[f"{name}_sum_immediate"] = refbin.reduce(0, a, axes, mode=1)
        g[f"{name}_amax"] = refbin.reduce(2, a, axes, mode=1)
        g[f"{name}_amin"] = refbin.reduce(3, a, axes, mode=1)
    g["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors_planner.npz")
    np.savez_compressed(out, **g)
    print(f"wrote {out}: {len(g) - 1} arrays, {os.path.getsize(out)} bytes")


if __name__ == "__main__":
    main()
