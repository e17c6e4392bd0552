"""Measure the max ulp distance of every device functor against the CPU oracle (glibc libm).
Run on the GPU box:  python tools/ulp_report.py > gpurun_out/ulp_report.json"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from oracle import oracle  # noqa: E402
from util import run_both, ulp_distance  # noqa: E402

xt = oracle.install()
UNARY = ["exp", "exp2", "expm1", "log", "log10", "log2", "log1p", "sqrt", "cbrt", "sin", "cos", "tan", "asin",
         "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh", "erf", "erfc", "tgamma", "lgamma"]
DOMAIN = {"log": (1e-6, 1e4), "log10": (1e-6, 1e4), "log2": (1e-6, 1e4), "log1p": (-0.999, 1e3), "sqrt": (0, 1e6),
          "asin": (-1, 1), "acos": (-1, 1), "acosh": (1, 1e4), "atanh": (-0.9999, 0.9999), "tgamma": (0.01, 30),
          "lgamma": (0.01, 1e3), "exp": (-80, 80), "exp2": (-100, 100), "expm1": (-20, 20), "sinh": (-20, 20),
          "cosh": (-20, 20), "sin": (-100, 100), "cos": (-100, 100), "tan": (-100, 100), "erf": (-5, 5),
          "erfc": (-3, 9), "tanh": (-10, 10)}
BINARY = {"pow": ((0.01, 20), (-5, 5)), "hypot": ((-1e3, 1e3), (-1e3, 1e3)), "atan2": ((-10, 10), (-10, 10)),
          "fmod": ((-1e3, 1e3), (0.1, 7)), "remainder": ((-1e3, 1e3), (0.1, 7))}
out = {}
n = 1 << 20
for dt in (np.float32, np.float64):
    for name in UNARY:
        lo, hi = DOMAIN.get(name, (-20.0, 20.0))
        a = np.random.default_rng(1).uniform(lo, hi, n).astype(dt)
        g, w = run_both(xt, lambda A: getattr(xt, name)(A), a)
        out[f"{name}/{np.dtype(dt).name}"] = ulp_distance(g, w)
    for name, (da, db) in BINARY.items():
        a = np.random.default_rng(2).uniform(*da, n).astype(dt)
        b = np.random.default_rng(3).uniform(*db, n).astype(dt)
        g, w = run_both(xt, lambda A, B: getattr(xt, name)(A, B), a, b)
        out[f"{name}/{np.dtype(dt).name}"] = ulp_distance(g, w)
print(json.dumps(out, indent=1))
