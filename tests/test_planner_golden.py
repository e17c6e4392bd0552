"""Reductions over the shapes the device planner rewrites into two passes or serves with scalar-access kernels
(xtb_reduce.cu: reduce_decomposed, V = 1), on RANDOM data, against outputs of the REAL reference
(tests/golden/ref_vectors_planner.npz, generator tests/golden/make_golden_planner.py; inputs are regenerated from the
recorded seeds and guarded by a checksum).

CPU: the oracle reproduces the reference's lazy and immediate results bit for bit at these sizes.
GPU: amax / amin bit-exact; sums within the north-star tolerance of the reference's result (1e-6 fp32 / 1e-12 fp64,
relative to the magnitude of the summands) and at least as close to the fp64 value as the reference itself.
"""
import json
import os

import numpy as np
import pytest

from util import assert_bit_exact

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors_planner.npz"))
META = json.loads(bytes(G["meta"]).decode())
NP = {"f32": np.float32, "f64": np.float64}


def case_input(m):
    a = np.random.default_rng(m["seed"]).uniform(-1, 1, m["shape"]).astype(NP[m["dtype"]])
    u = a.view(np.uint32 if a.dtype.itemsize == 4 else np.uint64).astype(np.uint64)
    # order-independent checksum of the bit patterns (modulo 2^64)
    assert int(u.sum(dtype=np.uint64)) == m["input_checksum"], "numpy's generator changed: regenerate the goldens"
    return a


@pytest.mark.parametrize("m", META, ids=[m["name"] for m in META])
def test_oracle_reproduces_the_reference(xt, m):
    a = case_input(m)
    H = xt.HostArray.from_numpy(a)
    assert_bit_exact(xt._run_reducer(xt.sum(H, m["axes"]), xt.HostArray, mode=0).numpy(), G[f"{m['name']}_sum_lazy"])
    assert_bit_exact(xt._run_reducer(xt.sum(H, m["axes"]), xt.HostArray, mode=1).numpy(), G[f"{m['name']}_sum_immediate"])
    assert_bit_exact(xt._run_reducer(xt.amax(H, m["axes"]), xt.HostArray, mode=1).numpy(), G[f"{m['name']}_amax"])
    assert_bit_exact(xt._run_reducer(xt.amin(H, m["axes"]), xt.HostArray, mode=1).numpy(), G[f"{m['name']}_amin"])


@pytest.mark.gpu
@pytest.mark.parametrize("m", META, ids=[m["name"] for m in META])
def test_device_against_the_reference(xt, gpu, m):
    a = case_input(m)
    D = xt.DeviceArray.from_numpy(a)
    assert_bit_exact(xt.evaluate(xt.amax(D, m["axes"])).numpy(), G[f"{m['name']}_amax"])
    assert_bit_exact(xt.evaluate(xt.amin(D, m["axes"])).numpy(), G[f"{m['name']}_amin"])
    got = xt.evaluate(xt.sum(D, m["axes"])).numpy().astype(np.float64)
    tol = 1e-6 if m["dtype"] == "f32" else 1e-12
    eps = np.finfo(NP[m["dtype"]]).eps
    scale = np.maximum(np.abs(a).sum(axis=tuple(m["axes"]), dtype=np.float64), 1.0)
    exact = a.astype(np.float64).sum(axis=tuple(m["axes"]))
    for mode in ("lazy", "immediate"):
        ref = G[f"{m['name']}_sum_{mode}"].astype(np.float64)
        # the reference's own sequential order drifts by up to ~n * eps / 2 of the summands' magnitude; the device
        # (blocked summation) must agree with it to that accuracy and to the north-star tolerance where that is wider
        n_terms = a.size / got.size
        bound = np.maximum(tol, n_terms * eps) * scale
        assert np.all(np.abs(got - ref) <= bound), float((np.abs(got - ref) / scale).max())
        assert np.abs(got - exact).max() <= np.abs(ref - exact).max() * 1.5 + eps * scale.max()
    assert np.all(np.abs(got - exact) <= tol * scale), float((np.abs(got - exact) / scale).max())
