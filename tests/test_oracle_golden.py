"""CPU: pin the oracle (oracle/xtb_oracle.cpp) against
  (1) tests/golden/ref_vectors.npz -- outputs of the REAL reference (xtensor 0.27.1 compiled
      here through the xtl stand-in; generator: tests/golden/make_golden.py), bit-exact;
  (2) tests/golden/reference_kats.json -- literal expectations of the reference's own tests.
Runs without a GPU and without /root/reference."""
import json
import os

import numpy as np
import pytest

from util import assert_bit_exact

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "ref_vectors.npz"))
KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))
F32 = np.float32


@pytest.fixture(scope="module")
def H(xt):
    return xt.HostArray.from_numpy


def test_baseline_configs(xt, H):
    assert_bit_exact(xt.evaluate(H(G["cfg1_a"]) + H(G["cfg1_b"])).numpy(), G["cfg1_out"])
    e = xt.sin(H(G["cfg2_a"])) * H(G["cfg2_b"]) + F32(2.0) * H(G["cfg2_d"])
    assert_bit_exact(xt.evaluate(e).numpy(), G["cfg2_out"])
    e = xt.transpose(H(G["cfg4_a"])) + xt.view(H(G["cfg4_b"]), slice(0, None, 2), slice(None))
    assert_bit_exact(xt.evaluate(e).numpy(), G["cfg4_out"])
    assert_bit_exact(xt.evaluate(xt.exp(H(G["cfg5_a"]) - H(G["cfg5_m"]))).numpy(), G["cfg5_out"])
    assert_bit_exact(xt.evaluate(3.0 * H(G["axmby_x"]) - 2.0 * H(G["axmby_y"])).numpy(), G["axmby_out"])


@pytest.mark.parametrize("i", range(4))
def test_broadcast_shapes(xt, H, i):
    assert_bit_exact(xt.evaluate(H(G[f"bcast{i}_a"]) + H(G[f"bcast{i}_b"])).numpy(), G[f"bcast{i}_out"])


UNARY = sorted({k.split("_")[1] for k in G.files if k.startswith("un_")})
BINARY = sorted({k[4:].rsplit("_", 2)[0] for k in G.files if k.startswith("bin_") and k.endswith("_out")})


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("name", UNARY)
def test_unary_functors(xt, H, name, tag):
    a = H(G[f"un_{name}_{tag}_in"])
    e = -a if name == "neg" else getattr(xt, name)(a)
    assert_bit_exact(xt.evaluate(e).numpy(), G[f"un_{name}_{tag}_out"])


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("name", BINARY)
def test_binary_functors(xt, H, name, tag):
    a, b = H(G[f"bin_{name}_{tag}_a"]), H(G[f"bin_{name}_{tag}_b"])
    T = G[f"bin_{name}_{tag}_a"].dtype.type
    ops = {"add": lambda: a + b, "sub": lambda: a - b, "mul": lambda: a * b, "div": lambda: a / b,
           "where_gt": lambda: xt.where(a > b, a, b * T(0.5)),
           "clip_fma": lambda: xt.clip(a, T(-1), T(1)) + xt.fma(a, b, a)}
    e = ops[name]() if name in ops else getattr(xt, name)(a, b)
    assert_bit_exact(xt.evaluate(e).numpy(), G[f"bin_{name}_{tag}_out"])


@pytest.mark.parametrize("n", ["int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64"])
def test_integer_expression(xt, H, n):
    a, b = H(G[f"int_{n}_a"]), H(G[f"int_{n}_b"])
    e = (a + b) * a - (a / b) + (a % b) + (a & b) - (a | b) + (a ^ b)
    assert_bit_exact(xt.evaluate(e).numpy(), G[f"int_{n}_out"])


RED_KEYS = [k for k in G.files if k.startswith("red_") and not k.endswith("_in") and not k.endswith("_keep")]


@pytest.mark.parametrize("key", RED_KEYS)
def test_reducers_lazy_and_immediate(xt, H, key):
    """Order-sensitive fp sums must match the reference bit for bit in BOTH evaluation orders."""
    _, tag, oname, axes, mode = key.split("_")
    a = H(G[f"red_{tag}_in"])
    r = getattr(xt, oname)(a, [int(c) for c in axes])
    got = xt._run_reducer(r, xt.HostArray, mode=1 if mode == "imm" else 0).numpy()
    assert_bit_exact(got, G[key])


def test_keep_dims_mean_variance(xt, H):
    for tag in ("f32", "f64", "i32"):
        got = xt.evaluate(xt.sum(H(G[f"red_{tag}_in"]), [1, 3], keep_dims=True)).numpy()
        assert_bit_exact(got, G[f"red_{tag}_sum_13_keep"])
    a = H(G["mean_in"])
    assert_bit_exact(xt.evaluate(xt.mean(a, [0])).numpy(), G["mean_0_f64"])
    assert_bit_exact(xt.evaluate(xt.mean(a, [1, 3])).numpy(), G["mean_13_f64"])
    assert_bit_exact(xt.evaluate(xt.mean(a, [2], dtype=xt.F32)).numpy(), G["mean_2_f32"])
    assert_bit_exact(xt.evaluate(xt.variance(a, [0, 2], dtype=xt.F32)).numpy(), G["var_02_f32"])
    ad = H(G["var_in_f64"])
    assert_bit_exact(xt.evaluate(xt.variance(ad, [0, 2])).numpy(), G["var_02_f64"])
    assert_bit_exact(xt.evaluate(xt.variance(ad, [0, 2], ddof=1)).numpy(), G["var_02_ddof1_f64"])
    # the reference's own check of these: numpy var/std with xt::allclose (rtol 1e-5, atol 1e-8)
    assert np.allclose(G["var_02_f64"], G["var_in_f64"].var(axis=(0, 2)), rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("tag", ["f32", "f64", "i32", "i16"])
@pytest.mark.parametrize("axis", [None, 0, 1, 2])
def test_cumsum(xt, H, tag, axis):
    got = xt.cumsum(H(G[f"cumsum_{tag}_in"]), axis).numpy()
    assert_bit_exact(got, G[f"cumsum_{tag}_{'flat' if axis is None else axis}"])


def test_kats_reducer_fixture(xt, H):
    k = KATS["reducer_fixture"]
    a = np.ones(k["shape"])
    a[1, :, 1, :, 1] = 2
    r = xt.evaluate(xt.sum(H(a), k["axes"])).numpy()
    assert r.shape == (3, 4, 5) and r[0, 0, 0] == k["red_000"] and r[1, 1, 1] == k["red_111"]
    assert float(xt.evaluate(xt.sum(H(a))).numpy()) == k["sum_all"]
    s = xt.evaluate(xt.sum(H(np.ones(1000, np.uint8)))).numpy()
    assert s.dtype == np.int32 and int(s) == k["uint8_ones_1000"]


def test_kats_accumulator(xt, H):
    k = KATS["accumulator_one_d"]
    r = xt.cumsum(H(np.array(k["input_int16"], np.int16))).numpy()
    assert r.dtype == np.int32 and r.tolist() == k["expected_int32"]
    r0 = xt.cumsum(H(np.array(k["input_int16"], np.int16)), 0).numpy()
    assert r0.tolist() == k["expected_int32"]
    k = KATS["accumulator_four_d"]
    a = np.arange(36, dtype=np.float64).reshape(k["shape"])
    assert xt.cumsum(H(a)).numpy().tolist() == k["flat"]
    assert xt.cumsum(H(a), 0).numpy().reshape(-1).tolist() == k["axis0"]
    assert xt.cumsum(H(a), 1).numpy().reshape(-1).tolist() == k["axis1"]
    one = np.array([[5.0, 6.0, 7.0]])           # dim_one (test_xaccumulator.cpp:36-45)
    assert np.array_equal(xt.cumsum(H(one), 0).numpy(), one)
    assert np.array_equal(xt.cumsum(xt.transpose(H(one)), 1).numpy(), one.T)


def test_kats_layout_fixture(xt, H):
    k = KATS["layout_fixture"]
    data = np.array(k["row_major_data"], np.int32).reshape(k["shape"])
    rm = H(data)
    cm = H(np.ascontiguousarray(data.transpose(2, 1, 0))).transpose([2, 1, 0])
    ctm = H(np.ascontiguousarray(data.transpose(0, 2, 1))).transpose([0, 2, 1])
    assert rm.strides == (8, 4, 1) and cm.strides == (1, 3, 6) and ctm.strides == (8, 1, 2)
    unit = H(np.ascontiguousarray(data[:, :1, :]))
    assert unit.strides == (4, 0, 1)             # stride 0 on the unit dimension
    for x in (rm, cm, ctm):
        assert np.array_equal(xt.evaluate(rm + x).numpy(), 2 * data)
        assert np.array_equal(xt.evaluate(rm * x - x).numpy(), data * data - data)
    assert np.array_equal(xt.evaluate(cm + unit).numpy(), data + data[:, :1, :])


def test_kats_result_types(xt, H):
    k = KATS["result_types"]
    name = {"bool": xt.BOOL, "i8": xt.I8, "u8": xt.U8, "i16": xt.I16, "u16": xt.U16, "i32": xt.I32, "u32": xt.U32,
            "i64": xt.I64, "u64": xt.U64, "f32": xt.F32, "f64": xt.F64}
    for a, b, op, res in k["cases"]:
        A = H(np.ones(3, xt.NP_OF[name[a]]))
        B = H(np.ones(3, xt.NP_OF[name[b]]))
        assert xt.Func(op, (A, B)).dtype == name[res], (a, b, op)
    f = H(np.ones((2, 3), np.float32))
    assert xt.mean(f, [0]).dtype == name[k["mean_of_f32_is"]]
    assert xt.sum(H(np.ones(3, np.uint8))).dtype == name[k["sum_of_u8_is"]]
    assert xt.cumsum(H(np.ones(3, np.int16))).dtype == name[k["cumsum_of_i16_is"]]
    assert xt.sqrt(H(np.ones(3, np.int32))).dtype == name[k["sqrt_of_i32_is"]]
    assert xt.sin(f).dtype == name[k["sin_of_f32_is"]]
