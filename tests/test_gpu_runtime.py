"""GPU: stream fork / join and CUDA-graph capture of a call sequence give the same results as the plain
sequence (the sharded cfg5 step overlaps its variance reduction with the exp(a - mean) map this way)."""
import ctypes as C

import numpy as np
import pytest

from util import assert_bit_exact

pytestmark = pytest.mark.gpu


def _step(xt, capi, lib, a, mean_, s_sq, o, fork):
    if fork:
        capi.check(lib.xtb_fork_begin())
    xt._run_reducer(xt.sum(xt.square(a - mean_), [0]), xt.DeviceArray, out=s_sq)
    if fork:
        capi.check(lib.xtb_fork_end())
    xt.assign(o, xt.exp(a - mean_))
    if fork:
        capi.check(lib.xtb_fork_join())


@pytest.mark.parametrize("graph", [False, True])
def test_fork_join_matches_sequential(xt, gpu, graph):
    from xtensor_b200 import capi
    lib = capi.lib()
    rng = np.random.default_rng(31)
    an = rng.uniform(-1, 1, (4096, 512)).astype(np.float32)
    a = xt.DeviceArray.from_numpy(an)
    mean_ = xt.DeviceArray.from_numpy(an.mean(axis=0).astype(np.float32))
    outs = []
    for fork in (False, True):
        s_sq = xt.DeviceArray.empty((512,), xt.F32)
        o = xt.DeviceArray.empty(an.shape, xt.F32)
        _step(xt, capi, lib, a, mean_, s_sq, o, fork)          # also sizes both scratch buffers
        if graph:
            g = C.c_void_p()
            capi.check(lib.xtb_graph_begin())
            _step(xt, capi, lib, a, mean_, s_sq, o, fork)
            capi.check(lib.xtb_graph_end(C.byref(g)))
            capi.check(lib.xtb_memset(C.c_void_p(s_sq.owner.ptr), 0, 512 * 4))
            capi.check(lib.xtb_memset(C.c_void_p(o.owner.ptr), 0, an.size * 4))
            for _ in range(3):
                capi.check(lib.xtb_graph_launch(g))
            capi.check(lib.xtb_sync())
            lib.xtb_graph_destroy(g)
        outs.append((s_sq.numpy(), o.numpy()))
    assert_bit_exact(outs[0][0], outs[1][0])
    assert_bit_exact(outs[0][1], outs[1][1])
    ref = ((an.astype(np.float64) - mean_.numpy()) ** 2).sum(axis=0)
    assert np.allclose(outs[1][0], ref, rtol=1e-5)


def test_fork_misuse_is_an_error(xt, gpu):
    from xtensor_b200 import capi
    lib = capi.lib()
    assert lib.xtb_fork_end() != 0
    capi.check(lib.xtb_fork_begin())
    assert lib.xtb_fork_begin() != 0
    capi.check(lib.xtb_fork_end())
    capi.check(lib.xtb_fork_join())
    capi.check(lib.xtb_sync())


def test_foreign_device_memory(xt, gpu):
    """xtb::adapt / DeviceArray.from_pointer: operands and destination in memory the library does not own
    (here: raw xtb_malloc allocations handled by the test), contiguous and with custom strides."""
    import ctypes as C
    from xtensor_b200 import capi
    rng = np.random.default_rng(11)
    a = rng.integers(-9, 10, (6, 40)).astype(np.float32)
    b = rng.integers(-9, 10, (40,)).astype(np.float32)
    ptrs = []
    for nbytes in (a.nbytes, b.nbytes, a.nbytes):
        p = C.c_void_p()
        capi.check(gpu.xtb_malloc(nbytes, C.byref(p)))
        ptrs.append(p)
    try:
        capi.check(gpu.xtb_memcpy(ptrs[0], C.c_void_p(a.ctypes.data), a.nbytes, capi.H2D))
        capi.check(gpu.xtb_memcpy(ptrs[1], C.c_void_p(b.ctypes.data), b.nbytes, capi.H2D))
        A = xt.DeviceArray.from_pointer(ptrs[0].value, a.shape, xt.F32)
        B = xt.DeviceArray.from_pointer(ptrs[1].value, b.shape, xt.F32)
        O = xt.DeviceArray.from_pointer(ptrs[2].value, a.shape, xt.F32)
        xt.assign(O, A * B + np.float32(2.0))                       # written straight into the foreign buffer
        got = np.empty_like(a)
        capi.check(gpu.xtb_memcpy(C.c_void_p(got.ctypes.data), ptrs[2], got.nbytes, capi.D2H))
        capi.check(gpu.xtb_sync())
        assert np.array_equal(got, a * b + np.float32(2.0))
        At = xt.DeviceArray.from_pointer(ptrs[0].value, (40, 6), xt.F32, strides=(1, 40))   # the same memory, transposed
        assert np.array_equal(xt.evaluate(xt.sum(At, [1])).numpy(), a.T.sum(axis=1))
        assert np.array_equal(At.numpy(), a.T)
    finally:
        for p in ptrs:
            gpu.xtb_free(p)


def test_npy_fixture_round_trip(xt, gpu, tmp_path):
    """.npy -> device -> expression -> .npy (the C++ side is xtb::load_npy / dump_npy, test_dropin.cpp)."""
    a = np.random.default_rng(3).integers(-9, 10, (33, 17)).astype(np.float32)
    src, dst = str(tmp_path / "a.npy"), str(tmp_path / "b.npy")
    np.save(src, a)
    d = xt.DeviceArray.from_npy(src)
    assert d.shape == a.shape and np.array_equal(d.numpy(), a)
    xt.evaluate(d * np.float32(2) + np.float32(1)).to_npy(dst)
    assert np.array_equal(np.load(dst), a * 2 + 1)
