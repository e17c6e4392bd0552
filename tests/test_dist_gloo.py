"""CPU, world_size 2 over gloo: the N > 1 path (row sharding, one allreduce of the per-rank
partials when the sharded axis is reduced, mean / variance finalised after the merge) with
host arrays evaluated by the oracle.  On GPUs the same functions use NCCL through the C ABI."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch
    import torch.distributed as dist
    from oracle import oracle
    from xtensor_b200 import capi, shard
    xt = oracle.install()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def allreduce(x, op):
        t = torch.from_numpy(np.ascontiguousarray(x).copy())
        dist.all_reduce(t, op={capi.RED_SUM: dist.ReduceOp.SUM, capi.RED_PROD: dist.ReduceOp.PRODUCT,
                               capi.RED_MAX: dist.ReduceOp.MAX, capi.RED_MIN: dist.ReduceOp.MIN}[op])
        return t.numpy()

    rows, cols = 37, 24          # ragged split: 19 + 18 rows
    full = np.random.default_rng(9).integers(-8, 9, (rows, cols)).astype(np.float32)
    mvec = np.random.default_rng(10).integers(-3, 4, (cols,)).astype(np.float32)
    b, e = shard.row_block(rows, rank, world)
    A = xt.HostArray.from_numpy(full)
    local = shard.shard_operand(A, rows, rank, world)
    M = shard.shard_operand(xt.HostArray.from_numpy(mvec), rows, rank, world)      # replicated
    res = {"block": (b, e), "local_shape": local.shape, "m_shape": M.shape}
    res["sum0"] = shard.sharded_reduce(capi.RED_SUM, local, [0], rows, world, allreduce).numpy()
    res["max0"] = shard.sharded_reduce(capi.RED_MAX, local, [0], rows, world, allreduce).numpy()
    res["sum1"] = shard.sharded_reduce(capi.RED_SUM, local, [1], rows, world, allreduce).numpy()    # stays sharded
    res["mean0"] = shard.sharded_mean(local, [0], rows, world, allreduce, dtype=xt.F32).numpy()
    res["var0"] = shard.sharded_variance(local, [0], rows, world, allreduce, dtype=xt.F32).numpy()
    res["map"] = xt.evaluate(xt.exp(local - M)).numpy()                                              # no exchange
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, res, full, mvec))


@pytest.mark.timeout(300)
def test_sharded_path_world2_gloo(xt):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = {}
    for _ in procs:
        rank, res, full, mvec = q.get(timeout=240)
        out[rank] = res
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0]["block"] == (0, 19) and out[1]["block"] == (19, 37)
    assert out[0]["m_shape"] == (24,)
    H = xt.HostArray.from_numpy
    for r in (0, 1):
        assert np.array_equal(out[r]["sum0"], full.sum(axis=0))           # integer-valued: exact in any order
        assert np.array_equal(out[r]["max0"], full.max(axis=0))
        assert np.array_equal(out[r]["mean0"], xt.evaluate(xt.mean(H(full), [0], dtype=xt.F32)).numpy())
        ref_var = xt.evaluate(xt.variance(H(full), [0], dtype=xt.F32)).numpy()
        assert np.allclose(out[r]["var0"], ref_var, rtol=1e-6, atol=0)
    b0, e0 = out[0]["block"]
    assert np.array_equal(np.concatenate([out[0]["sum1"], out[1]["sum1"]]), full.sum(axis=1))
    whole = xt.evaluate(xt.exp(H(full) - H(mvec))).numpy()
    assert np.array_equal(np.concatenate([out[0]["map"], out[1]["map"]]), whole)


class _FakeLib:
    """Stands in for libxtb200 in the rendezvous test: records the calls, rank `fail_rank` cannot map the windows."""

    def __init__(self, rank, fail_rank):
        self.rank, self.fail_rank, self.calls = rank, fail_rank, []

    def xtb_comm_unique_id(self, buf): return 0
    def xtb_comm_init(self, rank, world, ident): return 0
    def xtb_comm_p2p_handle(self, h): return 0
    def xtb_last_error(self): return b"mock"

    def xtb_comm_p2p_attach(self, handles, world):
        self.calls.append(world)
        return 1 if (world > 0 and self.rank == self.fail_rank) else 0


def _init_comm_worker(rank, world, port, fail_rank, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch.distributed as dist
    from xtensor_b200 import capi, shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fake = _FakeLib(rank, fail_rank)
    capi.lib = lambda: fake
    attached = shard.init_comm(dist, rank, world)
    q.put((rank, attached, fake.calls))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("fail_rank", [-1, 1])
def test_init_comm_is_collective(fail_rank):
    """shard.init_comm: the peer-memory windows are attached on every rank or on none (a rank that cannot map them
    makes all ranks detach and fall back to NCCL) -- otherwise the ranks would wait for each other on different routes."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_init_comm_worker, args=(r, 2, port, fail_rank, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if fail_rank < 0:
        assert res == [(0, True, [2]), (1, True, [2])]
    else:
        assert res == [(0, False, [2, 0]), (1, False, [2, 0])]
