"""Multi-GPU check of the exchange step (run under torchrun, one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dist_check.py
Covers xtb_allreduce on both routes (NVLink peer-memory kernel for <= 256 KB, NCCL above), every
32-/64-bit dtype and merge op, back-to-back calls (window slot reuse), calls from a forked stream and
from a replayed CUDA graph, and the sharded mean / variance / map of cfg5 against numpy on the whole
matrix.  Prints one JSON line on rank 0; exits non-zero on any mismatch."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xtensor_b200 import capi, shard  # noqa: E402
from xtensor_b200 import expr as xt  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    lib = capi.lib()
    capi.check(lib.xtb_init(local))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p2p = shard.init_comm(dist, rank, world)
    fails, checks = [], 0

    NP = {xt.F32: np.float32, xt.F64: np.float64, xt.I32: np.int32, xt.I64: np.int64, xt.U32: np.uint32, xt.U64: np.uint64}
    FN = {capi.RED_SUM: np.add, capi.RED_PROD: np.multiply, capi.RED_MAX: np.maximum, capi.RED_MIN: np.minimum}

    def part(r, n, dt, op):
        g = np.random.default_rng(1000 * r + n % 977)
        if op == capi.RED_PROD:
            return g.integers(1, 3, n).astype(NP[dt])
        lo = 0 if NP[dt] in (np.uint32, np.uint64) else -8
        return g.integers(lo, 9, n).astype(NP[dt])

    def want(n, dt, op):
        acc = part(0, n, dt, op)
        for r in range(1, world):
            acc = FN[op](acc, part(r, n, dt, op))
        return acc

    def allreduce_dev(d, n, dt, op):
        capi.check(lib.xtb_allreduce(C.c_void_p(d.owner.ptr), n, dt, op))

    for dt in NP:
        for op in FN:
            for n in (1, 7, 1000, 8192, 32768, 65536, 300000):
                d = xt.DeviceArray.from_numpy(part(rank, n, dt, op))
                allreduce_dev(d, n, dt, op)
                got = d.numpy()
                checks += 1
                if not np.array_equal(got, want(n, dt, op)):
                    fails.append(f"allreduce dtype {dt} op {op} n {n}")
    kernel_small = None
    d = xt.DeviceArray.from_numpy(part(rank, 8192, xt.F32, capi.RED_SUM))
    allreduce_dev(d, 8192, xt.F32, capi.RED_SUM)
    kernel_small = lib.xtb_last_kernel().decode()

    # back-to-back calls reuse the two payload slots; uneven arrival (rank-dependent delay)
    n = 8192
    for it in range(200):
        x = np.full(n, rank + it, np.float32)
        d = xt.DeviceArray.from_numpy(x)
        if it % 7 == rank % 7:
            torch.cuda._sleep(2_000_000)
        allreduce_dev(d, n, xt.F32, capi.RED_SUM)
        if it % 20 == 0 or it > 190:
            checks += 1
            if not np.array_equal(d.numpy(), np.full(n, sum(range(world)) + world * it, np.float32)):
                fails.append(f"back-to-back it {it}")

    # cfg5-shaped pipeline, small: mean / variance over the sharded axis + map, eager, then forked + graph
    rows, cols = 64 * world + 3, 8192
    full = np.random.default_rng(5).integers(-8, 9, (rows, cols)).astype(np.float32)
    b, e = shard.row_block(rows, rank, world)
    a = xt.DeviceArray.from_numpy(full[b:e])
    m = shard.sharded_mean(a, [0], rows, world, dtype=xt.F32).numpy()
    v = shard.sharded_variance(a, [0], rows, world, dtype=xt.F32).numpy()
    m_ref = full.sum(axis=0, dtype=np.float32) / np.float32(rows)       # integer-valued: sum exact in any order
    checks += 2
    if not np.array_equal(m, m_ref):
        fails.append("sharded mean")
    v_ref = np.square(full.astype(np.float64) - m_ref.astype(np.float64)).sum(axis=0) / rows
    if not np.allclose(v, v_ref, rtol=1e-6, atol=0):
        fails.append("sharded variance")

    # split reductions over the sharded axis: local merge + cross-GPU merge in one kernel
    fused_kernel = None
    for npdt in (np.float32, np.float64, np.int32, np.int64):
        for name, fn in (("sum", np.sum), ("amax", np.max), ("amin", np.min)):
            for cols_ in (1000, 8192, 131):
                rows_ = 256 * world + 5
                fl = np.random.default_rng(cols_).integers(-8, 9, (rows_, cols_)).astype(npdt)
                b_, e_ = shard.row_block(rows_, rank, world)
                got = xt._run_reducer(getattr(xt, name)(xt.DeviceArray.from_numpy(fl[b_:e_]), [0]), xt.DeviceArray, allreduce=True)
                if npdt is np.float32 and name == "sum" and cols_ == 8192:
                    fused_kernel = lib.xtb_last_kernel().decode()
                checks += 1
                if not np.array_equal(got.numpy().astype(np.float64), fn(fl, axis=0).astype(np.float64)):
                    fails.append(f"sharded {name} {npdt.__name__} cols {cols_}")

    # reductions the planner rewrites into two passes (narrow: < 1024 outputs; mixed: outer + innermost axis): the
    # cross-GPU merge, xt::initial and the finalize step belong to the second pass
    for shape_, axes_ in (((32768 * world + 7, 48), [0]), ((300 * world + 1, 70, 50, 6), [0, 2]), ((2048 * world, 64, 16), [0, 2])):
        fl = np.random.default_rng(len(shape_)).integers(-8, 9, shape_).astype(np.float32)
        b_, e_ = shard.row_block(shape_[0], rank, world)
        loc = xt.DeviceArray.from_numpy(fl[b_:e_])
        got = xt._run_reducer(xt.sum(loc, axes_, initial=np.float32(10)), xt.DeviceArray, allreduce=True)
        checks += 1
        if not np.array_equal(got.numpy(), fl.sum(axis=tuple(axes_), dtype=np.float64).astype(np.float32) + np.float32(10)):
            fails.append(f"sharded two-pass sum {shape_} {axes_}")
        got = xt._run_reducer(xt.amax(loc, axes_), xt.DeviceArray, allreduce=True)
        checks += 1
        if not np.array_equal(got.numpy(), fl.max(axis=tuple(axes_))):
            fails.append(f"sharded two-pass amax {shape_} {axes_}")

    s_sum, mean_, s_sq, var_ = (xt.DeviceArray.empty((cols,), xt.F32) for _ in range(4))
    o = xt.DeviceArray.empty((e - b, cols), xt.F32)
    n_rows = np.float32(rows)

    def pipeline():
        xt._run_reducer(xt.sum(a, [0]), xt.DeviceArray, allreduce=True, out=s_sum)
        xt.assign(mean_, s_sum / n_rows)
        capi.check(lib.xtb_fork_begin())
        xt._run_reducer(xt.sum(xt.square(a - mean_), [0]), xt.DeviceArray, allreduce=True, out=s_sq)
        xt.assign(var_, s_sq / n_rows)
        capi.check(lib.xtb_fork_end())
        xt.assign(o, xt.exp(a - mean_))
        capi.check(lib.xtb_fork_join())

    pipeline()
    capi.check(lib.xtb_sync())
    graph = C.c_void_p()
    capi.check(lib.xtb_graph_begin())
    pipeline()
    capi.check(lib.xtb_graph_end(C.byref(graph)))
    for _ in range(50):
        capi.check(lib.xtb_graph_launch(graph))
    capi.check(lib.xtb_sync())
    checks += 3
    if not np.array_equal(mean_.numpy(), m_ref):
        fails.append("graph mean")
    if not np.allclose(var_.numpy(), v_ref, rtol=1e-6, atol=0):
        fails.append("graph variance")
    map_ref = np.exp((full[b:e] - m_ref).astype(np.float64))
    if not np.allclose(o.numpy(), map_ref, rtol=3e-7, atol=0):
        fails.append("graph map")
    # every rank holds the same bits
    t = torch.from_numpy(var_.numpy().view(np.int32).copy()).cuda()
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    checks += 1
    if not torch.equal(lo, hi):
        fails.append("variance differs between ranks")
    lib.xtb_graph_destroy(graph)

    nf = torch.tensor([len(fails)], device="cuda")
    dist.all_reduce(nf)
    if fails:
        print(f"rank {rank} FAILED: {fails}", flush=True)
    if rank == 0:
        print(json.dumps({"dist_check": "ok" if int(nf.item()) == 0 else "FAILED", "world": world, "checks_per_rank": checks,
                          "peer_memory_allreduce": p2p, "small_allreduce_kernel": kernel_small,
                          "split_reduce_kernel": fused_kernel}), flush=True)
    dist.barrier()
    lib.xtb_comm_destroy()
    dist.destroy_process_group()
    sys.exit(1 if int(nf.item()) else 0)


if __name__ == "__main__":
    main()
