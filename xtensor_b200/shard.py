"""Leading-axis sharding of the hot path across GPUs (one process per GPU).

The reference has no distributed layer; its own partial -> merge -> finalize contract is
xblockwise_reducer (reducers/xblockwise_reducer.hpp:154-185, functors
xblockwise_reducer_functors.hpp:45-260).  Here:
  * elementwise assignment and reductions over non-sharded axes need no exchange: every
    rank evaluates its row block with the single-GPU kernels;
  * a reduction over the sharded axis 0 produces one partial per rank, merged with ONE
    allreduce (sum / prod / max / min) -- `allreduce` is NCCL through the C ABI on the
    device (xtb_reduce(..., allreduce=1)), or any callable in tests (gloo on host arrays);
  * mean finalises after the merge (divide by the GLOBAL count, cf. mean_functor::finalize,
    xblockwise_reducer_functors.hpp:175-185); variance is the reference's two-pass form with
    the merged mean broadcast back to every rank.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import numpy as np

from . import capi
from . import expr as xt


def row_block(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of this rank's contiguous block of the leading axis (remainder to low ranks)."""
    base, rem = divmod(n_rows, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_operand(a: xt.Array, out_rows: int, rank: int, world: int) -> xt.Array:
    """Slice `a` on axis 0 if it spans the sharded axis, else replicate it (broadcast operand)."""
    if a.ndim == 0 or a.shape[0] != out_rows or out_rows == 1:
        return a
    b, e = row_block(out_rows, rank, world)
    return a[b:e]


def init_comm(dist, rank: int, world: int, p2p: Optional[bool] = None) -> bool:
    """Bring up the device-side exchange for an initialised torch.distributed group `dist` (used for
    the rendezvous only): NCCL communicator through the C ABI, then -- unless `p2p` is False or
    XTB_NO_P2P is set -- the NVLink peer-memory windows that serve small allreduces (the (cols,)
    partials of an axis-0 reduction) with one kernel instead of an NCCL call.  Returns whether the
    peer-memory path is attached."""
    import ctypes as C
    import os
    lib = capi.lib()
    ident = [None]
    if rank == 0:
        buf = C.create_string_buffer(128)
        capi.check(lib.xtb_comm_unique_id(buf))
        ident[0] = buf.raw
    dist.broadcast_object_list(ident, src=0)
    capi.check(lib.xtb_comm_init(rank, world, C.create_string_buffer(ident[0], 128)))
    if p2p is None:
        p2p = os.environ.get("XTB_NO_P2P") is None
    if not p2p or world > 8:
        return False
    # every step is collective: a rank that cannot export or map a window makes ALL ranks fall back to NCCL
    h = C.create_string_buffer(64)
    ok = lib.xtb_comm_p2p_handle(h) == 0
    handles = [None] * world
    dist.all_gather_object(handles, h.raw if ok else None)
    if all(x is not None for x in handles):
        ok = lib.xtb_comm_p2p_attach(C.create_string_buffer(b"".join(handles), 64 * world), world) == 0
    else:
        ok = False
    oks = [None] * world
    dist.all_gather_object(oks, bool(ok))      # also the barrier: every rank attached before anyone's next allreduce
    if not all(oks):
        capi.check(lib.xtb_comm_p2p_attach(None, 0))
        return False
    return True


HostAllreduce = Callable[[np.ndarray, int], np.ndarray]


def sharded_reduce(op: int, local: xt.Expr, axes: Sequence[int], global_rows: int, world: int,
                   host_allreduce: Optional[HostAllreduce] = None, dtype=None) -> xt.Array:
    """Reduce a row-sharded expression.  `local` is this rank's block; if axis 0 is reduced the
    per-rank partials are merged by one allreduce, otherwise the result stays sharded."""
    r = xt.Reducer(op, local, list(axes), acc_dtype=dtype)
    kind = xt._leaf_kind(local) or xt.DeviceArray
    crosses = 0 in r.axes and world > 1
    if kind is xt.DeviceArray:
        return xt._run_reducer(r, kind, allreduce=crosses)
    part = xt._run_reducer(r, kind)
    if crosses:
        if host_allreduce is None:
            raise RuntimeError("host arrays need a host_allreduce callable")
        merged = host_allreduce(part.numpy(), op)
        return xt.HostArray.from_numpy(merged)
    return part


def sharded_mean(local: xt.Expr, axes: Sequence[int], global_rows: int, world: int,
                 host_allreduce: Optional[HostAllreduce] = None, dtype=None) -> xt.Array:
    """mean over `axes` of a row-sharded expression: merged sum / GLOBAL count."""
    full_shape = (global_rows,) + tuple(local.shape[1:])
    n = int(np.prod([full_shape[a] for a in axes], dtype=np.int64))
    vt = xt.F64 if dtype is None else dtype
    kind = xt._leaf_kind(local) or xt.DeviceArray
    if kind is xt.DeviceArray:
        # device: local partial, cross-GPU merge and the division by the GLOBAL count are one call -- the
        # division is the epilogue of the merge kernel (xtb_reduce_fin), applied once after the exchange
        r = xt.Reducer(capi.RED_SUM, local, list(axes), acc_dtype=dtype)
        out = xt.DeviceArray.empty(r.shape, vt)
        fin = capi.Finalize(capi.FIN_DIV, vt, xt._imm_bits(xt.NP_OF[vt](n), vt))
        return xt._run_reducer(r, kind, allreduce=(0 in r.axes and world > 1), out=out, fin=fin)
    s = sharded_reduce(capi.RED_SUM, local, axes, global_rows, world, host_allreduce, dtype)
    return xt.evaluate(s / xt.Scalar(xt.NP_OF[vt](n), vt))


def sharded_variance(local: xt.Expr, axes: Sequence[int], global_rows: int, world: int,
                     host_allreduce: Optional[HostAllreduce] = None, dtype=None) -> xt.Array:
    """Two-pass variance (core/xmath.hpp:2082-2105) on a row-sharded expression."""
    m = sharded_mean(local, axes, global_rows, world, host_allreduce, dtype)
    keep = [1 if d in axes else s for d, s in enumerate(local.shape)]
    mrv = m.reshape_view(keep)
    return sharded_mean(xt.square(local - mrv), axes, global_rows, world, host_allreduce, dtype)
