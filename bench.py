#!/usr/bin/env python
"""bench.py -- headline measurement of the xtensor hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): effective HBM GB/s of the fused broadcast assign + axis reductions of
BASELINE cfg5, the configuration north_star's multi-GPU clause is written for:

    a: fp32 (262144, 8192), sharded by rows over the N GPUs (STRONG scaling: 262144 / N rows per GPU)
    m   = xt::mean<float>(a, {0})                 pass 1: read a          (+ cross-GPU merge of the partials)
    v   = xt::variance<float>(a, {0})             pass 2: read a          (two-pass form, xmath.hpp:2082-2105;
                                                                           + cross-GPU merge)
    out = exp(a - m)                              pass 3: read a, write out (fused broadcast assign)

One "step" = that pipeline once.  Algorithmic bytes: every distinct input element read once
per pass, every output element written once = 4 x 8 GiB + 3 x 32 KiB = 34,359,836,672 B for the whole job at
every N.  `value` = those bytes / device time (CUDA events, max over ranks) = aggregate GB/s over all GPUs.

The step is recorded once into a CUDA graph through the C ABI (xtb_graph_*): five kernels -- the two
reductions with the division by the GLOBAL row count fused into their merge kernel (xtb_reduce_fin), the
cross-GPU exchange of the per-GPU partials fused into that same merge kernel over NVLink peer memory, and
the map -- with the variance chain on a forked stream so that its exchange hides behind the map.

Correctness is asserted IN the run, at every N, through the same graph: on an integer-valued tiled matrix
the mean must equal the exact value bit for bit on every rank, the variance must be within 1e-6 of the
fp64 two-pass value, every rank must hold identical bits, and the map must be within 2 ulp of fp64.
A failed check exits non-zero and prints no JSON line.

The JSON line also carries: roofline (the dominant kernel, the map, timed alone with CUDA events against the
measured copy peak), cpu_baseline / --impl reference (the REAL xtensor evaluating the same pipeline on a
bounded sample on the host cores), e2e (the same step from pinned HOST buffers: H2D of the shard and D2H of
out / mean / variance inside the timed region), clocks, gpu_launches, and the other BASELINE configs.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS, COLS, BLK = 262144, 8192, 4096
WORKLOAD = ("cfg5: fp32 a(262144,8192) sharded by rows: m = mean<float>(a,{0}), v = variance<float>(a,{0}) "
            "(two-pass), out = exp(a - m); per-GPU partials merged across GPUs")
METRIC = "effective HBM GB/s, fused broadcast assign + axis reductions (algorithmic bytes / device time)"
CONFIG = {"workload": WORKLOAD, "rows": ROWS, "cols": COLS,
          "l2": "every pass streams its whole shard (>= 1 GiB per GPU, larger than the 126 MB L2): no flush needed"}
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def job_bytes(rows=ROWS, cols=COLS):
    return 4 * rows * cols * 4 + 3 * cols * 4


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---- clocks ------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.dev, self.proc, self.lines = device_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.dev)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- device timing -----------------------------------------------------------------------------
def device_time_ms(lib, fn, iters, lead_in=0):
    """CUDA-event time of `iters` calls of fn on the library stream.  lead_in > 0 enqueues that many
    untimed calls right before the start event (no host sync in between): for a step that contains a
    cross-GPU exchange this lines the ranks' device queues up, so the timed region does not include the
    skew with which the host processes left the barrier."""
    from xtensor_b200 import capi
    e0, e1 = C.c_void_p(), C.c_void_p()
    capi.check(lib.xtb_event_create(C.byref(e0)))
    capi.check(lib.xtb_event_create(C.byref(e1)))
    capi.check(lib.xtb_sync())
    for _ in range(lead_in):
        fn()
    capi.check(lib.xtb_event_record(e0))
    for _ in range(iters):
        fn()
    capi.check(lib.xtb_event_record(e1))
    ms = C.c_float()
    capi.check(lib.xtb_event_elapsed_ms(e0, e1, C.byref(ms)))
    capi.check(lib.xtb_sync())
    lib.xtb_event_destroy(e0); lib.xtb_event_destroy(e1)
    return float(ms.value)


def fail(msg):
    print(f"bench.py: CHECK FAILED: {msg}", file=sys.stderr, flush=True)
    sys.exit(3)


# ---- the cfg5 step -------------------------------------------------------------------------------
class Cfg5:
    """The sharded pipeline on this rank's row block, through the C ABI (ctypes mirror of the expression API)."""

    def __init__(self, lib, xt, capi, rows, cols, total_rows, world):
        self.lib, self.xt, self.capi = lib, xt, capi
        self.rows, self.cols, self.total_rows, self.world = rows, cols, total_rows, world
        self.a = xt.DeviceArray.empty((rows, cols), xt.F32)
        self.o = xt.DeviceArray.empty((rows, cols), xt.F32)
        self.mean = xt.DeviceArray.empty((cols,), xt.F32)
        self.var = xt.DeviceArray.empty((cols,), xt.F32)
        n = np.float32(total_rows)
        self.fin = capi.Finalize(capi.FIN_DIV, xt.F32, xt._imm_bits(n, xt.F32))
        self.overlap = world > 1 and os.environ.get("XTB_BENCH_NO_FORK") is None
        self.graph = None

    def fill(self, blk):
        """a = blk tiled along the rows (H2D block by block)."""
        lib, capi = self.lib, self.capi
        assert self.rows % blk.shape[0] == 0 and blk.shape[1] == self.cols and blk.dtype == np.float32
        for r0 in range(0, self.rows, blk.shape[0]):
            capi.check(lib.xtb_memcpy(C.c_void_p(self.a.owner.ptr + r0 * self.cols * 4), C.c_void_p(blk.ctypes.data), blk.nbytes, capi.H2D))
        capi.check(lib.xtb_sync())

    # the three data passes, also timed one by one for the roofline
    def pass_mean(self):
        xt = self.xt
        xt._run_reducer(xt.sum(self.a, [0]), xt.DeviceArray, allreduce=self.world > 1, out=self.mean, fin=self.fin)

    def pass_var(self):
        xt = self.xt
        xt._run_reducer(xt.sum(xt.square(self.a - self.mean), [0]), xt.DeviceArray, allreduce=self.world > 1, out=self.var, fin=self.fin)

    def pass_map(self):
        self.xt.assign(self.o, self.xt.exp(self.a - self.mean))

    def pipeline(self):
        lib, capi = self.lib, self.capi
        self.pass_mean()
        # variance and map both need only the mean: on several GPUs the variance chain (kernel, merge + exchange
        # + finalize) runs on the forked stream so that its exchange hides behind the map
        if self.overlap:
            capi.check(lib.xtb_fork_begin())
        self.pass_var()
        if self.overlap:
            capi.check(lib.xtb_fork_end())
        self.pass_map()
        if self.overlap:
            capi.check(lib.xtb_fork_join())

    def capture(self):
        lib, capi = self.lib, self.capi
        for _ in range(2):
            self.pipeline()        # sizes internal scratch, compiles nothing (ahead-of-time kernels)
        capi.check(lib.xtb_sync())
        if os.environ.get("XTB_BENCH_NO_GRAPH") is not None:
            return
        g = C.c_void_p()
        capi.check(lib.xtb_graph_begin())
        self.pipeline()
        capi.check(lib.xtb_graph_end(C.byref(g)))
        self.graph = g

    def step(self):
        if self.graph is not None:
            self.capi.check(self.lib.xtb_graph_launch(self.graph))
        else:
            self.pipeline()

    def close(self):
        if self.graph is not None:
            self.lib.xtb_graph_destroy(self.graph)
            self.graph = None


def int_block(rank):
    return np.random.default_rng(900 + rank).integers(-8, 9, (BLK, COLS)).astype(np.float32)


def uni_block(rank):
    return np.random.default_rng(9 + rank).uniform(-1, 1, (BLK, COLS)).astype(np.float32)


def host_moments(blk, rows_per_rank, dist, m32=None):
    """fp64 column sums, sums of |x| and -- given the fp32 mean the device used -- sums of squared deviations of the
    WHOLE tiled matrix: every rank reduces its own generating block on the host, the ranks' fp64 partials are
    added with one torch.distributed all_reduce (test plumbing, not the measured path)."""
    reps = rows_per_rank // BLK
    b = blk.astype(np.float64)
    parts = [reps * b.sum(axis=0), reps * np.abs(b).sum(axis=0)]
    if m32 is not None:
        parts.append(reps * np.square(b - m32.astype(np.float64)).sum(axis=0))
    acc = np.stack(parts)
    if dist is not None:
        import torch
        t = torch.from_numpy(acc).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        acc = t.cpu().numpy()
    return acc


def verify(step: Cfg5, blk_of, rank, world, dist, exact, what):
    """Run one step on `blk_of(rank)` tiled and check mean / variance / map; all ranks must agree bit for bit."""
    xt, lib, capi = step.xt, step.lib, step.capi
    blk = blk_of(rank)
    step.fill(blk)
    step.step()
    capi.check(lib.xtb_sync())
    m, v = step.mean.numpy(), step.var.numpy()
    n = step.total_rows
    s64, sabs, sq64 = host_moments(blk, step.rows, dist, m32=m)
    if exact:
        # integers: |sum| <= 8 * 262144 < 2^24, exact in any order; the fp32 division is correctly rounded
        want_m = s64.astype(np.float32) / np.float32(n)
        if not np.array_equal(m, want_m):
            fail(f"{what}: mean differs from the exact value (max |diff| {np.abs(m - want_m).max():.3e})")
    else:
        err = np.abs(m.astype(np.float64) - s64 / n) / (sabs / n)
        if float(err.max()) > 1e-6:
            fail(f"{what}: mean error {err.max():.3e} of sum|x|/N exceeds 1e-6")
    v64 = sq64 / n
    verr = np.abs(v.astype(np.float64) - v64) / v64
    if float(verr.max()) > 1e-6:
        fail(f"{what}: variance relative error vs fp64 two-pass {verr.max():.3e} exceeds 1e-6")
    # the map: first generating block of this rank against fp64, <= 2 ulp of the fp32 result
    got = np.empty((BLK, COLS), np.float32)
    capi.check(lib.xtb_memcpy(C.c_void_p(got.ctypes.data), C.c_void_p(step.o.owner.ptr), got.nbytes, capi.D2H))
    arg32 = blk - m
    want = np.exp(arg32.astype(np.float64))
    ulps = np.abs(got.astype(np.float64) - want) / np.spacing(want.astype(np.float32)).astype(np.float64)
    if float(ulps.max()) > 2.0:
        fail(f"{what}: exp(a - mean) is {ulps.max():.2f} ulp from fp64")
    last = np.empty((BLK, COLS), np.float32)
    off = (step.rows - BLK) * COLS * 4
    capi.check(lib.xtb_memcpy(C.c_void_p(last.ctypes.data), C.c_void_p(step.o.owner.ptr + off), last.nbytes, capi.D2H))
    if not np.array_equal(got, last):
        fail(f"{what}: the map of the last tile differs from the first (same input block)")
    if dist is not None:
        import torch
        mine = torch.from_numpy(np.concatenate([m, v]).view(np.int32).copy()).cuda()
        allv = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        for r in range(world):
            if not torch.equal(allv[r], allv[0]):
                fail(f"{what}: rank {r} holds different mean / variance bits than rank 0")
    return {"mean_exact": bool(exact), "variance_max_rel_err_vs_fp64": float(verr.max()), "map_max_ulp": float(ulps.max()),
            "ranks_bit_identical": True}


# ---- our arm -----------------------------------------------------------------------------------
def run_ours(args):
    from xtensor_b200 import capi
    from xtensor_b200 import expr as xt
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun for --gpus > 1 (one process per GPU)")
    lib = capi.lib()
    capi.check(lib.xtb_init(local_rank))
    dist, p2p_on = None, False
    if world > 1:
        import torch
        import torch.distributed as dist_
        dist = dist_
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from xtensor_b200 import shard
        p2p_on = shard.init_comm(dist, rank, world)

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    total_rows = args.rows or ROWS
    rows = total_rows // world
    warmup = max(args.warmup, 3)
    step = Cfg5(lib, xt, capi, rows, COLS, total_rows, world)
    step.fill(uni_block(rank))
    step.capture()

    # ---- correctness, through the same graph, before anything is timed --------------------------------
    checks = {"integer_tiled": verify(step, int_block, rank, world, dist, True, "integer-valued matrix")}

    # ---- headline: K steps on the device -------------------------------------------------------------
    step.fill(uni_block(rank))
    for _ in range(warmup):
        step.step()
    capi.check(lib.xtb_sync())
    sampler = ClockSampler(local_rank)
    barrier()
    lib.xtb_launch_count(1)
    sampler.start()
    ms = device_time_ms(lib, step.step, args.steps, lead_in=2 if world > 1 else 0)
    clocks = sampler.stop()
    launches = int(lib.xtb_launch_count(0)) * args.steps // (args.steps + (2 if world > 1 else 0))
    barrier()
    ms_per_step = max_over_ranks(ms) / args.steps
    nbytes = job_bytes(total_rows, COLS)
    value = nbytes / (ms_per_step * 1e-3) / 1e9
    checks["timed_data"] = verify(step, uni_block, rank, world, dist, False, "U(-1,1) matrix (the timed data)")

    # ---- roofline: the dominant kernel (the map: 2 of the 4 passes) timed alone, live ---------------------
    peak, peak_src = peak_hbm()
    kern = {}
    for name, fn, b in (("mean_pass", step.pass_mean, rows * COLS * 4 + COLS * 4),
                        ("variance_pass", step.pass_var, rows * COLS * 4 + 2 * COLS * 4),
                        ("map", step.pass_map, 2 * rows * COLS * 4 + COLS * 4)):
        for _ in range(3):
            fn()
        n_it = 20
        t = max_over_ranks(device_time_ms(lib, fn, n_it)) / n_it
        kern[name] = {"ms": round(t, 5), "GBs": round(b / t / 1e6, 1), "frac_of_measured_peak": round(b / t / 1e6 / peak, 4),
                      "kernel": lib.xtb_last_kernel().decode(), "algorithmic_bytes_per_launch": b}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1 and total_rows == ROWS:
        try:
            traffic = json.load(open(tpath)).get("cfg5_map_dram_bytes_per_launch")
        except Exception:
            traffic = None
    mk = kern["map"]
    roofline = {"bound": "hbm", "achieved": mk["GBs"], "peak": peak, "unit": "GB/s", "frac": mk["frac_of_measured_peak"],
                "traffic": traffic, "traffic_source": "ncu --set full, offline (profiles/traffic.json)" if traffic else None,
                "kernel": mk["kernel"], "algorithmic_bytes_per_launch": mk["algorithmic_bytes_per_launch"],
                "launch_ms": mk["ms"], "peak_source": peak_src, "step_frac_of_measured_peak_per_gpu": round(value / world / peak, 4),
                "kernels": kern}

    # ---- end to end: the same step from pinned HOST buffers ----------------------------------------------
    e2e = run_e2e(args, lib, xt, capi, step, rank, world, barrier, max_over_ranks, nbytes)

    extra = {}
    step_info = {"rows_per_gpu": rows, "cuda_graph": step.graph is not None, "graph_kernels": int(lib.xtb_graph_kernel_count(step.graph)) if step.graph else None,
                 "variance_overlaps_map": step.overlap,
                 "exchange": ("k_reduce_merge + NVLink peer-memory exchange (one kernel)" if p2p_on else "nccl allreduce") if world > 1 else None}
    step.close()
    del step
    # ---- the same step and the same end-to-end loop through the DROP-IN C++ header (tests/cpp/bench_dropin.cpp) ----
    cpp = run_cpp_dropin(args, total_rows, rank, world)
    barrier()
    if rank == 0 and cpp is not None and "e2e_value" in cpp:
        # the headline end-to-end number is the one a C++ user of the drop-in gets; the ctypes loop stays beside it
        e2e = dict(e2e, ctypes_value=e2e["value"], ctypes_ms_per_step=e2e["ms_per_step"], value=cpp["e2e_value"],
                   ms_per_step=cpp["e2e_ms_per_step"], steps=cpp["e2e_steps"],
                   how="tests/cpp/bench_dropin.cpp, one process per GPU: xtb::copy_to_device(a, pinned host a); m = mean<float>(a,{0}) / "
                       "v = variance<float>(a,{0}) through xtb::dist::*_into; xt::noalias(out) = xt::exp(a - m); xtb::copy_to_host of "
                       "out / mean / variance into pinned host memory; wall clock, max over ranks; bytes per step are per rank. "
                       "ctypes_value = the same loop through the ctypes mirror in this process")
    if not args.no_extra and not args.quick and world == 1:
        extra = other_configs(lib, xt, capi, args)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": CONFIG,
            "pct_of_8TBs_per_gpu": round(100 * value / world / 8000.0, 2),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "checks": checks, "step": step_info,
            "cpp_dropin": cpp,
        }
        if total_rows != ROWS:
            line["config"] = dict(CONFIG, rows=total_rows, note="development run on a reduced row count")
        if world == 1:
            line["cpu_baseline"] = cpu_baseline_cfg5(args.cpu_sample_rows)
        line.update(extra)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        lib.xtb_comm_destroy()
        dist.destroy_process_group()


def run_cpp_dropin(args, total_rows, rank, world):
    """Run tests/cpp/_build/bench_dropin (the step written against include/xtb200/xtensor_b200.hpp) as this rank's
    child: the children rendezvous among themselves (xtb::dist::init_from_env, MASTER_PORT + 29).  Rank 0's child
    prints the JSON object that is returned; None when the binary is absent (it is built where /root/reference is)."""
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "bench_dropin")
    if not os.path.exists(exe):
        return {"error": "tests/cpp/_build/bench_dropin not built"} if rank == 0 else None
    n_e2e = 1 if args.quick else max(2, min(args.steps, 5))
    try:
        r = subprocess.run([exe, str(total_rows), str(args.steps), str(max(args.warmup, 3)), str(n_e2e)],
                           capture_output=True, text=True, timeout=900)
    except Exception as ex:
        return {"error": repr(ex)}
    if rank != 0:
        return None
    for ln in r.stdout.splitlines():
        if ln.startswith("{"):
            try:
                return json.loads(ln)
            except Exception:
                pass
    return {"error": f"rc={r.returncode} {r.stdout[-300:]} {r.stderr[-300:]}"}


def run_e2e(args, lib, xt, capi, step, rank, world, barrier, max_over_ranks, nbytes):
    """The public call with HOST buffers: every step copies this rank's shard from pinned host memory to the
    device, runs the pipeline, and reads out / mean / variance back into pinned host memory."""
    rows, cols = step.rows, step.cols
    shard_bytes = rows * cols * 4
    bufs = []
    try:
        for nb in (shard_bytes, shard_bytes, cols * 4, cols * 4):
            p = C.c_void_p()
            capi.check(lib.xtb_host_alloc(nb, C.byref(p)))
            bufs.append(p)
    except Exception as ex:
        for p in bufs:
            lib.xtb_host_free(p)
        return {"value": None, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(ex)}
    h_a, h_o, h_m, h_v = bufs
    blk = uni_block(rank)
    for r0 in range(0, rows, BLK):
        C.memmove(h_a.value + r0 * cols * 4, blk.ctypes.data, blk.nbytes)

    def e2e_step():
        capi.check(lib.xtb_memcpy(C.c_void_p(step.a.owner.ptr), h_a, shard_bytes, capi.H2D))
        step.step()
        capi.check(lib.xtb_memcpy(h_m, C.c_void_p(step.mean.owner.ptr), cols * 4, capi.D2H))   # D2H blocks: host sees the results
        capi.check(lib.xtb_memcpy(h_v, C.c_void_p(step.var.owner.ptr), cols * 4, capi.D2H))
        capi.check(lib.xtb_memcpy(h_o, C.c_void_p(step.o.owner.ptr), shard_bytes, capi.D2H))

    n = 1 if args.quick else max(2, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(n):
        e2e_step()
    dt = max_over_ranks((time.perf_counter() - t0) / n)
    res = np.ctypeslib.as_array(C.cast(h_o, C.POINTER(C.c_float)), shape=(rows * cols,))
    dev = np.empty(BLK * cols, np.float32)        # the device result of the last tile, read on its own
    capi.check(lib.xtb_memcpy(C.c_void_p(dev.ctypes.data), C.c_void_p(step.o.owner.ptr + (rows - BLK) * cols * 4), dev.nbytes, capi.D2H))
    ok = bool(np.array_equal(res[(rows - BLK) * cols:], dev)) and bool(np.array_equal(res[: BLK * cols], dev))
    out = {"value": round(nbytes / dt / 1e9, 2), "unit": "GB/s", "h2d_bytes_per_step": int(shard_bytes),
           "d2h_bytes_per_step": int(shard_bytes + 2 * cols * 4), "steps": n, "ms_per_step": round(dt * 1e3, 3),
           "matches_device_result": ok, "result_checksum": float(res[:: 4099].astype(np.float64).sum()),
           "how": "per rank: xtb_memcpy H2D of the shard from pinned host memory, the graph, xtb_memcpy D2H of out / mean / variance; "
                  "wall clock, max over ranks; bytes per step are per rank"}
    for p in bufs:
        lib.xtb_host_free(p)
    if not ok:
        fail("e2e: the host result differs from the device result")
    return out


# ---- the reference's CPU evaluation ----------------------------------------------------------------
def cpu_cfg5_sample(sample_rows, reps=1):
    """The REAL xtensor (oracle/_ref/libxtref_fast.so, prebuilt from /root/reference's headers) evaluating the same
    pipeline on the first `sample_rows` rows: xt::mean<float>(a,{0}), xt::variance<float>(a,{0}), exp(a - m).
    Falls back to the oracle restatement (kind "port") when the library is absent."""
    a = np.tile(uni_block(0), (max(1, sample_rows // BLK), 1))[:sample_rows]
    nb = job_bytes(sample_rows, COLS)
    try:
        from oracle import refbin
        if refbin.available(fast=True):
            L = refbin.lib(True)
            sh, ax = (C.c_int64 * 2)(sample_rows, COLS), (C.c_int32 * 1)(0)
            m, v, o = np.empty(COLS, np.float32), np.empty(COLS, np.float32), np.empty_like(a)
            p = lambda x: C.c_void_p(x.ctypes.data)
            best = None
            for _ in range(reps):
                t0 = time.perf_counter()
                L.xtref_mean_f32_f32(p(a), 2, sh, 1, ax, p(m))
                L.xtref_variance_f32_f32(p(a), 2, sh, 1, ax, p(v))
                L.xtref_cfg5_exp_sub_f32(p(a), p(m), p(o), C.c_int64(sample_rows), C.c_int64(COLS))
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            threads = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
            return {"value": nb / best / 1e9, "unit": "GB/s", "cores": threads, "kind": "reference", "seconds": best,
                    "sample": f"cfg5 pipeline on the first {sample_rows} of {ROWS} rows ({best:.2f} s): real xtensor 0.27.1 headers "
                              "(oracle/_ref/libxtref_fast.so: -O3 -march=x86-64-v3 -fopenmp -DXTENSOR_USE_OPENMP, xtl stand-in, no xsimd/TBB "
                              "in this image); xtensor runs the reducers single-threaded (lazy xreducer stepper) and the map through "
                              f"its OpenMP strided-loop assigner ({threads} threads available)"}
    except Exception:
        pass
    from oracle import oracle
    xt = oracle.install()
    A = xt.HostArray.from_numpy(a)
    t0 = time.perf_counter()
    m = xt.evaluate(xt.mean(A, [0], dtype=xt.F32))
    xt.evaluate(xt.variance(A, [0], dtype=xt.F32))
    xt.evaluate(xt.exp(A - m))
    dt = time.perf_counter() - t0
    return {"value": nb / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": "port", "seconds": dt,
            "sample": f"cfg5 pipeline on the first {sample_rows} of {ROWS} rows ({dt:.2f} s): oracle/xtb_oracle.cpp (scalar restatement)"}


def cpu_baseline_cfg5(sample_rows):
    r = cpu_cfg5_sample(sample_rows)
    r["value"] = round(r["value"], 4)
    r.pop("seconds", None)
    return r


def run_reference(args):
    """xtensor's own CPU evaluation of the cfg5 pipeline on the box's host cores (a bounded sample per step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_rows = args.ref_sample_rows
    vals, info = [], None
    for i in range(args.warmup + args.steps):
        info = cpu_cfg5_sample(sample_rows)
        if i >= args.warmup:
            vals.append(info["value"])
    v = float(np.mean(vals))
    ms = job_bytes(sample_rows, COLS) / (v * 1e9) * 1e3
    info["value"] = round(v, 4)
    info.pop("seconds", None)
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": CONFIG, "cpu_baseline": info,
            "e2e": {"value": round(v, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---- the other BASELINE configs (side measurements, N = 1) ---------------------------------------------
def other_configs(lib, xt, capi, args):
    out = {}
    peak, _ = peak_hbm()

    def timed(fn, nbytes, iters=10):
        for _ in range(3):
            fn()
        ms = device_time_ms(lib, fn, iters) / iters
        return {"ms": round(ms, 4), "GBs": round(nbytes / ms / 1e6, 1), "frac_of_measured_peak": round(nbytes / ms / 1e6 / peak, 4),
                "kernel": lib.xtb_last_kernel().decode()}

    try:
        rng = np.random.default_rng(1)
        # cfg1: fp64 1-D 2^24 a + b
        n = 1 << 24
        a, b = (xt.DeviceArray.from_numpy(rng.uniform(-1, 1, n)) for _ in range(2))
        c = xt.DeviceArray.empty((n,), xt.F64)
        out["cfg1_add_f64"] = timed(lambda: xt.assign(c, a + b), 3 * n * 8)
        del a, b, c
        # cfg2: fp32 c(1024,1024,64) = sin(a) * b(1,1024,1) + 2.0f * d
        shp = (1024, 1024, 64)
        a2 = xt.DeviceArray.from_numpy(rng.uniform(-np.pi, np.pi, shp).astype(np.float32))
        b2 = xt.DeviceArray.from_numpy(rng.uniform(0.5, 1.5, (1, 1024, 1)).astype(np.float32))
        d2 = xt.DeviceArray.from_numpy(rng.uniform(-np.pi, np.pi, shp).astype(np.float32))
        c2 = xt.DeviceArray.empty(shp, xt.F32)
        e2 = xt.sin(a2) * b2 + np.float32(2.0) * d2
        out["cfg2_fused_broadcast_f32"] = timed(lambda: xt.assign(c2, e2), 3 * int(np.prod(shp)) * 4 + 4096, iters=20)
        del a2, b2, d2, c2, e2
        # cfg3: fp32 (4096,4096,16) sum / amax over axis 0 and axis 2
        x = xt.DeviceArray.from_numpy(rng.uniform(-1, 1, (4096, 4096, 16)).astype(np.float32))
        nb = 4096 * 4096 * 16 * 4
        o0, o2 = xt.DeviceArray.empty((4096, 16), xt.F32), xt.DeviceArray.empty((4096, 4096), xt.F32)
        red = lambda r, o: xt._run_reducer(r, xt.DeviceArray, out=o)           # into an existing container, like cfg1 / 2 / 4
        out["cfg3_sum_axis0"] = timed(lambda: red(xt.sum(x, [0]), o0), nb + 4096 * 16 * 4)
        out["cfg3_amax_axis0"] = timed(lambda: red(xt.amax(x, [0]), o0), nb + 4096 * 16 * 4)
        out["cfg3_sum_axis2"] = timed(lambda: red(xt.sum(x, [2]), o2), nb + 4096 * 4096 * 4)
        out["cfg3_amax_axis2"] = timed(lambda: red(xt.amax(x, [2]), o2), nb + 4096 * 4096 * 4)
        out["cfg3_mean_axis0_fused_finalize"] = timed(lambda: xt.assign(o0, xt.mean(x, [0], dtype=xt.F32)), nb + 4096 * 16 * 4)
        # index-carrying reduction (xt::argmax over the strided axis): ONE pass over packed (order key, index) keys
        oi = xt.DeviceArray.empty((4096, 16), xt.U64)
        def argmax0():
            iop, oop = x.operand(), oi.operand()
            capi.check(lib.xtb_argreduce(capi.RED_MAX, C.byref(iop), 0, C.byref(oop)))
        out["cfg3_argmax_axis0"] = timed(argmax0, nb + 4096 * 16 * 8)
        # beyond the named axes: reductions the planner rewrites into two passes (reduce_decomposed)
        o01, o02 = xt.DeviceArray.empty((16,), xt.F32), xt.DeviceArray.empty((4096,), xt.F32)
        out["cfg3_shape_sum_axes01_narrow_two_pass"] = timed(lambda: red(xt.sum(x, [0, 1]), o01), nb + 16 * 4)
        out["cfg3_shape_sum_axes02_mixed_two_pass"] = timed(lambda: red(xt.sum(x, [0, 2]), o02), nb + 4096 * 4)
        del o0, o2, oi, o01, o02
        del x
        # odd extents: the row pitch (32760 bytes) rules out 128-bit vectors and tensor maps
        xo = xt.DeviceArray.from_numpy(rng.uniform(-1, 1, (8191, 8190)).astype(np.float32))
        oo0, oo1, yo = xt.DeviceArray.empty((8190,), xt.F32), xt.DeviceArray.empty((8191,), xt.F32), xt.DeviceArray.empty((8191, 8190), xt.F32)
        nbo = 8191 * 8190 * 4
        out["odd_8191x8190_sum_axis0"] = timed(lambda: red(xt.sum(xo, [0]), oo0), nbo)
        out["odd_8191x8190_sum_axis1"] = timed(lambda: red(xt.sum(xo, [1]), oo1), nbo)
        out["odd_8191x8190_cumsum_axis0"] = timed(lambda: xt.cumsum(xo, 0, out=yo), 2 * nbo)
        del xo, oo0, oo1, yo
        # cfg4: fp64 (8192,8192) transpose(a) + view(b, range(0,_,2), all())
        a = xt.DeviceArray.from_numpy(rng.uniform(-1, 1, (8192, 8192)))
        b = xt.DeviceArray.from_numpy(rng.uniform(-1, 1, (16384, 8192)))
        o = xt.DeviceArray.empty((8192, 8192), xt.F64)
        e = xt.transpose(a) + xt.view(b, slice(0, None, 2), slice(None))
        out["cfg4_transpose_view_f64"] = timed(lambda: xt.assign(o, e), 3 * 8192 * 8192 * 8, iters=5)
        del a, b, o, e
        # cumsum (north_star: xaccumulator as a decoupled look-back scan); 2 x element size per element
        x = xt.DeviceArray.from_numpy(np.random.default_rng(2).uniform(-1, 1, 1 << 26).astype(np.float32))
        y = xt.DeviceArray.empty((1 << 26,), xt.F32)
        out["cumsum_flat_f32_2^26"] = timed(lambda: xt.cumsum(x, out=y), 2 * (1 << 26) * 4)
        x2, y2 = x.reshape_view((8192, 8192)), y.reshape_view((8192, 8192))
        out["cumsum_axis1_f32_8192x8192"] = timed(lambda: xt.cumsum(x2, 1, out=y2), 2 * (1 << 26) * 4)
        out["cumsum_axis0_f32_8192x8192"] = timed(lambda: xt.cumsum(x2, 0, out=y2), 2 * (1 << 26) * 4)
        del x, x2, y, y2
        # an expression with no ahead-of-time instantiation: run-time specialised kernel
        n = 1 << 26
        a3 = [xt.DeviceArray.from_numpy(np.random.default_rng(3 + i).uniform(0.5, 2, n).astype(np.float32)) for i in range(3)]
        o3 = xt.DeviceArray.empty((n,), xt.F32)
        e3 = xt.sqrt(a3[0] * a3[0] + a3[1] * a3[1]) / (a3[2] + np.float32(1.0))
        xt.assign(o3, e3)
        out["jit_hypot_div_f32_2^26"] = timed(lambda: xt.assign(o3, e3), 16 * n)
        del a3, o3, e3
        # context only: xtensor's own CPU evaluation of the other configs on bounded samples (host cores)
        try:
            from oracle import refbin
            out["cpu_reference_context"] = refbin.run_context()
        except Exception as ex:
            out["cpu_reference_context"] = {"error": repr(ex)}
    except Exception as ex:  # the headline number must survive a failure of the side measurements
        out["other_configs_error"] = repr(ex)
    return {"other_configs": out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--quick", action="store_true", help="development: one e2e step, no side measurements")
    ap.add_argument("--rows", type=int, default=0, help="development: total rows of the cfg5 matrix (default 262144)")
    ap.add_argument("--no-extra", action="store_true", help="skip the side measurements of cfg1/2/3/4 and the scans")
    ap.add_argument("--cpu-sample-rows", type=int, default=16384, help="rows of the cpu_baseline sample (N = 1)")
    ap.add_argument("--ref-sample-rows", type=int, default=4096, help="rows per step of --impl reference")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
