"""GPU: the C++ drop-in boundary (include/xtb200/xtensor_b200.hpp) against the REAL xtensor.
tests/cpp/test_dropin.cpp evaluates the same expressions with host containers (xtensor's own CPU
loops) and with device containers (libxtb200) and compares them; it is compiled in the build
container (tests/cpp/Makefile, needs /root/reference) and only *run* here."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name", ["test_dropin", "test_dropin_reducers"])
def test_cpp_dropin(gpu, name):
    exe = os.path.join(HERE, "cpp", "_build", name)
    assert os.path.exists(exe), f"{exe} missing: run `make -C tests/cpp` where /root/reference exists"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "OK" in r.stdout


def _run(cmd, env=None, timeout=600):
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    print(r.stdout[-3000:], r.stderr[-3000:])
    return r


def test_cpp_dist_single_rank(gpu):
    """tests/cpp/test_dist.cpp at world size 1: xtb::dist degenerates to the single-GPU path."""
    exe = os.path.join(HERE, "cpp", "_build", "test_dist")
    assert os.path.exists(exe), f"{exe} missing: run `make -C tests/cpp` where /root/reference exists"
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = _run([exe], env=env)
    assert r.returncode == 0 and "OK test_dist" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_cpp_dist_two_ranks(gpu):
    """The same program as two processes on two GPUs (own rendezvous + NCCL / peer-memory bring-up from C++):
    sharded results equal the single-process ones.  Needs two devices."""
    import ctypes as C
    n = C.c_int(0)
    gpu.xtb_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip("one GPU on this box: the two-rank C++ test runs under `gpurun --gpus 2` (tools/_run_cpp_checks.sh)")
    exe = os.path.join(HERE, "cpp", "_build", "test_dist")
    import sys
    r = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
              "--master-port", "29571", "--no-python", exe])
    assert r.returncode == 0 and r.stdout.count("OK test_dist") == 2, r.stdout[-3000:] + r.stderr[-3000:]
