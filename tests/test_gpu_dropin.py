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
