"""ctypes loader of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, by __graft_entry__.smoke()
and by bench.py's cpu_baseline / --impl reference legs -- never by xtensor_b200/.
"""
import ctypes as C
import os
import subprocess

from xtensor_b200.capi import Operand, Program

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    path = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "xtb_oracle.cpp")
    if force or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return path


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        i32, i64, vp = C.c_int, C.c_int64, C.c_void_p
        L.xto_last_error.restype = C.c_char_p
        L.xto_assign.restype = i32
        L.xto_assign.argtypes = [C.POINTER(Program), C.POINTER(Operand), C.POINTER(Operand)]
        L.xto_reduce.restype = i32
        L.xto_reduce.argtypes = [i32, i32, C.POINTER(Program), C.POINTER(Operand), i32, C.POINTER(i64), i32,
                                 C.POINTER(C.c_int32), i32, vp, C.POINTER(Operand), i32]
        L.xto_scan.restype = i32
        L.xto_scan.argtypes = [i32, i32, C.POINTER(Operand), i32, C.POINTER(Operand)]
        L.xto_argreduce.restype = i32
        L.xto_argreduce.argtypes = [i32, C.POINTER(Operand), i32, C.POINTER(Operand)]
        _LIB = L
    return _LIB


def install():
    """Register the oracle as the evaluator of HostArray expressions (tests only)."""
    from xtensor_b200 import expr
    expr.set_host_backend(lib())
    return expr
