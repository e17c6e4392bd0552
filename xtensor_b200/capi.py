"""ctypes binding of the C ABI declared in include/xtb200.h.

This is plumbing for tests and bench.py: the product is libxtb200.so (CUDA, sm_100a)
and the header-only C++ boundary in include/xtb200/.  There is deliberately no
fallback: if the shared library is missing, or there is no GPU, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

MAX_DIM, MAX_LEAVES, MAX_INSNS, MAX_IMMS = 8, 8, 48, 16

# xtb_dtype
BOOL, I8, U8, I16, U16, I32, U32, I64, U64, F32, F64 = range(11)
# xtb_reduce_op
RED_SUM, RED_PROD, RED_MAX, RED_MIN, RED_NANMIN, RED_NANMAX = range(6)
# xtb_src
SRC_STACK, SRC_LEAF, SRC_IMM, SRC_REV = 0, 1, 2, 4
H2D, D2H, D2D = 1, 2, 3

OPCODES = dict(
    PUSH=0, CAST=1, NEG=2, NOT=3, BITNOT=4, ABS=5, EXP=6, EXP2=7, EXPM1=8, LOG=9, LOG10=10, LOG2=11,
    LOG1P=12, SQRT=13, CBRT=14, SIN=15, COS=16, TAN=17, ASIN=18, ACOS=19, ATAN=20, SINH=21, COSH=22,
    TANH=23, ASINH=24, ACOSH=25, ATANH=26, ERF=27, ERFC=28, TGAMMA=29, LGAMMA=30, CEIL=31, FLOOR=32,
    TRUNC=33, ROUND=34, NEARBYINT=35, RINT=36, ISFINITE=37, ISINF=38, ISNAN=39, SIGN=40, DEG2RAD=41,
    RAD2DEG=42, SQUARE=43, CUBE=44, ORDKEY=45,
    ADD=64, SUB=65, MUL=66, DIV=67, MOD=68, LOR=69, LAND=70, BOR=71, BAND=72, BXOR=73, SHL=74, SHR=75,
    LT=76, LE=77, GT=78, GE=79, EQ=80, NE=81, FMOD=82, REMAINDER=83, FMAX=84, FMIN=85, FDIM=86, POW=87,
    HYPOT=88, ATAN2=89, MAXIMUM=90, MINIMUM=91, NANMIN=92, NANMAX=93,
    WHERE=112, FMA=113, CLAMP=114,
)


class Operand(C.Structure):
    _fields_ = [("base", C.c_void_p), ("offset", C.c_int64), ("dtype", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int64 * MAX_DIM), ("stride", C.c_int64 * MAX_DIM)]


class Insn(C.Structure):
    _fields_ = [("op", C.c_uint8), ("type", C.c_uint8), ("src", C.c_uint8), ("arg", C.c_uint8)]


class Program(C.Structure):
    _fields_ = [("n_insns", C.c_int32), ("n_leaves", C.c_int32), ("n_imms", C.c_int32), ("reserved", C.c_int32),
                ("insns", Insn * MAX_INSNS), ("imms", C.c_uint64 * MAX_IMMS)]


class Finalize(C.Structure):
    """xtb_finalize: the step fused into the last store of a reduction (include/xtb200.h)."""
    _fields_ = [("op", C.c_int32), ("type", C.c_int32), ("imm", C.c_uint64)]


FIN_NONE, FIN_DIV, FIN_DIV_SQRT = 0, 1, 2


class XtbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"xtb status {code}: {msg}")
        self.code = code


ERR_INVALID, ERR_SHAPE, ERR_AXIS, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE, ERR_NCCL, ERR_OOM = range(-1, -9, -1)

# every symbol include/xtb200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "xtb_abi_version", "xtb_init", "xtb_device_count", "xtb_sync", "xtb_last_error", "xtb_set_stream",
    "xtb_get_stream", "xtb_malloc", "xtb_free", "xtb_memcpy", "xtb_memset", "xtb_host_alloc", "xtb_host_free",
    "xtb_event_create", "xtb_event_record", "xtb_event_elapsed_ms", "xtb_event_destroy",
    "xtb_graph_begin", "xtb_graph_end", "xtb_graph_launch", "xtb_graph_destroy", "xtb_fork_begin", "xtb_fork_end", "xtb_fork_join",
    "xtb_assign", "xtb_assign_host", "xtb_reduce", "xtb_scan", "xtb_comm_unique_id", "xtb_comm_init", "xtb_comm_destroy",
    "xtb_comm_info", "xtb_comm_p2p_handle", "xtb_comm_p2p_attach", "xtb_allreduce", "xtb_launch_count", "xtb_last_kernel", "xtb_program_result_type",
    "xtb_reduce_fin", "xtb_set_option", "xtb_get_option", "xtb_graph_kernel_count", "xtb_argreduce",
]

_LIB = None


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libxtb200.so")


def lib():
    """Load libxtb200.so (built in-tree by __graft_entry__.build()).  No fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the CUDA extension is the only execution path; there is no CPU fallback)")
    L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    vp, i32, i64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
    sig = {
        "xtb_abi_version": (i32, []),
        "xtb_init": (i32, [i32]),
        "xtb_device_count": (i32, [C.POINTER(i32)]),
        "xtb_sync": (i32, []),
        "xtb_fork_begin": (i32, []),
        "xtb_fork_end": (i32, []),
        "xtb_fork_join": (i32, []),
        "xtb_last_error": (C.c_char_p, []),
        "xtb_set_stream": (i32, [vp]),
        "xtb_get_stream": (vp, []),
        "xtb_malloc": (i32, [sz, C.POINTER(vp)]),
        "xtb_free": (i32, [vp]),
        "xtb_memcpy": (i32, [vp, vp, sz, i32]),
        "xtb_memset": (i32, [vp, i32, sz]),
        "xtb_host_alloc": (i32, [sz, C.POINTER(vp)]),
        "xtb_host_free": (i32, [vp]),
        "xtb_event_create": (i32, [C.POINTER(vp)]),
        "xtb_event_record": (i32, [vp]),
        "xtb_event_elapsed_ms": (i32, [vp, vp, C.POINTER(C.c_float)]),
        "xtb_event_destroy": (i32, [vp]),
        "xtb_graph_begin": (i32, []),
        "xtb_graph_end": (i32, [C.POINTER(vp)]),
        "xtb_graph_launch": (i32, [vp]),
        "xtb_graph_destroy": (i32, [vp]),
        "xtb_assign": (i32, [C.POINTER(Program), C.POINTER(Operand), C.POINTER(Operand)]),
        "xtb_assign_host": (i32, [C.POINTER(Program), C.POINTER(Operand), C.POINTER(Operand), i64]),
        "xtb_reduce": (i32, [i32, i32, C.POINTER(Program), C.POINTER(Operand), i32, C.POINTER(i64), i32,
                             C.POINTER(C.c_int32), i32, vp, C.POINTER(Operand), i32]),
        "xtb_reduce_fin": (i32, [i32, i32, C.POINTER(Program), C.POINTER(Operand), i32, C.POINTER(i64), i32,
                                 C.POINTER(C.c_int32), i32, vp, C.POINTER(Operand), i32, C.POINTER(Finalize)]),
        "xtb_set_option": (i32, [C.c_char_p, C.c_longlong]),
        "xtb_get_option": (C.c_longlong, [C.c_char_p]),
        "xtb_graph_kernel_count": (i32, [vp]),
        "xtb_scan": (i32, [i32, i32, C.POINTER(Operand), i32, C.POINTER(Operand)]),
        "xtb_argreduce": (i32, [i32, C.POINTER(Operand), i32, C.POINTER(Operand)]),
        "xtb_comm_unique_id": (i32, [vp]),
        "xtb_comm_init": (i32, [i32, i32, vp]),
        "xtb_comm_destroy": (i32, []),
        "xtb_comm_info": (i32, [C.POINTER(i32), C.POINTER(i32)]),
        "xtb_comm_p2p_handle": (i32, [vp]),
        "xtb_comm_p2p_attach": (i32, [vp, i32]),
        "xtb_allreduce": (i32, [vp, sz, i32, i32]),
        "xtb_launch_count": (i64, [i32]),
        "xtb_last_kernel": (C.c_char_p, []),
        "xtb_program_result_type": (i32, [C.POINTER(Program), C.POINTER(C.c_int32)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _LIB = L
    return L


def check(code: int):
    if code != 0:
        raise XtbError(code, lib().xtb_last_error().decode())
