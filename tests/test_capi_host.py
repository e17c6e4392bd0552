"""CPU: the C-ABI library loads, exports every symbol include/xtb200.h declares, validates
programs on the host, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from xtensor_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "xtb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xtb_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = capi.lib()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.xtb_abi_version() == 1


def test_struct_layouts_match_the_header():
    assert C.sizeof(capi.Operand) == 8 + 8 + 4 + 4 + 8 * 8 + 8 * 8
    assert C.sizeof(capi.Insn) == 4
    assert C.sizeof(capi.Program) == 16 + 4 * capi.MAX_INSNS + 8 * capi.MAX_IMMS


def test_opcode_table_matches_the_header():
    src = open(os.path.join(ROOT, "include", "xtb200.h")).read()
    for name, val in capi.OPCODES.items():
        m = re.search(rf"XTB_OP_{name}\s*=\s*(\d+)", src)
        assert m and int(m.group(1)) == val, name


def test_program_validation_on_host(xt):
    lib = capi.lib()
    a = xt.HostArray.from_numpy(np.ones((2, 3), np.float32))
    b = xt.HostArray.from_numpy(np.ones((3,), np.uint8))
    lw = xt.lower(xt.sin(a) * b + 2.0)
    prog = lw.program()
    dts = (C.c_int32 * 2)(xt.F32, xt.U8)
    assert lib.xtb_program_result_type(C.byref(prog), dts) == xt.F64
    # wrong leaf dtype, stack underflow, unknown opcode are rejected with a message
    bad = (C.c_int32 * 2)(xt.F64, xt.U8)
    assert lib.xtb_program_result_type(C.byref(prog), bad) == capi.ERR_INVALID
    assert b"dtype" in lib.xtb_last_error()
    prog.insns[0].op = 200
    assert lib.xtb_program_result_type(C.byref(prog), dts) == capi.ERR_INVALID
    empty = capi.Program()
    assert lib.xtb_program_result_type(C.byref(empty), None) == capi.ERR_INVALID


def test_lowering_is_canonical(xt):
    """The Python mirror must emit exactly the encoding the compile-time programs match."""
    a, b, d = (xt.HostArray.from_numpy(np.ones(s, np.float32)) for s in ((2, 3, 4), (1, 3, 1), (2, 3, 4)))
    lw = xt.lower(xt.sin(a) * b + np.float32(2.0) * d)
    O = capi.OPCODES
    assert lw.insns == [(O["PUSH"], xt.F32, capi.SRC_LEAF, 0), (O["SIN"], xt.F32, 0, 0), (O["MUL"], xt.F32, capi.SRC_LEAF, 1),
                        (O["PUSH"], xt.F32, capi.SRC_IMM, 0), (O["MUL"], xt.F32, capi.SRC_LEAF, 2), (O["ADD"], xt.F32, 0, 0)]
    lw = xt.lower(xt.exp(a - d))
    assert lw.insns == [(O["PUSH"], xt.F32, capi.SRC_LEAF, 0), (O["SUB"], xt.F32, capi.SRC_LEAF, 1), (O["EXP"], xt.F32, 0, 0)]
    lw = xt.lower(a + a)                       # the same container twice is one leaf
    assert len(lw.leaves) == 1
    lw = xt.lower(2.0 * a)                     # double scalar * float array computes in double
    assert lw.insns == [(O["PUSH"], xt.F32, capi.SRC_LEAF, 0), (O["CAST"], xt.F32, 0, xt.F64),
                        (O["MUL"], xt.F64, capi.SRC_IMM | capi.SRC_REV, 0)]


def test_views_are_descriptors(xt):
    a = xt.HostArray.from_numpy(np.arange(24, dtype=np.float64).reshape(2, 3, 4))
    t = xt.transpose(a)
    assert t.shape == (4, 3, 2) and t.strides == (1, 4, 12)
    v = xt.view(a, slice(None), 1, slice(0, None, 2))
    assert v.shape == (2, 2) and v.strides == (12, 2) and v.offset == 4
    assert np.array_equal(v.numpy(), np.arange(24).reshape(2, 3, 4)[:, 1, ::2])
    b = xt.broadcast(xt.HostArray.from_numpy(np.arange(4.0)), (3, 4))
    assert b.strides == (0, 1)
    with pytest.raises(xt.BroadcastError):
        xt.HostArray.from_numpy(np.ones(3)) + xt.HostArray.from_numpy(np.ones(4))
    with pytest.raises(RuntimeError, match="sorted"):
        xt.sum(a, [1, 0])
    with pytest.raises(RuntimeError, match="out of bounds"):
        xt.sum(a, [3])


def test_no_cpu_fallback_without_gpu():
    """On a box without a CUDA device every compute entry point must fail loudly."""
    lib = capi.lib()
    n = C.c_int()
    assert lib.xtb_device_count(C.byref(n)) == 0
    if n.value > 0:
        pytest.skip("a GPU is present")
    p = C.c_void_p()
    assert lib.xtb_malloc(16, C.byref(p)) == capi.ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.xtb_last_error()
    op, prog = capi.Operand(), capi.Program()
    prog.n_insns = 1
    prog.n_imms = 1
    prog.insns[0] = capi.Insn(0, capi.F32, capi.SRC_IMM, 0)
    op.dtype = capi.F32
    op.ndim = 1
    op.shape[0] = 4
    assert lib.xtb_assign(C.byref(prog), C.byref(op), None) == capi.ERR_NO_DEVICE
    # reductions too -- single pass, and a shape the planner would rewrite into two passes (narrow: 4 outputs)
    rp = capi.Program()
    rp.n_insns = 1
    rp.n_leaves = 1
    rp.insns[0] = capi.Insn(0, capi.F32, capi.SRC_LEAF, 0)
    for rows in (64, 1 << 20):
        leaf, out = capi.Operand(), capi.Operand()
        leaf.dtype = out.dtype = capi.F32
        leaf.ndim, out.ndim = 2, 1
        leaf.shape[0], leaf.shape[1], leaf.stride[0], leaf.stride[1] = rows, 4, 4, 1
        out.shape[0], out.stride[0] = 4, 1
        shape = (C.c_int64 * 2)(rows, 4)
        axes = (C.c_int32 * 1)(0)
        assert lib.xtb_reduce(capi.RED_SUM, capi.F32, C.byref(rp), C.byref(leaf), 2, shape, 1, axes, 0, None, C.byref(out), 0) == capi.ERR_NO_DEVICE


def test_process_options_round_trip():
    """Every switch the header documents is known to the library (no device needed), unknown names are rejected."""
    lib = capi.lib()
    doc = open(os.path.join(ROOT, "include", "xtb200.h")).read()
    names = ["no_static", "no_jit", "no_staged", "no_tma", "jit_min_elems", "jit_verbose", "scan_variant", "scan_nv",
             "tile_variant", "arg_two_pass", "no_pdl", "no_decompose", "reduce_split", "reduce_g"]
    for n in names:
        assert f'"{n}"' in doc, f"{n} is not documented in include/xtb200.h"
        old = lib.xtb_get_option(n.encode())
        assert old >= 0, n
        assert lib.xtb_set_option(n.encode(), old + 1) == 0
        assert lib.xtb_get_option(n.encode()) == old + 1
        assert lib.xtb_set_option(n.encode(), old) == 0
    assert lib.xtb_get_option(b"no_such_option") == -1
    assert lib.xtb_set_option(b"no_such_option", 1) != 0
