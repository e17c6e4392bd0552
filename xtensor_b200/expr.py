"""Python mirror of the part of xtensor's expression API that sits on the hot path.

Test / benchmark plumbing only -- the reference-facing host side is the C++ header
include/xtb200/xtensor_b200.hpp.  This module exists so that the parity tests read
like the reference's own tests (`xt.noalias(c).assign(xt.sin(a) * b + 2.0 * d)`,
`xt.sum(a, [0])`) while calling libxtb200 through its C ABI.

It mirrors, with the same names and argument meaning:
  containers / views  xtensor/xarray, xt::transpose (misc/xmanipulation.hpp:238-259),
                      xt::view with int / range / all / newaxis slices (views/xview.hpp:1824),
                      xt::broadcast (views/xbroadcast.hpp:158-237)
  lazy nodes          xfunction via operators and xt::sin... (core/xoperation.hpp:231-330,
                      core/xmath.hpp:443-1670); C++ usual arithmetic conversions decide every
                      node's value type (core/xfunction.hpp:132-142)
  evaluation          xt::noalias(out) = expr (core/xnoalias.hpp:160-230), resize to the
                      broadcast shape (core/xassign.hpp:581-607), broadcast errors
  reducers            xt::sum / prod / amax / amin / mean / variance (core/xmath.hpp:367-421,
                      777-800, 1803-1912, 2074-2105) with keep_dims / initial options
  accumulators        xt::cumsum / cumprod (core/xmath.hpp:2247-2297)

Arrays are either DeviceArray (HBM, evaluated by libxtb200) or HostArray (numpy,
evaluated by whatever host backend the *tests* register -- the CPU oracle).  The
package itself never imports the oracle.
"""
from __future__ import annotations

import builtins
import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import capi
from .capi import (BOOL, F32, F64, I8, I16, I32, I64, U8, U16, U32, U64, OPCODES, SRC_IMM, SRC_LEAF,
                   SRC_REV, SRC_STACK)

NP_OF = {BOOL: np.bool_, I8: np.int8, U8: np.uint8, I16: np.int16, U16: np.uint16, I32: np.int32,
         U32: np.uint32, I64: np.int64, U64: np.uint64, F32: np.float32, F64: np.float64}
DT_OF = {np.dtype(v): k for k, v in NP_OF.items()}
SIZE_OF = {k: np.dtype(v).itemsize for k, v in NP_OF.items()}


class BroadcastError(ValueError):
    """xt::broadcast_error (core/xstrides.hpp:748-777)."""


# ---- C++ type rules -----------------------------------------------------------------
def regtype(dt: int) -> int:
    """Integral promotion: bool/int8/uint8/int16/uint16 -> int."""
    return I32 if dt < I32 else dt


def common_type(a: int, b: int) -> int:
    """Usual arithmetic conversions on two (already promoted) types."""
    a, b = regtype(a), regtype(b)
    if F64 in (a, b):
        return F64
    if F32 in (a, b):
        return F32
    if a == b:
        return a
    rank = {I32: 1, U32: 1, I64: 2, U64: 2}
    signed = {I32: True, U32: False, I64: True, U64: False}
    if signed[a] == signed[b]:
        return a if rank[a] >= rank[b] else b
    u, s = (a, b) if not signed[a] else (b, a)
    if rank[u] >= rank[s]:
        return u
    return s  # int64 represents every uint32


def dtype_of_scalar(x) -> int:
    if isinstance(x, (bool, np.bool_)):
        return BOOL
    if isinstance(x, np.generic):
        return DT_OF[np.dtype(type(x))]
    if isinstance(x, int):
        return I32
    if isinstance(x, float):
        return F64
    raise TypeError(f"unsupported scalar {x!r}")


def compute_strides(shape: Sequence[int]) -> Tuple[int, ...]:
    """Row-major strides in elements, 0 for extent-1 dims (core/xstrides.hpp:503-580)."""
    st, acc = [], 1
    for ext in reversed(shape):
        st.append(0 if ext == 1 else acc)
        acc *= ext
    return tuple(reversed(st))


def broadcast_shape(shapes: Sequence[Sequence[int]]) -> Tuple[int, ...]:
    """xt::broadcast_shape (core/xstrides.hpp:737-780): right aligned, 1 stretches."""
    nd = max((len(s) for s in shapes), default=0)
    out = [1] * nd
    for s in shapes:
        off = nd - len(s)
        for i, ext in enumerate(s):
            cur = out[off + i]
            if cur == 1:
                out[off + i] = ext
            elif ext != 1 and ext != cur:
                raise BroadcastError(f"Incompatible dimension of arrays: {tuple(shapes)}")
    return tuple(out)


# ---- expressions -----------------------------------------------------------------------
class Expr:
    """CRTP-base stand-in: anything with a shape and a value dtype that can be lowered."""
    dtype: int
    shape: Tuple[int, ...]
    __array_ufunc__ = None      # numpy scalars must defer to our reflected operators

    def _bin(self, other, op, rev=False):
        other = as_expr(other)
        return Func(op, (other, self) if rev else (self, other))

    def __add__(self, o): return self._bin(o, "ADD")
    def __radd__(self, o): return self._bin(o, "ADD", True)
    def __sub__(self, o): return self._bin(o, "SUB")
    def __rsub__(self, o): return self._bin(o, "SUB", True)
    def __mul__(self, o): return self._bin(o, "MUL")
    def __rmul__(self, o): return self._bin(o, "MUL", True)
    def __truediv__(self, o): return self._bin(o, "DIV")
    def __rtruediv__(self, o): return self._bin(o, "DIV", True)
    def __mod__(self, o): return self._bin(o, "MOD")
    def __and__(self, o): return self._bin(o, "BAND")
    def __or__(self, o): return self._bin(o, "BOR")
    def __xor__(self, o): return self._bin(o, "BXOR")
    def __lshift__(self, o): return self._bin(o, "SHL")
    def __rshift__(self, o): return self._bin(o, "SHR")
    def __lt__(self, o): return self._bin(o, "LT")
    def __le__(self, o): return self._bin(o, "LE")
    def __gt__(self, o): return self._bin(o, "GT")
    def __ge__(self, o): return self._bin(o, "GE")
    def eq(self, o): return self._bin(o, "EQ")          # xt::equal
    def ne(self, o): return self._bin(o, "NE")          # xt::not_equal
    def __neg__(self): return Func("NEG", (self,))
    def __pos__(self): return self
    def __invert__(self): return Func("BITNOT", (self,))
    __hash__ = object.__hash__


class Scalar(Expr):
    """xscalar<T> (containers/xscalar.hpp:84-300): 0-d, broadcasts trivially."""

    def __init__(self, value, dtype: Optional[int] = None):
        self.dtype = dtype_of_scalar(value) if dtype is None else dtype
        self.value = NP_OF[self.dtype](value)
        self.shape = ()


def as_expr(x) -> Expr:
    return x if isinstance(x, Expr) else Scalar(x)


FLOAT_UNARY = {"EXP", "EXP2", "EXPM1", "LOG", "LOG10", "LOG2", "LOG1P", "SQRT", "CBRT", "SIN", "COS", "TAN", "ASIN",
               "ACOS", "ATAN", "SINH", "COSH", "TANH", "ASINH", "ACOSH", "ATANH", "ERF", "ERFC", "TGAMMA", "LGAMMA",
               "CEIL", "FLOOR", "TRUNC", "ROUND", "NEARBYINT", "RINT"}
PRED_UNARY = {"NOT", "ISFINITE", "ISINF", "ISNAN"}
CMP_BINARY = {"LT", "LE", "GT", "GE", "EQ", "NE", "LOR", "LAND"}
FLOAT_BINARY = {"FMOD", "REMAINDER", "FMAX", "FMIN", "FDIM", "POW", "HYPOT", "ATAN2"}
INT_BINARY = {"MOD", "BOR", "BAND", "BXOR", "SHL", "SHR"}


def _float_of(dt: int) -> int:
    """std::sin(int) etc. return double; float stays float."""
    return dt if dt in (F32, F64) else F64


class Func(Expr):
    """xfunction<F, CT...>: value type = decltype(f(value_types...))."""

    def __init__(self, op: str, args: Tuple[Expr, ...], cast_to: Optional[int] = None):
        self.op, self.args, self.cast_to = op, tuple(args), cast_to
        self.shape = broadcast_shape([a.shape for a in self.args])
        ats = [a.dtype for a in self.args]
        if op == "CAST":
            self.operand_type, self.dtype = regtype(ats[0]), cast_to
        elif op in FLOAT_UNARY or op in ("DEG2RAD", "RAD2DEG"):
            self.operand_type = self.dtype = _float_of(regtype(ats[0]))
        elif op in PRED_UNARY:
            self.operand_type, self.dtype = regtype(ats[0]), BOOL
        elif len(ats) == 1:  # NEG BITNOT ABS SIGN SQUARE CUBE: result of promoted operand
            self.operand_type = self.dtype = regtype(ats[0])
        elif op in CMP_BINARY:
            self.operand_type, self.dtype = common_type(ats[0], ats[1]), BOOL
        elif op in FLOAT_BINARY:
            self.operand_type = self.dtype = _float_of(common_type(ats[0], ats[1]))
        elif op in ("SHL", "SHR"):
            self.operand_type = self.dtype = regtype(ats[0])
        elif op == "WHERE":
            self.operand_type = self.dtype = common_type(ats[1], ats[2])
            if ats[1] == ats[2]:
                self.dtype = ats[1]       # c ? a : b keeps a narrow type when both agree
        elif op in ("FMA", "CLAMP"):
            t = common_type(common_type(ats[0], ats[1]), ats[2])
            self.operand_type = self.dtype = _float_of(t) if op == "FMA" else t
            if op == "CLAMP" and ats[0] == ats[1] == ats[2]:
                self.dtype = ats[0]
        elif op in ("MAXIMUM", "MINIMUM"):
            self.operand_type = self.dtype = common_type(ats[0], ats[1])
            if ats[0] == ats[1]:
                self.dtype = ats[0]       # select(c, t1, t2) with identical types
        else:  # arithmetic / bitwise
            self.operand_type = self.dtype = common_type(ats[0], ats[1])
        if op in INT_BINARY and self.operand_type in (F32, F64):
            raise TypeError(f"{op} needs integer operands")


class Reducer(Expr):
    """Lazy xreducer node (reducers/xreducer.hpp:803-909); evaluated eagerly when used."""

    def __init__(self, op: int, e: Expr, axes, keep_dims=False, initial=None, acc_dtype: Optional[int] = None):
        self.op, self.e, self.keep_dims, self.initial = op, as_expr(e), bool(keep_dims), initial
        nd = len(self.e.shape)
        if axes is None:
            axes = list(range(nd))
        elif isinstance(axes, (int, np.integer)):
            axes = [int(axes)]
        # normalize_axis (utils/xutils.hpp:349-460): negative axes wrap; order is kept, so an
        # unsorted list still raises "Reducing axes should be sorted."
        self.axes = [int(a) + nd if int(a) < 0 else int(a) for a in axes]
        for i in range(1, len(self.axes)):
            if self.axes[i] < self.axes[i - 1]:
                raise RuntimeError("Reducing axes should be sorted.")
            if self.axes[i] == self.axes[i - 1]:
                raise RuntimeError("Reducing axes should not contain duplicates.")
        if self.axes and self.axes[-1] > nd - 1:
            raise RuntimeError(f"Axis {self.axes[-1]} out of bounds for reduction.")
        init_t = self.e.dtype if acc_dtype is None else acc_dtype
        # result_type = decltype(reduce(init, x)) (reducers/xreducer.hpp:292-298)
        if op in (capi.RED_SUM, capi.RED_PROD):
            self.acc = common_type(init_t, self.e.dtype)
            self.dtype = self.acc
        else:  # maximum/minimum select: both are value_type
            self.acc = regtype(self.e.dtype)
            self.dtype = self.e.dtype
        sh = []
        for d, ext in enumerate(self.e.shape):
            if d in self.axes:
                if self.keep_dims:
                    sh.append(1)
            else:
                sh.append(ext)
        self.shape = tuple(sh)


# ---- arrays ---------------------------------------------------------------------------------
class Array(Expr):
    """Shape/strides/offset over a flat buffer: xstrided_container + strided views."""

    def __init__(self, shape, strides, offset, dtype, owner):
        self.shape = tuple(int(s) for s in shape)
        self.strides = tuple(int(s) for s in strides)
        self.offset = int(offset)
        self.dtype = dtype
        self.owner = owner      # the allocation (DeviceBuffer or numpy array)

    # -- to be provided by subclasses
    def base_ptr(self) -> int: raise NotImplementedError
    def _like(self, shape, strides, offset): raise NotImplementedError

    @property
    def size(self) -> int:
        return int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1

    @property
    def ndim(self) -> int:
        return len(self.shape)

    def operand(self) -> capi.Operand:
        op = capi.Operand()
        op.base = self.base_ptr()
        op.offset = self.offset
        op.dtype = self.dtype
        op.ndim = len(self.shape)
        for i, (s, st) in enumerate(zip(self.shape, self.strides)):
            op.shape[i] = s
            op.stride[i] = st
        return op

    # xt::transpose(e[, permutation]) -- permuted (shape, strides), misc/xmanipulation.hpp:120-259
    def transpose(self, perm: Optional[Sequence[int]] = None):
        nd = len(self.shape)
        perm = list(reversed(range(nd))) if perm is None else [int(p) for p in perm]
        if sorted(perm) != list(range(nd)):
            raise RuntimeError("Permutation does not have the same size as shape / contains duplicates")
        return self._like([self.shape[p] for p in perm], [self.strides[p] for p in perm], self.offset)

    # xt::view(e, slices...) restricted to strided slices (views/xview.hpp:169-176, 1268-1345)
    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        n_idx = builtins.sum(1 for k in key if k is not None and k is not Ellipsis)
        if Ellipsis in key:
            i = key.index(Ellipsis)
            key = key[:i] + (slice(None),) * (len(self.shape) - n_idx) + key[i + 1:]
        shape, strides, offset, d = [], [], self.offset, 0
        for k in key:
            if k is None:                       # xt::newaxis
                shape.append(1); strides.append(0)
                continue
            ext, st = self.shape[d], self.strides[d]
            if isinstance(k, (int, np.integer)):
                k = int(k) + ext if k < 0 else int(k)
                if not 0 <= k < ext:
                    raise IndexError("index out of range")
                offset += k * st
            elif isinstance(k, slice):          # xt::range(a, b, step) / xt::all()
                a, b, step = k.indices(ext)
                n = max(0, (b - a + (step - (1 if step > 0 else -1))) // step)
                offset += a * st
                shape.append(n); strides.append(0 if n == 1 else st * step)
            else:
                raise TypeError("only int / range / all / newaxis slices are strided views")
            d += 1
        for dd in range(d, len(self.shape)):
            shape.append(self.shape[dd]); strides.append(self.strides[dd])
        return self._like(shape, strides, offset)

    def broadcast(self, shape: Sequence[int]):
        """xt::broadcast(e, shape): stride-0 descriptor (views/xbroadcast.hpp:158-237)."""
        shape = tuple(int(s) for s in shape)
        broadcast_shape([self.shape, shape])
        off = len(shape) - len(self.shape)
        if off < 0:
            raise BroadcastError("broadcast to a lower rank")
        st = [0] * off + [0 if self.shape[i] == 1 else self.strides[i] for i in range(len(self.shape))]
        for i, ext in enumerate(self.shape):
            if ext != 1 and ext != shape[off + i]:
                raise BroadcastError(f"cannot broadcast {self.shape} to {shape}")
        return self._like(shape, st, self.offset)

    def reshape_view(self, shape: Sequence[int]):
        if self.strides != compute_strides(self.shape):
            raise RuntimeError("reshape_view needs a contiguous operand")
        return self._like(shape, compute_strides(shape), self.offset)


class HostArray(Array):
    """numpy-backed twin of DeviceArray: same descriptors, host pointers (oracle input)."""

    def __init__(self, shape, strides, offset, dtype, owner):
        super().__init__(shape, strides, offset, dtype, owner)

    @staticmethod
    def from_numpy(a: np.ndarray) -> "HostArray":
        a = np.asarray(a, order="C")
        return HostArray(a.shape, compute_strides(a.shape), 0, DT_OF[a.dtype], a.reshape(-1).copy())

    @staticmethod
    def empty(shape, dtype: int) -> "HostArray":
        n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
        return HostArray(shape, compute_strides(shape), 0, dtype, np.zeros(n, dtype=NP_OF[dtype]))

    def base_ptr(self) -> int:
        return self.owner.ctypes.data

    def _like(self, shape, strides, offset):
        return HostArray(shape, strides, offset, self.dtype, self.owner)

    def numpy(self) -> np.ndarray:
        isz = self.owner.itemsize
        v = np.lib.stride_tricks.as_strided(self.owner[self.offset:], shape=self.shape,
                                            strides=[s * isz for s in self.strides], writeable=False)
        return np.array(v)


class DeviceBuffer:
    """xtb::device_uvector: owns HBM; contents uninitialised (uvector contract, xstorage.hpp:217-228)."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        capi.check(capi.lib().xtb_malloc(max(self.nbytes, 1), C.byref(p)))
        self.ptr = p.value or 0

    def __del__(self):
        try:
            if getattr(self, "ptr", 0):
                capi.lib().xtb_free(C.c_void_p(self.ptr))
                self.ptr = 0
        except Exception:
            pass


class ForeignBuffer:
    """xtb::device_span: device memory owned by somebody else (cudaMalloc, torch, DLPack); never freed here.
    `keepalive` pins the owner object for as long as arrays over the buffer exist."""

    def __init__(self, ptr: int, nbytes: int, keepalive=None):
        self.ptr, self.nbytes, self.keepalive = int(ptr), int(nbytes), keepalive


class DeviceArray(Array):
    """Container with device storage (xtensor_container<device_uvector<T>, ...>)."""

    @staticmethod
    def from_pointer(ptr: int, shape, dtype: int, strides=None, keepalive=None) -> "DeviceArray":
        """xtb::adapt(device_ptr, shape[, strides]) (xt::adapt with no_ownership, containers/xadapt.hpp:105-215):
        an array over foreign device memory, element strides, no copy."""
        shape = tuple(int(s) for s in shape)
        strides = compute_strides(shape) if strides is None else tuple(int(s) for s in strides)
        span = 0 if 0 in shape else 1 + builtins.sum((e - 1) * builtins.abs(st) for e, st in zip(shape, strides))
        return DeviceArray(shape, strides, 0, dtype, ForeignBuffer(ptr, span * SIZE_OF[dtype], keepalive))

    @staticmethod
    def from_torch(t) -> "DeviceArray":
        """Zero-copy view of a CUDA torch tensor (data_ptr / shape / stride); the tensor is kept alive.  The
        caller orders the streams (torch.cuda.synchronize() or a shared stream via xtb_set_stream)."""
        import torch
        names = {torch.float32: F32, torch.float64: F64, torch.int32: I32, torch.int64: I64, torch.int16: I16,
                 torch.int8: I8, torch.uint8: U8, torch.bool: BOOL}
        if not t.is_cuda:
            raise ValueError("from_torch needs a CUDA tensor (there is no CPU path)")
        return DeviceArray.from_pointer(t.data_ptr(), tuple(t.shape), names[t.dtype], tuple(t.stride()), keepalive=t)

    @staticmethod
    def empty(shape, dtype: int) -> "DeviceArray":
        shape = tuple(int(s) for s in shape)
        n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
        return DeviceArray(shape, compute_strides(shape), 0, dtype, DeviceBuffer(n * SIZE_OF[dtype]))

    @staticmethod
    def from_numpy(a: np.ndarray) -> "DeviceArray":
        a = np.asarray(a, order="C")
        d = DeviceArray.empty(a.shape, DT_OF[a.dtype])
        if a.nbytes:
            capi.check(capi.lib().xtb_memcpy(C.c_void_p(d.owner.ptr), C.c_void_p(a.ctypes.data), a.nbytes, capi.H2D))
            capi.check(capi.lib().xtb_sync())
        return d

    @staticmethod
    def from_npy(path: str) -> "DeviceArray":
        """.npy fixture -> device (mirror of xtb::load_npy, include/xtb200/xtb_npy.hpp)."""
        return DeviceArray.from_numpy(np.load(path))

    def to_npy(self, path: str) -> None:
        """device -> .npy (mirror of xtb::dump_npy)."""
        np.save(path, self.numpy())

    def base_ptr(self) -> int:
        return self.owner.ptr

    def _like(self, shape, strides, offset):
        return DeviceArray(shape, strides, offset, self.dtype, self.owner)

    def numpy(self) -> np.ndarray:
        """Device -> host copy of the *viewed* elements (whole buffer is fetched)."""
        n = self.owner.nbytes // SIZE_OF[self.dtype]
        host = np.empty(n, dtype=NP_OF[self.dtype])
        if host.nbytes:
            capi.check(capi.lib().xtb_memcpy(C.c_void_p(host.ctypes.data), C.c_void_p(self.owner.ptr), host.nbytes, capi.D2H))
        return HostArray(self.shape, self.strides, self.offset, self.dtype, host).numpy()


# ---- lowering: expression tree -> postfix program + leaf descriptors --------------------------
def _imm_bits(value, rt: int) -> int:
    v = NP_OF[rt](value)
    raw = v.tobytes()
    return int.from_bytes(raw, "little")


class Lowered:
    def __init__(self):
        self.insns: List[Tuple[int, int, int, int]] = []
        self.leaves: List[Array] = []
        self.imms: List[int] = []

    def leaf_index(self, a: Array) -> int:
        for i, l in enumerate(self.leaves):
            if l is a:
                return i
        if len(self.leaves) >= capi.MAX_LEAVES:
            raise RuntimeError("too many tensor leaves in one expression")
        self.leaves.append(a)
        return len(self.leaves) - 1

    def imm_index(self, bits: int) -> int:
        if len(self.imms) >= capi.MAX_IMMS:
            raise RuntimeError("too many scalar immediates in one expression")
        self.imms.append(bits)
        return len(self.imms) - 1

    def emit(self, op, type_, src=0, arg=0):
        if len(self.insns) >= capi.MAX_INSNS:
            raise RuntimeError("expression too long")
        self.insns.append((OPCODES[op], type_, src, arg))

    def program(self) -> capi.Program:
        p = capi.Program()
        p.n_insns, p.n_leaves, p.n_imms = len(self.insns), len(self.leaves), len(self.imms)
        for i, (op, t, s, a) in enumerate(self.insns):
            p.insns[i] = capi.Insn(op, t, s, a)
        for i, b in enumerate(self.imms):
            p.imms[i] = b
        return p

    def operands(self):
        arr = (capi.Operand * max(1, len(self.leaves)))()
        for i, l in enumerate(self.leaves):
            arr[i] = l.operand()
        return arr


def _materialise(e: Expr) -> Expr:
    """Reducers nested in a tree are evaluated into temporaries first (like eval())."""
    if isinstance(e, Reducer):
        return evaluate(e)
    if isinstance(e, Func):
        args = tuple(_materialise(a) for a in e.args)
        if any(a is not b for a, b in zip(args, e.args)):
            return Func(e.op, args, e.cast_to)
    return e


def _is_simple(e: Expr, t: int) -> bool:
    """Operand that a binary instruction can fetch itself: a scalar, or a leaf stored as t."""
    return isinstance(e, Scalar) or (isinstance(e, Array) and e.dtype == t and t >= I32)


def _emit_value(lw: Lowered, e: Expr, want: Optional[int]) -> int:
    """Emit code leaving e's value (register type) on the stack; cast to `want` if given."""
    if isinstance(e, Scalar):
        rt = regtype(e.dtype) if want is None else want
        lw.emit("PUSH", rt, SRC_IMM, lw.imm_index(_imm_bits(e.value, rt)))
        return rt
    if isinstance(e, Array):
        lw.emit("PUSH", e.dtype, SRC_LEAF, lw.leaf_index(e))
        rt = regtype(e.dtype)
    elif isinstance(e, Func):
        rt = _emit_func(lw, e)
    else:
        raise TypeError(f"cannot lower {type(e).__name__}")
    if want is not None and rt != want:
        lw.emit("CAST", rt, 0, want)
        rt = want
    return rt


def _fused_src(lw: Lowered, e: Expr, t: int) -> Tuple[int, int]:
    if isinstance(e, Scalar):
        return SRC_IMM, lw.imm_index(_imm_bits(e.value, t))
    return SRC_LEAF, lw.leaf_index(e)


def _emit_func(lw: Lowered, f: Func) -> int:
    op, t = f.op, f.operand_type
    if len(f.args) == 1:
        _emit_value(lw, f.args[0], t)
        if op == "CAST":
            lw.emit("CAST", t, 0, f.cast_to)
            return regtype(f.cast_to)
        lw.emit(op, t)
        return I32 if op in PRED_UNARY else t
    if len(f.args) == 2:
        a, b = f.args
        if _is_simple(b, t):
            _emit_value(lw, a, t)
            src, arg = _fused_src(lw, b, t)
            lw.emit(op, t, src, arg)
        elif _is_simple(a, t):
            _emit_value(lw, b, t)
            src, arg = _fused_src(lw, a, t)
            lw.emit(op, t, src | SRC_REV, arg)
        else:
            _emit_value(lw, a, t)
            _emit_value(lw, b, t)
            lw.emit(op, t, SRC_STACK, 0)
        return I32 if op in CMP_BINARY else t
    a, b, c = f.args
    if op == "WHERE":
        rt = _emit_value(lw, a, None)
        if rt != I32:                       # condition -> bool
            lw.emit("CAST", rt, 0, BOOL)
    else:
        _emit_value(lw, a, t)
    _emit_value(lw, b, t)
    _emit_value(lw, c, t)
    lw.emit(op, t)
    return t


def lower(e: Expr) -> Lowered:
    lw = Lowered()
    rt = _emit_value(lw, _materialise(as_expr(e)), None)
    # narrow value types (e.g. cast<int8>(x), where(c, u8, u8)) wrap before the store
    if isinstance(e, Func) and e.dtype < I32 and e.dtype != BOOL and e.op != "CAST":
        lw.emit("CAST", rt, 0, e.dtype)
    return lw


# ---- evaluation --------------------------------------------------------------------------------
_HOST_BACKEND = None


def set_host_backend(backend):
    """Tests register the CPU oracle here; the package never imports it."""
    global _HOST_BACKEND
    _HOST_BACKEND = backend


def _kind(leaves: Sequence[Array], out: Optional[Array] = None):
    arrs = list(leaves) + ([out] if out is not None else [])
    kinds = {type(a) for a in arrs}
    if len(kinds) > 1:
        raise TypeError("host and device arrays cannot be mixed in one expression")
    return kinds.pop() if kinds else DeviceArray


def _backend_for(kind):
    if kind is DeviceArray:
        return capi.lib(), "xtb_"
    if _HOST_BACKEND is None:
        raise RuntimeError("no host backend registered: HostArray expressions are evaluated by the test oracle only")
    return _HOST_BACKEND, "xto_"


def _check(kind, code):
    if code == 0:
        return
    if kind is DeviceArray:
        msg = capi.lib().xtb_last_error().decode()
    else:
        msg = _HOST_BACKEND.xto_last_error().decode()
    if code == capi.ERR_SHAPE:
        raise BroadcastError(msg)
    if code == capi.ERR_AXIS:
        raise RuntimeError(msg)
    raise capi.XtbError(code, msg)


def assign(out: Array, e) -> Array:
    """xt::noalias(out) = e: out must already have the broadcast shape (views) ."""
    e = as_expr(e)
    shp = broadcast_shape([e.shape])
    if broadcast_shape([shp, out.shape]) != out.shape:
        raise BroadcastError(f"cannot assign shape {shp} into {out.shape}")
    fr = _finalized_reducer(e)
    if fr is not None and isinstance(out, DeviceArray) and (_leaf_kind(e) or DeviceArray) is DeviceArray and out.shape == fr[0].shape:
        # device: the division is the reduction kernel's epilogue (one launch fewer, same arithmetic)
        _run_reducer(fr[0], DeviceArray, out=out, fin=fr[1])
        return out
    lw = lower(e)
    kind = _kind(lw.leaves, out)
    be, pre = _backend_for(kind)
    prog, ops, oop = lw.program(), lw.operands(), out.operand()
    _check(kind, getattr(be, pre + "assign")(C.byref(prog), C.byref(oop), ops))
    return out


class noalias:
    """xt::noalias(a) = expr; += -= *= /= map to the computed assigns (core/xnoalias.hpp:160-230)."""

    def __init__(self, out: Array):
        self.out = out

    def assign(self, e): return assign(self.out, e)
    def plus_assign(self, e): return assign(self.out, self.out + as_expr(e))
    def minus_assign(self, e): return assign(self.out, self.out - as_expr(e))
    def multiplies_assign(self, e): return assign(self.out, self.out * as_expr(e))
    def divides_assign(self, e): return assign(self.out, self.out / as_expr(e))


def _alloc_like(kind, shape, dtype):
    return kind.empty(shape, dtype)


def _leaf_kind(e: Expr):
    if isinstance(e, Array):
        return type(e)
    if isinstance(e, Func):
        for a in e.args:
            k = _leaf_kind(a)
            if k is not None:
                return k
    if isinstance(e, Reducer):
        return _leaf_kind(e.e)
    return None


def _finalized_reducer(e: Expr):
    """`sum(...) / scalar` (xt::mean, xt::variance: core/xmath.hpp:1827-1852, 2082-2105) and sqrt of it (xt::stddev):
    returns (reducer, capi.Finalize) when the division can be fused into the reduction's last store, else None.
    The fused step computes exactly what the separate kernel would: T(sum) / T(divisor) in the node's value type."""
    fin_op = capi.FIN_DIV
    if isinstance(e, Func) and e.op == "SQRT" and isinstance(e.args[0], Func) and e.dtype == e.args[0].dtype:
        e, fin_op = e.args[0], capi.FIN_DIV_SQRT
    if not (isinstance(e, Func) and e.op == "DIV" and isinstance(e.args[0], Reducer) and isinstance(e.args[1], Scalar)):
        return None
    r, t = e.args[0], e.dtype
    if t not in (F32, F64) or r.op != capi.RED_SUM or r.acc < I32:
        return None
    fin = capi.Finalize(fin_op, t, _imm_bits(e.args[1].value, t))
    return r, fin


def evaluate(e, dtype: Optional[int] = None) -> Array:
    """xt::eval / container construction from an expression: allocate + assign."""
    e = as_expr(e)
    if isinstance(e, Array) and dtype is None:
        return e
    kind = _leaf_kind(e) or DeviceArray
    if isinstance(e, Reducer):
        return _run_reducer(e, kind, dtype)
    out = _alloc_like(kind, e.shape, e.dtype if dtype is None else dtype)
    return assign(out, e)


def _run_reducer(r: Reducer, kind, out_dtype: Optional[int] = None, allreduce: bool = False, mode: int = 0,
                 out: Optional[Array] = None, fin=None) -> Array:
    inner = _materialise(r.e)
    lw = Lowered()
    _emit_value(lw, inner, None)
    if out is None:
        out = _alloc_like(kind, r.shape, r.dtype if out_dtype is None else out_dtype)
    prog, ops, oop = lw.program(), lw.operands(), out.operand()
    nd = len(r.e.shape)
    shape = (C.c_int64 * max(nd, 1))(*r.e.shape)
    axes = (C.c_int32 * max(len(r.axes), 1))(*r.axes)
    init_p = None
    if r.initial is not None:
        buf = NP_OF[r.acc](r.initial).tobytes()
        init_p = C.cast(C.create_string_buffer(buf, len(buf)), C.c_void_p)
    be, pre = _backend_for(kind)
    last = int(allreduce) if kind is DeviceArray else int(mode)
    if fin is not None:
        assert kind is DeviceArray
        _check(kind, be.xtb_reduce_fin(r.op, r.acc, C.byref(prog), ops, nd, shape, len(r.axes), axes,
                                       int(r.keep_dims), init_p, C.byref(oop), last, C.byref(fin)))
        return out
    _check(kind, getattr(be, pre + "reduce")(r.op, r.acc, C.byref(prog), ops, nd, shape, len(r.axes), axes,
                                             int(r.keep_dims), init_p, C.byref(oop), last))
    return out


# ---- free functions (same names as xt::) --------------------------------------------------------
def _unary(op):
    def f(e):
        return Func(op, (as_expr(e),))
    f.__name__ = op.lower()
    return f


abs = fabs = _unary("ABS")
exp, exp2, expm1, log, log10, log2, log1p = map(_unary, ["EXP", "EXP2", "EXPM1", "LOG", "LOG10", "LOG2", "LOG1P"])
sqrt, cbrt = _unary("SQRT"), _unary("CBRT")
sin, cos, tan, asin, acos, atan = map(_unary, ["SIN", "COS", "TAN", "ASIN", "ACOS", "ATAN"])
sinh, cosh, tanh, asinh, acosh, atanh = map(_unary, ["SINH", "COSH", "TANH", "ASINH", "ACOSH", "ATANH"])
erf, erfc, tgamma, lgamma = map(_unary, ["ERF", "ERFC", "TGAMMA", "LGAMMA"])
ceil, floor, trunc, round, nearbyint, rint = map(_unary, ["CEIL", "FLOOR", "TRUNC", "ROUND", "NEARBYINT", "RINT"])
isfinite, isinf, isnan = map(_unary, ["ISFINITE", "ISINF", "ISNAN"])
sign, deg2rad, rad2deg, square, cube = map(_unary, ["SIGN", "DEG2RAD", "RAD2DEG", "SQUARE", "CUBE"])
logical_not = _unary("NOT")


def _binary(op):
    def f(a, b):
        return Func(op, (as_expr(a), as_expr(b)))
    f.__name__ = op.lower()
    return f


fmod, remainder, fmax, fmin, fdim, pow, hypot, atan2 = map(
    _binary, ["FMOD", "REMAINDER", "FMAX", "FMIN", "FDIM", "POW", "HYPOT", "ATAN2"])
maximum, minimum = _binary("MAXIMUM"), _binary("MINIMUM")
logical_or, logical_and = _binary("LOR"), _binary("LAND")
equal, not_equal = _binary("EQ"), _binary("NE")


def where(c, a, b): return Func("WHERE", (as_expr(c), as_expr(a), as_expr(b)))
def fma(a, b, c): return Func("FMA", (as_expr(a), as_expr(b), as_expr(c)))
def clip(e, lo, hi): return Func("CLAMP", (as_expr(e), as_expr(lo), as_expr(hi)))
def cast(e, dtype: int): return Func("CAST", (as_expr(e),), cast_to=dtype)
def transpose(a: Array, perm=None): return a.transpose(perm)
def broadcast(a: Array, shape): return a.broadcast(shape)
def view(a: Array, *slices): return a[tuple(slices)]
def eval(e, dtype=None): return evaluate(e, dtype)


def sum(e, axes=None, keep_dims=False, initial=None, dtype=None):
    return Reducer(capi.RED_SUM, e, axes, keep_dims, initial, dtype)


def prod(e, axes=None, keep_dims=False, initial=None, dtype=None):
    return Reducer(capi.RED_PROD, e, axes, keep_dims, initial, dtype)


def amax(e, axes=None, keep_dims=False, initial=None):
    return Reducer(capi.RED_MAX, e, axes, keep_dims, initial)


def amin(e, axes=None, keep_dims=False, initial=None):
    return Reducer(capi.RED_MIN, e, axes, keep_dims, initial)


def _size(shape):
    return int(np.prod(shape, dtype=np.int64)) if len(shape) else 1


def _mean_divisor(e, s_shape, axes, dtype, ddof):
    n_out = _size(s_shape)
    vt = F64 if dtype is None else dtype
    if axes is None:
        div = NP_OF[vt](_size(e.shape) - ddof)
    else:
        div = NP_OF[vt]((_size(e.shape) - ddof) // n_out) if n_out else NP_OF[vt](0)
    return Scalar(div, vt)


def mean(e, axes=None, dtype=None, ddof=0, keep_dims=False):
    """sum<T>(e, axes) / static_cast<V>(size / result_size), V = double unless T given
    (core/xmath.hpp:1827-1852).  ddof is subtracted from e.size() first (:1848-1851)."""
    e = as_expr(e)
    s = sum(e, axes, keep_dims=keep_dims, dtype=dtype)
    return s / _mean_divisor(e, s.shape, axes, dtype, ddof)


def average(e, weights=None, axes=None, dtype=None):
    """xt::average (core/xmath.hpp:1925-2010): sum<T>(e * w, axes) / sum<T>(w, axes, immediate), `w` 1-d along the
    first given axis or of e's shape; without weights it is mean<T>(e)."""
    e = as_expr(e)
    if weights is None:
        return mean(e, axes, dtype=dtype)
    w = weights
    if axes is None:
        if tuple(w.shape) != tuple(e.shape):
            raise RuntimeError("Weights need to have the same shape as expression.")
        div = _run_reducer(sum(w, dtype=dtype), _leaf_kind(w) or DeviceArray, mode=1)
        return sum(e * w, dtype=dtype) / div
    nd = len(e.shape)
    ax = [axes] if isinstance(axes, (int, np.integer)) else list(axes)
    ax = [int(a) + nd if int(a) < 0 else int(a) for a in ax]
    if len(w.shape) == 1:
        if w.size != e.shape[ax[0]]:
            raise RuntimeError("Weights need to have the same shape as expression at axes.")
        bshape = [1] * nd
        bshape[ax[0]] = w.size
    else:
        if tuple(w.shape) != tuple(e.shape):
            raise RuntimeError("Weights with dim > 1 need to have the same shape as expression.")
        bshape = list(e.shape)
    wv = w.reshape_view(bshape)
    scl = _run_reducer(sum(wv, ax, dtype=dtype), _leaf_kind(w) or DeviceArray, mode=1)
    return sum(e * wv, ax, dtype=dtype) / scl


def variance(e, axes=None, dtype=None, ddof=0):
    """Two-pass, as the reference (core/xmath.hpp:2082-2105): inner_mean =
    eval(mean<T>(e, axes, immediate)) -- the *immediate* evaluation order --, reshaped with
    reduced dims = 1, then the lazy mean<T>(square(e - inner_mean), axes, ddof)."""
    e = as_expr(e)
    nd = len(e.shape)
    ax = list(range(nd)) if axes is None else ([axes] if isinstance(axes, (int, np.integer)) else list(axes))
    ax = [a + nd if a < 0 else a for a in ax]
    s = sum(e, ax, dtype=dtype)
    s_arr = _run_reducer(s, _leaf_kind(e) or DeviceArray, mode=1)
    inner_mean = evaluate(s_arr / _mean_divisor(e, s.shape, ax, dtype, 0))
    keep = [1 if d in ax else sh for d, sh in enumerate(e.shape)]
    mrv = inner_mean.reshape_view(keep)
    return mean(square(e - mrv), ax, dtype=dtype, ddof=ddof)


def stddev(e, axes=None, dtype=None, ddof=0):
    return sqrt(variance(e, axes, dtype, ddof))


# ---- nan-aware reducers, counts (core/xmath.hpp:2307-2860) ---------------------------------------
# The reference builds these from custom reducer functors (nan_plus: "!isnan(rhs) ? lhs + rhs : lhs", ...).
# Here they are the SAME values written as fused map-reduce programs over the existing merge operators, so
# they run on the map-reduce kernels with one pass over the data:
#   nan_plus over x        ==  plus over where(isnan(x), 0, x)       (acc + 0 == acc: the accumulator starts at
#                                                                      +0 and can never be -0)
#   nan_multiplies over x  ==  multiplies over where(isnan(x), 1, x)
#   count_nonzero          ==  plus over size_t(x != 0)
def _const_like(e: Expr, v) -> Scalar:
    return Scalar(NP_OF[e.dtype](v), e.dtype)


def nan_to_num(e):
    """nan -> 0, +inf -> max, -inf -> lowest (detail::nan_to_num_functor, core/xmath.hpp:2309-2331)."""
    e = as_expr(e)
    if e.dtype not in (F32, F64):
        return e
    fi = np.finfo(NP_OF[e.dtype])
    return where(isnan(e), _const_like(e, 0), where(isinf(e), where(e < _const_like(e, 0), _const_like(e, fi.min), _const_like(e, fi.max)), e))


def nansum(e, axes=None, keep_dims=False, dtype=None):
    e = as_expr(e)
    return sum(where(isnan(e), _const_like(e, 0), e), axes, keep_dims, dtype=dtype)


def nanprod(e, axes=None, keep_dims=False, dtype=None):
    e = as_expr(e)
    return prod(where(isnan(e), _const_like(e, 1), e), axes, keep_dims, dtype=dtype)


def count_nonzero(e, axes=None, keep_dims=False):
    """result type std::size_t (xreducer_size_type_t, reducers/xreducer.hpp:1178-1185)."""
    e = as_expr(e)
    return sum(cast(e.ne(_const_like(e, 0)), U64), axes, keep_dims)


def count_nonnan(e, axes=None, keep_dims=False):
    """count_nonzero(!isnan(e)) (core/xmath.hpp:2540-2570)."""
    return count_nonzero(logical_not(isnan(as_expr(e))), axes, keep_dims)


def nanmin(e, axes=None, keep_dims=False):
    """XTENSOR_REDUCER_FUNCTION(nanmin, detail::nan_min, value_type, std::nan("0")) (core/xmath.hpp:2333-2346, 2427):
    starts from NaN and skips NaN operands -- the minimum of the non-NaN values, NaN where there is none."""
    return Reducer(capi.RED_NANMIN, as_expr(e), axes, keep_dims)


def nanmax(e, axes=None, keep_dims=False):
    """detail::nan_max (core/xmath.hpp:2348-2363, 2442)."""
    return Reducer(capi.RED_NANMAX, as_expr(e), axes, keep_dims)


def minmax(e):
    """xt::minmax (core/xmath.hpp:2195-2228): {min, max} over the whole expression (a reducer over every axis whose
    value type is std::array<value_type, 2>); returned as a 2-element array of e's dtype."""
    e = as_expr(e)
    kind = _leaf_kind(e) or DeviceArray
    out = _alloc_like(kind, (2,), e.dtype)
    nd = len(e.shape)
    # r[0] = std::min(r[0], v), r[1] = std::max(r[1], v) from {max(), lowest()}: std::min / std::max never take a NaN
    # operand, i.e. the NaN-skipping extremes merged once with the init
    npt = NP_OF[e.dtype]
    lim = np.finfo(npt) if e.dtype in (F32, F64) else np.iinfo(npt if e.dtype != BOOL else np.uint8)
    acc = NP_OF[regtype(e.dtype)]
    for k, (op, init) in enumerate(((capi.RED_NANMIN, lim.max), (capi.RED_NANMAX, lim.min))):
        _run_reducer(Reducer(op, e, list(range(nd)), initial=acc(init)), kind, out=out[k])
    return out


def _arg(op, e, axis):
    """xt::argmin / xt::argmax (misc/xsort.hpp:1237-1295): eval(e), then the first extreme along `axis` (or of the
    flattened row-major traversal); result std::size_t."""
    a = evaluate(as_expr(e))
    kind = type(a)
    if axis is None:
        if a.strides != compute_strides(a.shape):
            a = assign(_alloc_like(kind, a.shape, a.dtype), a)      # eval() of a strided view is a dense copy
        out = _alloc_like(kind, (), U64)
        ax = -1
    else:
        nd = len(a.shape)
        ax = axis + nd if axis < 0 else axis
        if not 0 <= ax < nd:
            raise RuntimeError(f"Axis {axis} out of bounds")
        out = _alloc_like(kind, tuple(s for d, s in enumerate(a.shape) if d != ax), U64)
    be, pre = _backend_for(kind)
    iop, oop = a.operand(), out.operand()
    _check(kind, getattr(be, pre + "argreduce")(op, C.byref(iop), ax, C.byref(oop)))
    return out


def argmin(e, axis=None): return _arg(capi.RED_MIN, e, axis)
def argmax(e, axis=None): return _arg(capi.RED_MAX, e, axis)


def nanmean(e, axes=None, dtype=None, keep_dims=False):
    """nansum<sum_type>(e, axes) / cast<value_type>(count_nonnan(e, axes)); value_type = double unless T is
    given, sum_type = common_type(E::value_type, double) unless T is given (core/xmath.hpp:2694-2712)."""
    e = as_expr(e)
    vt = F64 if dtype is None else dtype
    st = common_type(e.dtype, F64) if dtype is None else dtype
    return nansum(e, axes, keep_dims, dtype=st) / cast(count_nonnan(e, axes, keep_dims), vt)


def nanvar(e, axes=None, dtype=None):
    """nanmean<R>(square(cast<R>(e) - reshape_view(nanmean<R>(e, axes), keep_dims shape)), axes), R = double
    unless T is given (core/xmath.hpp:2785-2808)."""
    e = as_expr(e)
    nd = len(e.shape)
    ax = list(range(nd)) if axes is None else ([axes] if isinstance(axes, (int, np.integer)) else list(axes))
    ax = [a + nd if a < 0 else a for a in ax]
    rt = F64 if dtype is None else dtype
    inner = evaluate(nanmean(e, ax, dtype=rt))
    keep = [1 if d in ax else sh for d, sh in enumerate(e.shape)]
    ce = e if e.dtype == rt else cast(e, rt)
    return nanmean(square(ce - inner.reshape_view(keep)), ax, dtype=rt)


def nanstd(e, axes=None, dtype=None):
    return sqrt(nanvar(e, axes, dtype))


# ---- norms (reducers/xnorm.hpp:369-620) ---------------------------------------------------------------
# reducers  r + g(v)  (merge std::plus) or  std::max(r, |v|)  (merge math::maximum) from 0, with
#   l0: g = (v != 0), result unsigned long long        l1: g = std::abs(v), result big_promote_type_t<V>
#   sq: g = v * v (in V's promoted type), same result   linf: result decltype(std::abs(v)) (V itself when unsigned)
#   lp_to_p: g = pow(rt(|v|), rt(p)), rt = real_promote_type_t<V>, result norm_type_t = double
#   l2 = sqrt(sq), lp = pow(lp_to_p, 1 / p)
def _big_promote(dt: int) -> int:
    if dt in (F32, F64):
        return F64
    return U64 if dt in (BOOL, U8, U16, U32, U64) else I64


def _is_unsigned(dt: int) -> bool:
    return dt in (BOOL, U8, U16, U32, U64)


def _abs_v(e: Expr) -> Expr:
    return e if _is_unsigned(e.dtype) else abs(e)


def norm_l0(e, axes=None, keep_dims=False):
    e = as_expr(e)
    return sum(cast(e.ne(_const_like(e, 0)), U64), axes, keep_dims)


def norm_l1(e, axes=None, keep_dims=False):
    e = as_expr(e)
    return sum(cast(_abs_v(e), _big_promote(e.dtype)), axes, keep_dims)


def norm_sq(e, axes=None, keep_dims=False):
    e = as_expr(e)
    return sum(cast(e * e, _big_promote(e.dtype)), axes, keep_dims)


def norm_l2(e, axes=None, keep_dims=False):
    return sqrt(norm_sq(e, axes, keep_dims))


def norm_linf(e, axes=None, keep_dims=False):
    """std::max<result_type>(r, std::abs(v)) never takes a NaN operand: the NaN-skipping maximum, merged once with
    the reducer's own init 0."""
    e = as_expr(e)
    a = _abs_v(e)
    return Reducer(capi.RED_NANMAX, a, axes, keep_dims, initial=NP_OF[regtype(a.dtype)](0))


def norm_lp_to_p(e, p, axes=None, keep_dims=False):
    e = as_expr(e)
    if p == 0:
        return sum(cast(e.ne(_const_like(e, 0)), F64), axes, keep_dims)
    rt = e.dtype if e.dtype in (F32, F64) else F64
    a = _abs_v(e)
    a = a if a.dtype == rt else cast(a, rt)
    g = pow(a, Scalar(NP_OF[rt](p), rt))
    return sum(g if rt == F64 else cast(g, F64), axes, keep_dims)


def norm_lp(e, p, axes=None, keep_dims=False):
    if p == 0:
        raise ValueError("norm_lp(): p must be nonzero, use norm_l0() instead.")
    return pow(norm_lp_to_p(e, p, axes, keep_dims), Scalar(np.float64(1.0 / p), F64))


def _scan(op, e, axis, dtype, out=None):
    a = evaluate(e)
    kind = type(a)
    init_t = a.dtype if dtype is None else dtype
    acc = common_type(init_t, a.dtype)          # decltype(T() + x) (xaccumulator.hpp:224-234)
    if axis is None:
        out = _alloc_like(kind, (a.size,), acc) if out is None else out
        ax = -1
    else:
        ax = int(axis) + a.ndim if int(axis) < 0 else int(axis)
        if ax >= a.ndim:
            raise RuntimeError("Axis larger than expression dimension in accumulator.")
        out = _alloc_like(kind, a.shape, acc) if out is None else out
    be, pre = _backend_for(kind)
    iop, oop = a.operand(), out.operand()
    _check(kind, getattr(be, pre + "scan")(op, acc, C.byref(iop), ax, C.byref(oop)))
    return out


def cumsum(e, axis=None, dtype=None, out=None): return _scan(capi.RED_SUM, e, axis, dtype, out)
def cumprod(e, axis=None, dtype=None, out=None): return _scan(capi.RED_PROD, e, axis, dtype, out)


def sync():
    capi.check(capi.lib().xtb_sync())
