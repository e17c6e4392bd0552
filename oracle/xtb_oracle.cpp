// xtb_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code).
//
// A plain, single-threaded C++ restatement of xtensor 0.27.1's *scalar*
// (non-xsimd) evaluation of the hot path, used only by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs
// as the checker.  Nothing under xtensor_b200/ or include/ may call it.
//
// It takes the same descriptors as libxtb200 (include/xtb200.h) but with HOST
// pointers, and follows these reference loops:
//   xto_assign : stepper_assigner::run            include/xtensor/core/xassign.hpp:644-695
//                stepper_tools::increment_stepper include/xtensor/core/xiterator.hpp:589-631
//                xstepper::step / reset           include/xtensor/core/xiterator.hpp:489-525
//                (row-major odometer; every leaf steps by stride[dim - offset] and
//                 resets by backstride; leading dims a leaf lacks are no-ops; extent-1
//                 dims have stride 0, include/xtensor/core/xstrides.hpp:503-530)
//                functor semantics            include/xtensor/core/xoperation.hpp:30-164,
//                                             include/xtensor/core/xmath.hpp:82-866
//                store cast                   include/xtensor/core/xassign.hpp:613-667
//   xto_reduce : mode 0 (lazy)      xreducer_stepper::aggregate_impl
//                                   include/xtensor/reducers/xreducer.hpp:1778-1868
//                mode 1 (immediate) reduce_immediate  :289-565 (operates on eval(e))
//   xto_scan   : accumulator_impl   include/xtensor/reducers/xaccumulator.hpp:215-341
//
// Transcendentals are glibc libm (xtensor's scalar path calls std::sin etc.,
// xmath.hpp:150-215); the xsimd path uses different approximations and is NOT
// what this oracle models.  Build with -O2 -ffp-contract=off -fno-fast-math so
// that no FMA contraction or reassociation happens.
//
// Parity pinning: tests/test_oracle_golden.py checks this file against the
// literal expectations of the reference's own tests (tests/golden/*.json,
// transcribed from test/test_xreducer.cpp, test_xaccumulator.cpp, test_xnoalias /
// test_xsemantic fixtures, test_extended_broadcast_view.cpp) and against outputs
// of the real reference headers compiled here (oracle/_ref, see oracle/Makefile).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>
#include "../include/xtb200.h"

namespace {

thread_local char g_err[256] = "";
int fail(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

int dsize(int dt) {
    switch (dt) {
        case XTB_BOOL: case XTB_I8: case XTB_U8: return 1;
        case XTB_I16: case XTB_U16: return 2;
        case XTB_I32: case XTB_U32: case XTB_F32: return 4;
        default: return 8;
    }
}
int regtype(int dt) { return dt < XTB_I32 ? (int) XTB_I32 : dt; }

// A typed scalar: exactly one member is meaningful, selected by `t`.
struct Val {
    int t;
    union {
        int32_t i32; uint32_t u32; int64_t i64; uint64_t u64; float f32; double f64;
    };
};

Val load(const char* p, int dt) {
    Val v;
    v.t = regtype(dt);
    v.u64 = 0;
    switch (dt) {
        case XTB_BOOL: v.i32 = *(const uint8_t*) p != 0; break;
        case XTB_I8: v.i32 = *(const int8_t*) p; break;
        case XTB_U8: v.i32 = *(const uint8_t*) p; break;
        case XTB_I16: v.i32 = *(const int16_t*) p; break;
        case XTB_U16: v.i32 = *(const uint16_t*) p; break;
        case XTB_I32: v.i32 = *(const int32_t*) p; break;
        case XTB_U32: v.u32 = *(const uint32_t*) p; break;
        case XTB_I64: v.i64 = *(const int64_t*) p; break;
        case XTB_U64: v.u64 = *(const uint64_t*) p; break;
        case XTB_F32: v.f32 = *(const float*) p; break;
        case XTB_F64: v.f64 = *(const double*) p; break;
    }
    return v;
}

template <class F> auto visit(const Val& v, F&& f) {
    switch (v.t) {
        case XTB_I32: return f(v.i32);
        case XTB_U32: return f(v.u32);
        case XTB_I64: return f(v.i64);
        case XTB_U64: return f(v.u64);
        case XTB_F32: return f(v.f32);
        default: return f(v.f64);
    }
}
Val mk(int32_t x) { Val v; v.u64 = 0; v.t = XTB_I32; v.i32 = x; return v; }
Val mk(uint32_t x) { Val v; v.u64 = 0; v.t = XTB_U32; v.u32 = x; return v; }
Val mk(int64_t x) { Val v; v.u64 = 0; v.t = XTB_I64; v.i64 = x; return v; }
Val mk(uint64_t x) { Val v; v.u64 = 0; v.t = XTB_U64; v.u64 = x; return v; }
Val mk(float x) { Val v; v.u64 = 0; v.t = XTB_F32; v.f32 = x; return v; }
Val mk(double x) { Val v; v.u64 = 0; v.t = XTB_F64; v.f64 = x; return v; }
Val mk(bool x) { return mk((int32_t) x); }

// static_cast<dtype>(v), result widened back to its register type
Val cast_to(const Val& v, int dt) {
    return visit(v, [dt](auto x) -> Val {
        switch (dt) {
            case XTB_BOOL: return mk((int32_t) (x != 0));
            case XTB_I8: return mk((int32_t) (int8_t) x);
            case XTB_U8: return mk((int32_t) (uint8_t) x);
            case XTB_I16: return mk((int32_t) (int16_t) x);
            case XTB_U16: return mk((int32_t) (uint16_t) x);
            case XTB_I32: return mk((int32_t) x);
            case XTB_U32: return mk((uint32_t) x);
            case XTB_I64: return mk((int64_t) x);
            case XTB_U64: return mk((uint64_t) x);
            case XTB_F32: return mk((float) x);
            default: return mk((double) x);
        }
    });
}

void store(char* p, int dt, const Val& v) {
    visit(v, [p, dt](auto x) {
        switch (dt) {
            case XTB_BOOL: *(uint8_t*) p = (uint8_t) (x != 0); break;
            case XTB_I8: *(int8_t*) p = (int8_t) x; break;
            case XTB_U8: *(uint8_t*) p = (uint8_t) x; break;
            case XTB_I16: *(int16_t*) p = (int16_t) x; break;
            case XTB_U16: *(uint16_t*) p = (uint16_t) x; break;
            case XTB_I32: *(int32_t*) p = (int32_t) x; break;
            case XTB_U32: *(uint32_t*) p = (uint32_t) x; break;
            case XTB_I64: *(int64_t*) p = (int64_t) x; break;
            case XTB_U64: *(uint64_t*) p = (uint64_t) x; break;
            case XTB_F32: *(float*) p = (float) x; break;
            case XTB_F64: *(double*) p = (double) x; break;
        }
        return 0;
    });
}

Val imm(uint64_t bits, int rt) {
    Val v;
    v.t = rt;
    v.u64 = 0;
    switch (rt) {
        case XTB_I32: case XTB_U32: case XTB_F32: { uint32_t lo = (uint32_t) bits; memcpy(&v.u32, &lo, 4); break; }
        default: v.u64 = bits; break;
    }
    return v;
}

template <class T> T sign_of(T x) {  // math::sign_impl xmath.hpp:826-852
    if constexpr (std::is_floating_point<T>::value) {
        return std::isnan(x) ? std::numeric_limits<T>::quiet_NaN() : x == 0 ? T(std::copysign(T(0), x)) : T(std::copysign(T(1), x));
    } else if constexpr (std::is_signed<T>::value) {
        return x == 0 ? T(0) : (x < 0 ? T(-1) : T(1));
    } else {
        return T(x > T(0));
    }
}

template <class T> Val unary_float(int op, T x) {
    constexpr T PI = (T) 3.141592653589793238463;  // xt::numeric_constants<T>::PI xmath.hpp:41
    switch (op) {
        case XTB_OP_NEG: return mk((T) -x);
        case XTB_OP_ABS: return mk((T) std::fabs(x));
        case XTB_OP_EXP: return mk((T) std::exp(x));
        case XTB_OP_EXP2: return mk((T) std::exp2(x));
        case XTB_OP_EXPM1: return mk((T) std::expm1(x));
        case XTB_OP_LOG: return mk((T) std::log(x));
        case XTB_OP_LOG10: return mk((T) std::log10(x));
        case XTB_OP_LOG2: return mk((T) std::log2(x));
        case XTB_OP_LOG1P: return mk((T) std::log1p(x));
        case XTB_OP_SQRT: return mk((T) std::sqrt(x));
        case XTB_OP_CBRT: return mk((T) std::cbrt(x));
        case XTB_OP_SIN: return mk((T) std::sin(x));
        case XTB_OP_COS: return mk((T) std::cos(x));
        case XTB_OP_TAN: return mk((T) std::tan(x));
        case XTB_OP_ASIN: return mk((T) std::asin(x));
        case XTB_OP_ACOS: return mk((T) std::acos(x));
        case XTB_OP_ATAN: return mk((T) std::atan(x));
        case XTB_OP_SINH: return mk((T) std::sinh(x));
        case XTB_OP_COSH: return mk((T) std::cosh(x));
        case XTB_OP_TANH: return mk((T) std::tanh(x));
        case XTB_OP_ASINH: return mk((T) std::asinh(x));
        case XTB_OP_ACOSH: return mk((T) std::acosh(x));
        case XTB_OP_ATANH: return mk((T) std::atanh(x));
        case XTB_OP_ERF: return mk((T) std::erf(x));
        case XTB_OP_ERFC: return mk((T) std::erfc(x));
        case XTB_OP_TGAMMA: return mk((T) std::tgamma(x));
        case XTB_OP_LGAMMA: return mk((T) std::lgamma(x));
        case XTB_OP_CEIL: return mk((T) std::ceil(x));
        case XTB_OP_FLOOR: return mk((T) std::floor(x));
        case XTB_OP_TRUNC: return mk((T) std::trunc(x));
        case XTB_OP_ROUND: return mk((T) std::round(x));
        case XTB_OP_NEARBYINT: return mk((T) std::nearbyint(x));
        case XTB_OP_RINT: return mk((T) std::rint(x));
        case XTB_OP_ISFINITE: return mk((bool) std::isfinite(x));
        case XTB_OP_ISINF: return mk((bool) std::isinf(x));
        case XTB_OP_ISNAN: return mk((bool) std::isnan(x));
        case XTB_OP_NOT: return mk((bool) !x);
        case XTB_OP_SIGN: return mk(sign_of(x));
        case XTB_OP_DEG2RAD: return mk((T) (x * PI / T(180.0)));   // xmath.hpp:626-631
        case XTB_OP_RAD2DEG: return mk((T) (x * T(180.0) / PI));   // xmath.hpp:653-658
        case XTB_OP_SQUARE: return mk((T) (x * x));                // xmath.hpp:1100-1108
        case XTB_OP_CUBE: return mk((T) (x * x * x));              // xmath.hpp:1118-1127
    }
    return mk(x);
}
template <class T> Val unary_int(int op, T x) {
    switch (op) {
        case XTB_OP_NEG: return mk((T) (T(0) - x));
        case XTB_OP_BITNOT: return mk((T) ~x);
        case XTB_OP_NOT: return mk((bool) !x);
        case XTB_OP_ABS:
            if constexpr (std::is_signed<T>::value) return mk((T) (x < 0 ? T(0) - x : x));
            else return mk(x);
        case XTB_OP_SIGN: return mk(sign_of(x));
        case XTB_OP_SQUARE: return mk((T) (x * x));
        case XTB_OP_CUBE: return mk((T) (x * x * x));
        case XTB_OP_ISFINITE: return mk(true);
        case XTB_OP_ISINF: return mk(false);
        case XTB_OP_ISNAN: return mk(false);
    }
    return mk(x);
}

Val unary(int op, const Val& a, int arg) {
    if (op == XTB_OP_CAST) return cast_to(a, arg);
    if (op == XTB_OP_ORDKEY) {      // see xtb200.h: order key of a 32-bit value in the high half of a u64
        uint32_t k;
        if (a.t == XTB_F32) {
            if (a.f32 != a.f32) k = 0xffffffffu;
            else {
                const float z = a.f32 + 0.0f;
                uint32_t b;
                std::memcpy(&b, &z, 4);
                k = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
                if (arg) k = ~k;
            }
        } else if (a.t == XTB_I32) {
            k = (uint32_t) a.i32 ^ 0x80000000u;
            if (arg) k = ~k;
        } else {
            k = a.u32;
            if (arg) k = ~k;
        }
        return mk((uint64_t) k << 32);
    }
    return visit(a, [op](auto x) -> Val {
        using T = decltype(x);
        if constexpr (std::is_floating_point<T>::value) return unary_float<T>(op, x);
        else return unary_int<T>(op, x);
    });
}

template <class T> Val binary_t(int op, T x, T y) {
    switch (op) {
        case XTB_OP_ADD: return mk((T) (x + y));
        case XTB_OP_SUB: return mk((T) (x - y));
        case XTB_OP_MUL: return mk((T) (x * y));
        case XTB_OP_DIV:
            if constexpr (std::is_floating_point<T>::value) return mk((T) (x / y));
            else return mk((T) (y == 0 ? T(0) : x / y));
        case XTB_OP_LT: return mk(x < y);
        case XTB_OP_LE: return mk(x <= y);
        case XTB_OP_GT: return mk(x > y);
        case XTB_OP_GE: return mk(x >= y);
        case XTB_OP_EQ: return mk(x == y);
        case XTB_OP_NE: return mk(x != y);
        case XTB_OP_LOR: return mk((bool) (x || y));
        case XTB_OP_LAND: return mk((bool) (x && y));
        case XTB_OP_MAXIMUM: return mk((T) (x > y ? x : y));  // xtl::select(t1 > t2, t1, t2) xmath.hpp:588-602
        case XTB_OP_MINIMUM: return mk((T) (x < y ? x : y));  // xmath.hpp:570-586
        // detail::nan_min / nan_max, xmath.hpp:2333-2363
        case XTB_OP_NANMIN:
            if constexpr (std::is_floating_point<T>::value) return mk((T) (std::isnan(x) ? y : (std::isnan(y) ? x : (x < y ? x : y))));
            else return mk((T) (x < y ? x : y));
        case XTB_OP_NANMAX:
            if constexpr (std::is_floating_point<T>::value) return mk((T) (std::isnan(x) ? y : (std::isnan(y) ? x : (x > y ? x : y))));
            else return mk((T) (x > y ? x : y));
    }
    if constexpr (std::is_floating_point<T>::value) {
        switch (op) {
            case XTB_OP_FMOD: return mk((T) std::fmod(x, y));
            case XTB_OP_REMAINDER: return mk((T) std::remainder(x, y));
            case XTB_OP_FMAX: return mk((T) std::fmax(x, y));
            case XTB_OP_FMIN: return mk((T) std::fmin(x, y));
            case XTB_OP_FDIM: return mk((T) std::fdim(x, y));
            case XTB_OP_POW: return mk((T) std::pow(x, y));
            case XTB_OP_HYPOT: return mk((T) std::hypot(x, y));
            case XTB_OP_ATAN2: return mk((T) std::atan2(x, y));
        }
    } else {
        switch (op) {
            case XTB_OP_MOD: return mk((T) (y == 0 ? T(0) : x % y));
            case XTB_OP_BOR: return mk((T) (x | y));
            case XTB_OP_BAND: return mk((T) (x & y));
            case XTB_OP_BXOR: return mk((T) (x ^ y));
            case XTB_OP_SHL: return mk((T) (x << y));
            case XTB_OP_SHR: return mk((T) (x >> y));
        }
    }
    return mk(x);
}
Val binary(int op, const Val& a, const Val& b) {
    return visit(a, [op, &b](auto x) -> Val {
        using T = decltype(x);
        T y = visit(b, [](auto q) { return (T) q; });  // same type by construction
        return binary_t<T>(op, x, y);
    });
}
Val ternary(int op, const Val& a, const Val& b, const Val& c) {
    if (op == XTB_OP_WHERE) return a.i32 ? b : c;
    return visit(a, [op, &b, &c](auto x) -> Val {
        using T = decltype(x);
        T y = visit(b, [](auto q) { return (T) q; });
        T z = visit(c, [](auto q) { return (T) q; });
        if (op == XTB_OP_FMA) {
            if constexpr (std::is_floating_point<T>::value) return mk((T) std::fma(x, y, z));
            else return mk((T) (x * y + z));
        }
        // clamp_fun: select(lo < hi, select(v < lo, lo, select(hi < v, hi, v)), hi)  xmath.hpp:604-617
        return mk((T) ((y < z) ? ((x < y) ? y : ((z < x) ? z : x)) : z));
    });
}

// evaluate the postfix program for one element; leaf_ptr[k] points at the element
Val eval_program(const xtb_program* p, const char* const* leaf_ptr) {
    Val st[XTB_MAX_STACK + 1];
    int n = 0;
    for (int pc = 0; pc < p->n_insns; ++pc) {
        const xtb_insn in = p->insns[pc];
        if (in.op == XTB_OP_PUSH) {
            st[n++] = (in.src == XTB_SRC_LEAF) ? load(leaf_ptr[in.arg], in.type) : imm(p->imms[in.arg], in.type);
        } else if (in.op < XTB_OP_ADD) {
            st[n - 1] = unary(in.op, st[n - 1], in.arg);
        } else if (in.op < XTB_OP_WHERE) {
            const int kind = in.src & 3;
            if (kind == XTB_SRC_STACK) {
                st[n - 2] = binary(in.op, st[n - 2], st[n - 1]);
                --n;
            } else {
                Val y = (kind == XTB_SRC_LEAF) ? load(leaf_ptr[in.arg], in.type) : imm(p->imms[in.arg], in.type);
                st[n - 1] = (in.src & XTB_SRC_REV) ? binary(in.op, y, st[n - 1]) : binary(in.op, st[n - 1], y);
            }
        } else {
            st[n - 3] = ternary(in.op, st[n - 3], st[n - 2], st[n - 1]);
            n -= 2;
        }
    }
    return st[0];
}

// A stepper over one operand in an `ndim`-dimensional iteration (xstepper,
// xiterator.hpp:104-147, 484-587): offset = ndim - operand rank.
struct Stepper {
    const char* p;
    int size;
    int offset;
    int64_t stride[XTB_MAX_DIM], backstride[XTB_MAX_DIM];
    void init(const xtb_operand* op, int ndim, const int64_t* shape) {
        size = dsize(op->dtype);
        p = (const char*) op->base + op->offset * size;
        offset = ndim - op->ndim;
        for (int d = 0; d < op->ndim; ++d) {
            // stride 0 for extent-1 dims (xstrides.hpp:503-530); a leaf of extent 1 is broadcast
            const int64_t s = (op->shape[d] == 1) ? 0 : op->stride[d];
            stride[d] = s;
            backstride[d] = (shape[d + offset] - 1) * s;
        }
    }
    void step(int dim) { if (dim >= offset) p += stride[dim - offset] * size; }
    void reset(int dim) { if (dim >= offset) p -= backstride[dim - offset] * size; }
};

int check_broadcast(const xtb_operand* op, int ndim, const int64_t* shape) {
    if (op->ndim > ndim) return fail(XTB_ERR_SHAPE, "operand rank exceeds iteration rank");
    const int off = ndim - op->ndim;
    for (int d = 0; d < op->ndim; ++d)
        if (op->shape[d] != shape[d + off] && op->shape[d] != 1) return fail(XTB_ERR_SHAPE, "incompatible dimensions");
    return XTB_OK;
}

}  // namespace

extern "C" {

const char* xto_last_error(void) { return g_err; }

// stepper_assigner::run (xassign.hpp:654-667): for i in 0..size: *lhs = cast(*rhs); increment_stepper
int xto_assign(const xtb_program* prog, const xtb_operand* out, const xtb_operand* leaves) {
    const int nd = out->ndim;
    int64_t total = 1;
    for (int d = 0; d < nd; ++d) total *= out->shape[d];
    Stepper ls[XTB_MAX_LEAVES], os;
    for (int k = 0; k < prog->n_leaves; ++k) {
        int r = check_broadcast(&leaves[k], nd, out->shape);
        if (r) return r;
        ls[k].init(&leaves[k], nd, out->shape);
    }
    os.init(out, nd, out->shape);
    int64_t idx[XTB_MAX_DIM] = {0};
    const char* lp[XTB_MAX_LEAVES];
    for (int64_t i = 0; i < total; ++i) {
        for (int k = 0; k < prog->n_leaves; ++k) lp[k] = ls[k].p;
        store((char*) os.p, out->dtype, eval_program(prog, lp));
        // increment_stepper<row_major> (xiterator.hpp:589-631)
        int d = nd;
        while (d != 0) {
            --d;
            if (idx[d] != out->shape[d] - 1) {
                ++idx[d];
                for (int k = 0; k < prog->n_leaves; ++k) ls[k].step(d);
                os.step(d);
                break;
            }
            idx[d] = 0;
            if (d != 0) {
                for (int k = 0; k < prog->n_leaves; ++k) ls[k].reset(d);
                os.reset(d);
            }
        }
    }
    return XTB_OK;
}

static Val identity_of(int op, int rt) {
    // nanmin / nanmax start from NaN (XTENSOR_REDUCER_FUNCTION(nanmin, detail::nan_min, ..., std::nan("0")), xmath.hpp:2427-2442)
    if (op == XTB_RED_NANMIN || op == XTB_RED_NANMAX) {
        if (rt == XTB_F32) return mk(std::numeric_limits<float>::quiet_NaN());
        if (rt == XTB_F64) return mk(std::numeric_limits<double>::quiet_NaN());
        op = (op == XTB_RED_NANMIN) ? XTB_RED_MIN : XTB_RED_MAX;
    }
    switch (op) {
        case XTB_RED_SUM: return cast_to(mk((int32_t) 0), rt);
        case XTB_RED_PROD: return cast_to(mk((int32_t) 1), rt);
        case XTB_RED_MAX:
            switch (rt) {
                case XTB_I32: return mk(std::numeric_limits<int32_t>::lowest());
                case XTB_U32: return mk(std::numeric_limits<uint32_t>::lowest());
                case XTB_I64: return mk(std::numeric_limits<int64_t>::lowest());
                case XTB_U64: return mk(std::numeric_limits<uint64_t>::lowest());
                case XTB_F32: return mk(std::numeric_limits<float>::lowest());
                default: return mk(std::numeric_limits<double>::lowest());
            }
        default:
            switch (rt) {
                case XTB_I32: return mk(std::numeric_limits<int32_t>::max());
                case XTB_U32: return mk(std::numeric_limits<uint32_t>::max());
                case XTB_I64: return mk(std::numeric_limits<int64_t>::max());
                case XTB_U64: return mk(std::numeric_limits<uint64_t>::max());
                case XTB_F32: return mk(std::numeric_limits<float>::max());
                default: return mk(std::numeric_limits<double>::max());
            }
    }
}
static int binop_of(int op) {
    switch (op) {
        case XTB_RED_SUM: return XTB_OP_ADD;
        case XTB_RED_PROD: return XTB_OP_MUL;
        case XTB_RED_MAX: return XTB_OP_MAXIMUM;
        case XTB_RED_NANMIN: return XTB_OP_NANMIN;
        case XTB_RED_NANMAX: return XTB_OP_NANMAX;
        default: return XTB_OP_MINIMUM;
    }
}

struct LazyCtx {
    const xtb_program* prog;
    int n_leaves, ndim, n_axes, binop, acc;
    const int32_t* axes;
    const int64_t* shape;
    Stepper ls[XTB_MAX_LEAVES];
    Val init;
};
// reduce(res, x): x is converted to the accumulator type by usual arithmetic
// conversion (acc type = decltype(reduce(init, x)), xreducer.hpp:292-298)
static Val red(const LazyCtx& c, const Val& a, const Val& x) { return binary(c.binop, a, cast_to(x, c.acc)); }
static Val deref(LazyCtx& c) {
    const char* lp[XTB_MAX_LEAVES];
    for (int k = 0; k < c.n_leaves; ++k) lp[k] = c.ls[k].p;
    return eval_program(c.prog, lp);
}
// xreducer_stepper::aggregate_impl(dim, false_type)  xreducer.hpp:1802-1828
static Val aggregate(LazyCtx& c, int dim) {
    Val res;
    const int index = c.axes[dim];
    const int64_t size = c.shape[index];
    if (dim != c.n_axes - 1) {
        res = aggregate(c, dim + 1);
        for (int64_t i = 1; i != size; ++i) {
            for (int k = 0; k < c.n_leaves; ++k) c.ls[k].step(index);
            res = binary(c.binop, res, aggregate(c, dim + 1));  // merge
        }
    } else {
        res = red(c, c.init, deref(c));
        for (int64_t i = 1; i != size; ++i) {
            for (int k = 0; k < c.n_leaves; ++k) c.ls[k].step(index);
            res = red(c, res, deref(c));
        }
    }
    for (int k = 0; k < c.n_leaves; ++k) c.ls[k].reset(index);
    return res;
}

// mode 0: lazy xreducer evaluated by stepper_assigner; mode 1: reduce_immediate
int xto_reduce(int op, int acc_type, const xtb_program* prog, const xtb_operand* leaves, int ndim, const int64_t* shape,
               int n_axes, const int32_t* axes, int keep_dims, const void* initial, const xtb_operand* out, int mode) {
    // xreducer.hpp:336-350
    for (int i = 1; i < n_axes; ++i) {
        if (axes[i] < axes[i - 1]) return fail(XTB_ERR_AXIS, "Reducing axes should be sorted.");
        if (axes[i] == axes[i - 1]) return fail(XTB_ERR_AXIS, "Reducing axes should not contain duplicates.");
    }
    if (n_axes > 0 && (axes[0] < 0 || axes[n_axes - 1] > ndim - 1)) return fail(XTB_ERR_AXIS, "Axis out of bounds for reduction.");
    bool red_axis[XTB_MAX_DIM] = {false};
    for (int i = 0; i < n_axes; ++i) red_axis[axes[i]] = true;
    int64_t total = 1, K = 1;
    for (int d = 0; d < ndim; ++d) {
        total *= shape[d];
        if (!red_axis[d]) K *= shape[d];
    }
    for (int k = 0; k < prog->n_leaves; ++k) {
        int r = check_broadcast(&leaves[k], ndim, shape);
        if (r) return r;
    }
    const int binop = binop_of(op);
    Val init = identity_of(op, acc_type);
    Val initial_v;
    if (initial) {
        uint64_t bits = 0;
        memcpy(&bits, initial, dsize(acc_type));
        initial_v = imm(bits, acc_type);
    }
    const int osz = dsize(out->dtype);
    char* obase = (char*) out->base + out->offset * osz;
    // out strides over the iteration dims
    int64_t ostride[XTB_MAX_DIM] = {0};
    for (int d = 0, od = 0; d < ndim; ++d) {
        if (red_axis[d]) {
            if (keep_dims) ++od;
        } else {
            ostride[d] = out->shape[od] == 1 ? 0 : out->stride[od];
            ++od;
        }
    }
    if (K == 0) return XTB_OK;

    if (mode == 0 || n_axes == 0 || total == 0) {
        LazyCtx c;
        c.prog = prog; c.n_leaves = prog->n_leaves; c.ndim = ndim; c.n_axes = n_axes; c.binop = binop; c.acc = acc_type;
        c.axes = axes; c.shape = shape; c.init = init;
        for (int k = 0; k < prog->n_leaves; ++k) c.ls[k].init(&leaves[k], ndim, shape);
        // iterate output positions in row-major order over kept dims (stepper_assigner over the reducer)
        int64_t idx[XTB_MAX_DIM] = {0};
        for (int64_t o = 0; o < K; ++o) {
            int64_t ooff = 0;
            for (int d = 0; d < ndim; ++d) ooff += idx[d] * ostride[d];
            Val res;
            if (total == 0) res = initial ? initial_v : init;                          // xreducer.hpp:1782-1785
            else if (n_axes == 0) res = red(c, initial ? initial_v : init, deref(c));  // :1786-1789
            else {
                res = aggregate(c, 0);
                if (initial) res = binary(binop, initial_v, res);                      // :1793-1796
            }
            store(obase + ooff * osz, out->dtype, res);
            // advance kept coordinates
            int d = ndim;
            while (d != 0) {
                --d;
                if (red_axis[d]) continue;
                if (idx[d] != shape[d] - 1) {
                    ++idx[d];
                    for (int k = 0; k < prog->n_leaves; ++k) c.ls[k].step(d);
                    break;
                }
                idx[d] = 0;
                for (int k = 0; k < prog->n_leaves; ++k) c.ls[k].reset(d);
            }
        }
        return XTB_OK;
    }

    // ---- mode 1: reduce_immediate on eval(e) (dense row-major temporary) ----
    // evaluate the expression into a dense row-major buffer of its value type
    Val probe;
    std::vector<Val> e((size_t) total);
    {
        Stepper ls[XTB_MAX_LEAVES];
        for (int k = 0; k < prog->n_leaves; ++k) ls[k].init(&leaves[k], ndim, shape);
        int64_t idx[XTB_MAX_DIM] = {0};
        const char* lp[XTB_MAX_LEAVES];
        for (int64_t i = 0; i < total; ++i) {
            for (int k = 0; k < prog->n_leaves; ++k) lp[k] = ls[k].p;
            e[(size_t) i] = eval_program(prog, lp);
            int d = ndim;
            while (d != 0) {
                --d;
                if (idx[d] != shape[d] - 1) {
                    ++idx[d];
                    for (int k = 0; k < prog->n_leaves; ++k) ls[k].step(d);
                    break;
                }
                idx[d] = 0;
                if (d != 0) for (int k = 0; k < prog->n_leaves; ++k) ls[k].reset(d);
            }
        }
    }
    (void) probe;
    auto reduce_fct = [&](const Val& a, const Val& x) { return binary(binop, a, cast_to(x, acc_type)); };
    // dense result in row-major order of the kept dims
    std::vector<Val> result((size_t) K);
    // e strides (row major, 0 for extent 1: xstrides.hpp:503-530)
    int64_t estride[XTB_MAX_DIM], rstride_dense[XTB_MAX_DIM] = {0};
    {
        int64_t s = 1;
        for (int d = ndim - 1; d >= 0; --d) { estride[d] = shape[d] == 1 ? 0 : s; s *= shape[d]; }
        int64_t rs = 1;
        for (int d = ndim - 1; d >= 0; --d) {
            if (red_axis[d]) continue;
            rstride_dense[d] = shape[d] == 1 ? 0 : rs;
            rs *= shape[d];
        }
    }
    if (n_axes == ndim) {
        // fast track for complete reduction (:354-360): std::accumulate over storage from init (or initial)
        Val tmp = initial ? initial_v : init;
        for (int64_t i = 0; i < total; ++i) tmp = reduce_fct(tmp, e[(size_t) i]);
        result[0] = tmp;
    } else {
        const int leading_ax = axes[n_axes - 1];
        // find the next non-zero stride towards the front (:365-374)
        int sf = leading_ax;
        int64_t inner_stride = estride[sf];
        while (inner_stride == 0 && sf != 0) { --sf; inner_stride = estride[sf]; }
        if (inner_stride == 0) {
            for (int64_t i = 0; i < total; ++i) result[(size_t) i] = reduce_fct(init, e[(size_t) i]);
        } else {
            const int64_t inner_loop_size = inner_stride;
            int64_t outer_loop_size = shape[leading_ax];
            // merge adjacent reduction axes at the end (:386-402)
            int last_ax = axes[n_axes - 1];
            for (int i = n_axes - 2; i >= 0; --i) {
                if (std::abs(axes[i] - last_ax) == 1) { last_ax = axes[i]; outer_loop_size *= shape[last_ax]; }
            }
            // iteration over the dims in front of last_ax; strides into the dense result, 0 for reduced (:404-420)
            std::vector<int64_t> iter_shape, iter_strides;
            for (int d = 0; d < last_ax; ++d) { iter_shape.push_back(shape[d]); iter_strides.push_back(red_axis[d] ? 0 : rstride_dense[d]); }
            std::vector<int64_t> tidx(iter_shape.size(), 0);
            auto next_idx = [&]() -> std::pair<bool, int64_t> {
                size_t i = iter_shape.size();
                for (; i > 0; --i) {
                    if (tidx[i - 1] >= iter_shape[i - 1] - 1) tidx[i - 1] = 0;
                    else { tidx[i - 1]++; break; }
                }
                int64_t off = 0;
                for (size_t j = 0; j < tidx.size(); ++j) off += tidx[j] * iter_strides[j];
                return {i == 0, off};
            };
            size_t begin = 0;
            int64_t out = 0, merge_border = 0;
            bool merge = false;
            std::pair<bool, int64_t> idx_res(false, 0);
            if (inner_stride == 1) {  // :482-511
                while (!idx_res.first) {
                    Val tmp = init;
                    for (int64_t i = 0; i < outer_loop_size; ++i) tmp = reduce_fct(tmp, e[begin + (size_t) i]);
                    result[(size_t) out] = merge ? binary(binop, result[(size_t) out], tmp) : tmp;
                    begin += (size_t) outer_loop_size;
                    idx_res = next_idx();
                    out = idx_res.second;
                    if (out > merge_border) { merge = false; merge_border = out; }
                    else merge = true;
                }
            } else {  // :512-551
                while (!idx_res.first) {
                    for (int64_t j = 0; j < inner_loop_size; ++j)
                        result[(size_t) (out + j)] = merge ? reduce_fct(result[(size_t) (out + j)], e[begin + (size_t) j])
                                                            : reduce_fct(init, e[begin + (size_t) j]);
                    begin += (size_t) inner_stride;
                    for (int64_t i = 1; i < outer_loop_size; ++i) {
                        for (int64_t j = 0; j < inner_loop_size; ++j)
                            result[(size_t) (out + j)] = reduce_fct(result[(size_t) (out + j)], e[begin + (size_t) j]);
                        begin += (size_t) inner_stride;
                    }
                    idx_res = next_idx();
                    out = idx_res.second;
                    if (out > merge_border) { merge = false; merge_border = out; }
                    else merge = true;
                }
            }
        }
        if (initial)  // :552-563
            for (int64_t i = 0; i < K; ++i) result[(size_t) i] = binary(binop, result[(size_t) i], initial_v);
    }
    // scatter the dense result to `out`
    {
        int64_t idx[XTB_MAX_DIM] = {0};
        for (int64_t o = 0; o < K; ++o) {
            int64_t ooff = 0;
            for (int d = 0; d < ndim; ++d) ooff += idx[d] * ostride[d];
            store(obase + ooff * osz, out->dtype, result[(size_t) o]);
            for (int d = ndim - 1; d >= 0; --d) {
                if (red_axis[d]) continue;
                if (++idx[d] < shape[d]) break;
                idx[d] = 0;
            }
        }
    }
    return XTB_OK;
}

// accumulator_impl (xaccumulator.hpp:215-341).  `in` is copied (with conversion
// to the result value type) into dense row-major `out`, then scanned in place:
//   axis >= 0: res[pos + inner_stride] = f(res[pos], res[pos + inner_stride])  (:282-294)
//   axis <  0: flat scan in row-major traversal order into a 1-D result        (:299-341)
int xto_scan(int op, int acc_type, const xtb_operand* in, int axis, const xtb_operand* out) {
    const int nd = in->ndim;
    if (axis >= nd) return fail(XTB_ERR_AXIS, "Axis larger than expression dimension in accumulator.");
    int64_t total = 1;
    for (int d = 0; d < nd; ++d) total *= in->shape[d];
    if (total == 0) return XTB_OK;
    const int binop = (op == XTB_RED_PROD) ? XTB_OP_MUL : XTB_OP_ADD;
    // result = e  (dense row-major, value type = acc_type)
    std::vector<Val> res((size_t) total);
    {
        Stepper s;
        s.init(in, nd, in->shape);
        int64_t idx[XTB_MAX_DIM] = {0};
        for (int64_t i = 0; i < total; ++i) {
            res[(size_t) i] = cast_to(load(s.p, in->dtype), acc_type);
            int d = nd;
            while (d != 0) {
                --d;
                if (idx[d] != in->shape[d] - 1) { ++idx[d]; s.step(d); break; }
                idx[d] = 0;
                if (d != 0) s.reset(d);
            }
        }
    }
    if (axis < 0) {
        for (int64_t i = 0; i + 1 < total; ++i) res[(size_t) i + 1] = binary(binop, res[(size_t) i], res[(size_t) i + 1]);
    } else {
        int64_t inner = 1;
        for (int d = axis + 1; d < nd; ++d) inner *= in->shape[d];
        const int64_t n = in->shape[axis];
        const int64_t outer = total / (inner * n);
        for (int64_t o = 0; o < outer; ++o)
            for (int64_t i = 1; i < n; ++i)
                for (int64_t j = 0; j < inner; ++j) {
                    const size_t pos = (size_t) ((o * n + i) * inner + j);
                    res[pos] = binary(binop, res[pos - (size_t) inner], res[pos]);
                }
    }
    // store into out (dense or strided)
    Stepper so;
    const int ond = out->ndim;
    so.init(out, ond, out->shape);
    int64_t idx[XTB_MAX_DIM] = {0};
    for (int64_t i = 0; i < total; ++i) {
        store((char*) so.p, out->dtype, res[(size_t) i]);
        int d = ond;
        while (d != 0) {
            --d;
            if (idx[d] != out->shape[d] - 1) { ++idx[d]; so.step(d); break; }
            idx[d] = 0;
            if (d != 0) so.reset(d);
        }
    }
    return XTB_OK;
}

// argmin / argmax: detail::arg_func_impl (misc/xsort.hpp:1150-1231) and the flat overloads (:1237-1245, 1267-1275,
// std::min_element / std::max_element).  Per lane along `axis`:
//     val = x[0]; idx = 0;  for i = 1..n-1: if (cmp(x[i], val)) { val = x[i]; idx = i; }
// cmp = std::less (argmin) / std::greater (argmax) on the element type -- ties keep the first index, a NaN never
// replaces the running value, a NaN at x[0] is never replaced.  (std::min_element is the same loop with
// `*it < *smallest`, std::max_element with `*largest < *it`.)
int xto_argreduce(int op, const xtb_operand* in, int axis, const xtb_operand* out) {
    xtb_operand x = *in;
    if (axis < 0) {
        int64_t total = 1;
        for (int d = 0; d < in->ndim; ++d) total *= in->shape[d];
        // flat: row-major traversal; gather through the stepper into a dense copy first
        x.ndim = 1;
        x.shape[0] = total;
        axis = 0;
    } else if (axis >= in->ndim) {
        return fail(XTB_ERR_AXIS, "Axis out of bounds");
    }
    // dense row-major copy of the operand in its element type (what eval(e) gives)
    int64_t total = 1;
    for (int d = 0; d < in->ndim; ++d) total *= in->shape[d];
    if (total == 0) return XTB_OK;
    std::vector<Val> v((size_t) total);
    {
        Stepper s;
        s.init(in, in->ndim, in->shape);
        int64_t idx[XTB_MAX_DIM] = {0};
        for (int64_t i = 0; i < total; ++i) {
            v[(size_t) i] = load(s.p, in->dtype);
            int d = in->ndim;
            while (d != 0) {
                --d;
                if (idx[d] != in->shape[d] - 1) { ++idx[d]; s.step(d); break; }
                idx[d] = 0;
                if (d != 0) s.reset(d);
            }
        }
    }
    int64_t inner = 1;
    for (int d = axis + 1; d < x.ndim; ++d) inner *= x.shape[d];
    const int64_t n = x.shape[axis];
    const int64_t outer = total / (inner * n);
    const int cmp = op == XTB_RED_MIN ? XTB_OP_LT : XTB_OP_GT;
    std::vector<uint64_t> res((size_t) (outer * inner));
    for (int64_t o = 0; o < outer; ++o)
        for (int64_t j = 0; j < inner; ++j) {
            Val best = v[(size_t) (o * n * inner + j)];
            uint64_t bi = 0;
            for (int64_t i = 1; i < n; ++i) {
                const Val& c = v[(size_t) ((o * n + i) * inner + j)];
                if (binary(cmp, c, best).i32) { best = c; bi = (uint64_t) i; }
            }
            res[(size_t) (o * inner + j)] = bi;
        }
    Stepper so;
    so.init(out, out->ndim, out->shape);
    int64_t idx[XTB_MAX_DIM] = {0};
    for (size_t i = 0; i < res.size(); ++i) {
        store((char*) so.p, out->dtype, mk((uint64_t) res[i]));
        int d = out->ndim;
        while (d != 0) {
            --d;
            if (idx[d] != out->shape[d] - 1) { ++idx[d]; so.step(d); break; }
            idx[d] = 0;
            if (d != 0) so.reset(d);
        }
    }
    return XTB_OK;
}

}  // extern "C"
