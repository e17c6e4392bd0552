// xtb_static_programs.cuh -- expression programs that are instantiated at
// compile time (fully unrolled evaluation, all leaf loads issued up front).
// The instruction encoding is exactly what include/xtb200/lower.hpp emits for
// the named xtensor expressions; any other program runs on the interpreter.
#pragma once
#include "xtb_ops.cuh"

namespace xtb {
namespace sprogs {

constexpr int kCap = XTB_MAX_INSNS;
using SP = SProg<kCap>;

constexpr SInsn push_leaf(int k, int dt) { return {XTB_OP_PUSH, dt, XTB_SRC_LEAF, k}; }
constexpr SInsn push_imm(int k, int rt) { return {XTB_OP_PUSH, rt, XTB_SRC_IMM, k}; }
constexpr SInsn un(int op, int rt, int arg = 0) { return {op, rt, 0, arg}; }
constexpr SInsn bin(int op, int rt) { return {op, rt, XTB_SRC_STACK, 0}; }
constexpr SInsn bin_leaf(int op, int rt, int k, bool rev = false) {
    return {op, rt, XTB_SRC_LEAF | (rev ? XTB_SRC_REV : 0), k};
}
constexpr SInsn bin_imm(int op, int rt, int k, bool rev = false) {
    return {op, rt, XTB_SRC_IMM | (rev ? XTB_SRC_REV : 0), k};
}

template <int N> constexpr SP make(const SInsn (&ins)[N], int n_leaves, int n_imms) {
    static_assert(N <= kCap, "static program too long");
    SP p{};
    p.n = N;
    for (int i = 0; i < N; ++i) p.ins[i] = ins[i];
    p.n_leaves = n_leaves;
    p.n_imms = n_imms;
    return p;
}

constexpr bool is64(const SP& p) {
    for (int i = 0; i < p.n; ++i) {
        const int t = p.ins[i].type;
        if (t == XTB_I64 || t == XTB_U64 || t == XTB_F64) return true;
    }
    return false;
}

// register type of the value a program leaves on the stack
constexpr int result_type(const SP& p) {
    int st[XTB_MAX_STACK + 1] = {0};
    int n = 0;
    for (int i = 0; i < p.n; ++i) {
        const SInsn in = p.ins[i];
        if (in.op == XTB_OP_PUSH) st[n++] = (in.src == XTB_SRC_LEAF) ? regtype_of(in.type) : in.type;
        else if (in.op < XTB_OP_ADD) {
            if (in.op == XTB_OP_CAST) st[n - 1] = regtype_of(in.arg);
            else if (in.op == XTB_OP_ORDKEY) st[n - 1] = XTB_U64;
            else if (is_pred_op(in.op)) st[n - 1] = XTB_I32;
        } else if (in.op < XTB_OP_WHERE) {
            if ((in.src & 3) == XTB_SRC_STACK) --n;
            st[n - 1] = is_cmp_op(in.op) ? (int) XTB_I32 : in.type;
        } else {
            n -= 2;
            st[n - 1] = in.type;
        }
    }
    return st[0];
}

// storage dtype of leaf k, read off the instruction that references it
constexpr int leaf_dtype(const SP& p, int k) {
    for (int i = 0; i < p.n; ++i) {
        const SInsn in = p.ins[i];
        const bool push_leaf = in.op == XTB_OP_PUSH && in.src == XTB_SRC_LEAF;
        const bool fused_leaf = in.op >= XTB_OP_ADD && in.op < XTB_OP_WHERE && (in.src & 3) == XTB_SRC_LEAF;
        if ((push_leaf || fused_leaf) && in.arg == k) return in.type;
    }
    return XTB_F32;
}

#define XTB_SP_TYPED(NAME, T)                                                                      \
    /* out = a */                                                                                   \
    inline constexpr SP copy_##NAME = make({push_leaf(0, T)}, 1, 0);                                \
    /* a + b, a - b, a * b, a / b  (xoperation.hpp:231-330) */                                      \
    inline constexpr SP add_##NAME = make({push_leaf(0, T), bin_leaf(XTB_OP_ADD, T, 1)}, 2, 0);     \
    inline constexpr SP sub_##NAME = make({push_leaf(0, T), bin_leaf(XTB_OP_SUB, T, 1)}, 2, 0);     \
    inline constexpr SP mul_##NAME = make({push_leaf(0, T), bin_leaf(XTB_OP_MUL, T, 1)}, 2, 0);     \
    inline constexpr SP div_##NAME = make({push_leaf(0, T), bin_leaf(XTB_OP_DIV, T, 1)}, 2, 0);     \
    /* s * a, a * s, a + s */                                                                       \
    inline constexpr SP smul_##NAME = make({push_imm(0, T), bin_leaf(XTB_OP_MUL, T, 0)}, 1, 1);     \
    inline constexpr SP muls_##NAME = make({push_leaf(0, T), bin_imm(XTB_OP_MUL, T, 0)}, 1, 1);     \
    inline constexpr SP adds_##NAME = make({push_leaf(0, T), bin_imm(XTB_OP_ADD, T, 0)}, 1, 1);     \
    /* s0 * x - s1 * y  (benchmark/benchmark_assign.cpp:80-90) and s0 * x + s1 * y */               \
    inline constexpr SP axmby_##NAME = make({push_imm(0, T), bin_leaf(XTB_OP_MUL, T, 0),            \
                                             push_imm(1, T), bin_leaf(XTB_OP_MUL, T, 1),            \
                                             bin(XTB_OP_SUB, T)}, 2, 2);                            \
    inline constexpr SP axpby_##NAME = make({push_imm(0, T), bin_leaf(XTB_OP_MUL, T, 0),            \
                                             push_imm(1, T), bin_leaf(XTB_OP_MUL, T, 1),            \
                                             bin(XTB_OP_ADD, T)}, 2, 2);                            \
    /* sin(a) * b + s * d  (BASELINE cfg2) */                                                       \
    inline constexpr SP sinmul_axpy_##NAME = make({push_leaf(0, T), un(XTB_OP_SIN, T),              \
                                                   bin_leaf(XTB_OP_MUL, T, 1), push_imm(0, T),      \
                                                   bin_leaf(XTB_OP_MUL, T, 2), bin(XTB_OP_ADD, T)}, 3, 1); \
    /* exp(a - m), square(a - m)  (BASELINE cfg5: map and variance second pass) */                  \
    inline constexpr SP exp_sub_##NAME = make({push_leaf(0, T), bin_leaf(XTB_OP_SUB, T, 1),         \
                                               un(XTB_OP_EXP, T)}, 2, 0);                           \
    inline constexpr SP sq_sub_##NAME = make({push_leaf(0, T), bin_leaf(XTB_OP_SUB, T, 1),          \
                                              un(XTB_OP_SQUARE, T)}, 2, 0);                         \
    /* unary maps */                                                                                \
    inline constexpr SP exp_##NAME = make({push_leaf(0, T), un(XTB_OP_EXP, T)}, 1, 0);              \
    inline constexpr SP sin_##NAME = make({push_leaf(0, T), un(XTB_OP_SIN, T)}, 1, 0);              \
    inline constexpr SP square_##NAME = make({push_leaf(0, T), un(XTB_OP_SQUARE, T)}, 1, 0);

XTB_SP_TYPED(f32, XTB_F32)
XTB_SP_TYPED(f64, XTB_F64)
XTB_SP_TYPED(i32, XTB_I32)

}  // namespace sprogs
}  // namespace xtb

#define XTB_FOR_EACH_TYPED(M, NAME) \
    M(copy_##NAME) M(add_##NAME) M(sub_##NAME) M(mul_##NAME) M(div_##NAME) M(smul_##NAME) M(muls_##NAME) \
    M(adds_##NAME) M(axmby_##NAME) M(axpby_##NAME)
#define XTB_FOR_EACH_FLOAT(M, NAME) \
    M(sinmul_axpy_##NAME) M(exp_sub_##NAME) M(sq_sub_##NAME) M(exp_##NAME) M(sin_##NAME) M(square_##NAME)

#define XTB_STATIC_LIST_F32(M) XTB_FOR_EACH_TYPED(M, f32) XTB_FOR_EACH_FLOAT(M, f32)
#define XTB_STATIC_LIST_F64(M) XTB_FOR_EACH_TYPED(M, f64) XTB_FOR_EACH_FLOAT(M, f64)
#define XTB_STATIC_LIST_I32(M) XTB_FOR_EACH_TYPED(M, i32)
