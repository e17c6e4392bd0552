// xtb_arg.cu -- xtb_argreduce: index of the first minimum / maximum along an axis.
//
// Replaces xt::argmin / xt::argmax and detail::arg_func_impl
// (include/xtensor/misc/xsort.hpp:1150-1300).  The reference walks every lane along `axis` sequentially:
//     val = x[0]; idx = 0;  for i in 1..n-1:  if (cmp(x[i], val)) { val = x[i]; idx = i; }
// with cmp = std::less (argmin) / std::greater (argmax).  Consequences the device path reproduces:
//   * ties keep the FIRST index;
//   * a NaN never replaces the running value (cmp is false), and if x[0] is NaN nothing ever replaces it:
//     the result is 0 when x[0] is NaN, otherwise the first extreme among the non-NaN elements.
// Device formulations (order independent, so they run on the split / merged reduction kernels of xtb_reduce.cuh;
// `i` is a small index vector broadcast along the other dims):
//   * element types of <= 32 bits, ONE pass: a MIN reduction over packed 64-bit keys
//         key = (isnan(x) && i == 0) ? 0 : (ordkey(x) << 32 | i)
//     where ordkey is a monotone 32-bit image of x (of -x for argmax; -0.0 == +0.0; NaN -> 0xffffffff, XTB_OP_ORDKEY):
//     the smallest key is the extreme value with the smallest index; the index is the key's low half;
//   * 64-bit element types, two passes:
//         m   = nan_min / nan_max over the axis            (the extreme of the non-NaN elements; keep_dims)
//         idx = min over the axis of  ((x == m) || (isnan(x) && i == 0)) ? i : SIZE_MAX
#include <vector>
#include "xtb_common.hpp"

namespace xtb {

__global__ void __launch_bounds__(256) k_iota_u64(unsigned long long* p, unsigned long long n) {
    const unsigned long long i = (unsigned long long) blockIdx.x * 256 + threadIdx.x;
    if (i < n) p[i] = i;
}

static int regtype_of_dt(int dt) { return dt < XTB_I32 ? (int) XTB_I32 : dt; }

struct PoolBuf {
    void* p = nullptr;
    ~PoolBuf() { if (p) xtb_free(p); }
    int alloc(size_t bytes) { return xtb_malloc(bytes ? bytes : 1, &p); }
};

static void emit(xtb_program& pr, int op, int type, int src = 0, int arg = 0) {
    pr.insns[pr.n_insns++] = xtb_insn{(uint8_t) op, (uint8_t) type, (uint8_t) src, (uint8_t) arg};
}

}  // namespace xtb

using namespace xtb;

extern "C" int xtb_argreduce(int op, const xtb_operand* in, int axis, const xtb_operand* out) {
    if (!in || !out) XTB_FAIL(XTB_ERR_INVALID, "null argument");
    if (op != XTB_RED_MIN && op != XTB_RED_MAX) XTB_FAIL(XTB_ERR_INVALID, "argreduce supports XTB_RED_MIN / XTB_RED_MAX");
    if (in->ndim < 0 || in->ndim > XTB_MAX_DIM) XTB_FAIL(XTB_ERR_INVALID, "rank %d out of range", in->ndim);
    if (in->dtype < 0 || in->dtype >= XTB_DTYPE_COUNT) XTB_FAIL(XTB_ERR_INVALID, "bad dtype");
    if (out->dtype != XTB_U64 && out->dtype != XTB_I64) XTB_FAIL(XTB_ERR_INVALID, "argreduce writes std::size_t (u64 / i64) indices");
    DeviceCtx* ctx;
    XTB_TRY(get_ctx(&ctx));
    XTB_LAUNCH_LOCK(ctx);

    // iteration space: the operand's own shape, or -- flat -- its row-major traversal as one dim
    xtb_operand x = *in;
    if (axis < 0) {
        // the flattened traversal must be affine: dense row-major storage (the header evaluates anything else first,
        // as the reference does with eval(e), xsort.hpp:1240-1241)
        int64_t expect = 1, total = 1;
        bool dense = true;
        for (int d = in->ndim - 1; d >= 0; --d) {
            if (in->shape[d] != 1 && in->stride[d] != expect) dense = false;
            expect *= in->shape[d];
            total *= in->shape[d];
        }
        if (!dense) XTB_FAIL(XTB_ERR_UNSUPPORTED, "flat argmin / argmax needs a dense row-major operand");
        x.ndim = 1;
        x.shape[0] = total;
        x.stride[0] = 1;
        axis = 0;
        if (out->ndim != 0) XTB_FAIL(XTB_ERR_SHAPE, "flat argmin / argmax writes a 0-d result");
    } else {
        if (axis >= in->ndim) XTB_FAIL(XTB_ERR_AXIS, "Axis %d out of bounds for reduction.", axis);
        if (out->ndim != in->ndim - 1) XTB_FAIL(XTB_ERR_SHAPE, "argmin / argmax output has rank %d, expected %d", out->ndim, in->ndim - 1);
    }
    const int nd = x.ndim;
    const int64_t n = x.shape[axis];
    int64_t K = 1;
    for (int d = 0; d < nd; ++d)
        if (d != axis) K *= x.shape[d];
    if (K == 0) return XTB_OK;
    if (n == 0) XTB_FAIL(XTB_ERR_INVALID, "argmin / argmax of an empty axis");
    const int rt = regtype_of_dt(x.dtype);
    const int rsz = (rt == XTB_I64 || rt == XTB_U64 || rt == XTB_F64) ? 8 : 4;

    // m (keep_dims, register-typed, dense) and the index vector
    PoolBuf mbuf, ibuf;
    XTB_TRY(mbuf.alloc((size_t) K * rsz));
    XTB_TRY(ibuf.alloc((size_t) n * 8));
    k_iota_u64<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->stream>>>((unsigned long long*) ibuf.p, (unsigned long long) n);
    note_launch("k_iota_u64");
    XTB_TRY(check_launch("k_iota_u64"));

    xtb_operand m{};
    m.base = mbuf.p;
    m.dtype = rt;
    m.ndim = nd;
    {
        int64_t st = 1;
        for (int d = nd - 1; d >= 0; --d) {
            m.shape[d] = d == axis ? 1 : x.shape[d];
            m.stride[d] = (d == axis || x.shape[d] == 1) ? 0 : st;
            if (d != axis) st *= x.shape[d];
        }
    }
    xtb_operand iota{};
    iota.base = ibuf.p;
    iota.dtype = XTB_U64;
    iota.ndim = nd;
    for (int d = 0; d < nd; ++d) {
        iota.shape[d] = d == axis ? n : 1;
        iota.stride[d] = (d == axis && n > 1) ? 1 : 0;
    }
    int64_t shape[XTB_MAX_DIM] = {0};
    for (int d = 0; d < nd; ++d) shape[d] = x.shape[d];
    const int32_t ax = axis;

    const bool fp = rt == XTB_F32 || rt == XTB_F64;
    xtb_operand o = *out;
    o.dtype = XTB_U64;   // i64 storage holds the same bits for every valid index
    if (rsz == 4 && n < (1ll << 32) && !options().arg_two_pass) {
        // ---- one pass: MIN over packed (order key, index) ----
        xtb_program pk{};
        xtb_operand leaves[2] = {x, iota};
        pk.n_leaves = 2;
        pk.n_imms = 1;
        pk.imms[0] = 0;
        if (fp) {
            emit(pk, XTB_OP_PUSH, x.dtype, XTB_SRC_LEAF, 0);
            emit(pk, XTB_OP_ISNAN, rt);                                 // isnan(x)
            emit(pk, XTB_OP_PUSH, XTB_U64, XTB_SRC_LEAF, 1);
            emit(pk, XTB_OP_EQ, XTB_U64, XTB_SRC_IMM, 0);               // i == 0
            emit(pk, XTB_OP_LAND, XTB_I32, XTB_SRC_STACK, 0);           // a NaN in front is never replaced: the best key
            emit(pk, XTB_OP_PUSH, XTB_U64, XTB_SRC_IMM, 0);
        }
        emit(pk, XTB_OP_PUSH, x.dtype, XTB_SRC_LEAF, 0);
        emit(pk, XTB_OP_ORDKEY, rt, 0, op == XTB_RED_MAX ? 1 : 0);
        emit(pk, XTB_OP_BOR, XTB_U64, XTB_SRC_LEAF, 1);                 // | i
        if (fp) emit(pk, XTB_OP_WHERE, XTB_U64);
        XTB_TRY(xtb_reduce(XTB_RED_MIN, XTB_U64, &pk, leaves, nd, shape, 1, &ax, 0, nullptr, &o, 0));
        // the index is the low half of the winning key
        xtb_program pd{};
        pd.n_leaves = 1;
        pd.n_imms = 1;
        pd.imms[0] = 0xffffffffull;
        emit(pd, XTB_OP_PUSH, XTB_U64, XTB_SRC_LEAF, 0);
        emit(pd, XTB_OP_BAND, XTB_U64, XTB_SRC_IMM, 0);
        return xtb_assign(&pd, &o, &o);
    }
    // pass 1: m = nan_min / nan_max(x) along the axis
    {
        xtb_program p1{};
        p1.n_leaves = 1;
        emit(p1, XTB_OP_PUSH, x.dtype, XTB_SRC_LEAF, 0);
        XTB_TRY(xtb_reduce(op == XTB_RED_MIN ? XTB_RED_NANMIN : XTB_RED_NANMAX, rt, &p1, &x, nd, shape, 1, &ax, 1, nullptr, &m, 0));
    }
    // pass 2: first index whose element equals m (or index 0 when x[0] is NaN)
    {
        xtb_program p2{};
        xtb_operand leaves[3] = {x, m, iota};
        p2.n_leaves = 3;
        p2.n_imms = 2;
        p2.imms[0] = ~0ull;   // SIZE_MAX: "not a candidate"
        p2.imms[1] = 0;
        emit(p2, XTB_OP_PUSH, x.dtype, XTB_SRC_LEAF, 0);
        emit(p2, XTB_OP_EQ, rt, XTB_SRC_LEAF, 1);                       // x == m
        if (fp) {
            emit(p2, XTB_OP_PUSH, x.dtype, XTB_SRC_LEAF, 0);
            emit(p2, XTB_OP_ISNAN, rt);                                 // isnan(x)
            emit(p2, XTB_OP_PUSH, XTB_U64, XTB_SRC_LEAF, 2);
            emit(p2, XTB_OP_EQ, XTB_U64, XTB_SRC_IMM, 1);               // i == 0
            emit(p2, XTB_OP_LAND, XTB_I32, XTB_SRC_STACK, 0);
            emit(p2, XTB_OP_LOR, XTB_I32, XTB_SRC_STACK, 0);
        }
        emit(p2, XTB_OP_PUSH, XTB_U64, XTB_SRC_LEAF, 2);
        emit(p2, XTB_OP_PUSH, XTB_U64, XTB_SRC_IMM, 0);
        emit(p2, XTB_OP_WHERE, XTB_U64);
        XTB_TRY(xtb_reduce(XTB_RED_MIN, XTB_U64, &p2, leaves, nd, shape, 1, &ax, 0, nullptr, &o, 0));
    }
    return XTB_OK;
}
