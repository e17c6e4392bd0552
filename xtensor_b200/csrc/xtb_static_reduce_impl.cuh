// xtb_static_reduce_impl.cuh -- compile-time instantiations of the reduction kernels for
// the common (expression, reducer, accumulator) combinations: plain sum / prod /
// amax / amin of f32, f64, i32 containers and sum(square(a - m)) (the second pass
// of xt::variance, include/xtensor/core/xmath.hpp:2082-2105).
// Included by two translation units: xtb_static_reduce.cu (128-bit vectors: V = 4 / 2) and
// xtb_static_reduce_v1.cu (V = 1: rows whose pitch or base is not a multiple of 16 bytes -- odd extents,
// offset views -- are read with coalesced scalar loads by the same kernels instead of falling to the interpreter).
// XTB_SR_SCALAR (0 / 1) and XTB_SR_TABLE (the exported table function) are set by the including file.
#include <utility>
#include "xtb_reduce.cuh"

namespace xtb {
namespace {

struct Combo {
    sprogs::SP prog;
    int binop, acc_rt;
};
struct Tbl {
    static constexpr sprogs::SP progs[] = {sprogs::copy_f32, sprogs::copy_f64, sprogs::copy_i32,
                                           sprogs::sq_sub_f32, sprogs::sq_sub_f64, sprogs::square_f32,
                                           sprogs::square_f64};
};
static const char* const kProgNames[] = {"copy_f32", "copy_f64", "copy_i32", "sq_sub_f32", "sq_sub_f64",
                                         "square_f32", "square_f64"};
struct ComboId {
    int prog, binop, acc_rt;
};
constexpr ComboId kCombos[] = {
    {0, XTB_OP_ADD, XTB_F32}, {0, XTB_OP_MUL, XTB_F32}, {0, XTB_OP_MAXIMUM, XTB_F32}, {0, XTB_OP_MINIMUM, XTB_F32},
    {1, XTB_OP_ADD, XTB_F64}, {1, XTB_OP_MUL, XTB_F64}, {1, XTB_OP_MAXIMUM, XTB_F64}, {1, XTB_OP_MINIMUM, XTB_F64},
    {2, XTB_OP_ADD, XTB_I32}, {2, XTB_OP_MUL, XTB_I32}, {2, XTB_OP_MAXIMUM, XTB_I32}, {2, XTB_OP_MINIMUM, XTB_I32},
    {3, XTB_OP_ADD, XTB_F32}, {4, XTB_OP_ADD, XTB_F64}, {5, XTB_OP_ADD, XTB_F32}, {6, XTB_OP_ADD, XTB_F64},
};
constexpr int kCount = (int) (sizeof(kCombos) / sizeof(kCombos[0]));

template <int C> int launch_one(const RdParams& p, DeviceCtx* ctx, bool inner) {
    constexpr ComboId c = kCombos[C];
    constexpr bool k64 = sprogs::is64(Tbl::progs[c.prog]);
    using Slot = std::conditional_t<k64, uint64_t, uint32_t>;
    return launch_reduce<StaticEval<Tbl, c.prog>, StaticAcc<c.binop, c.acc_rt>, Slot, (XTB_SR_SCALAR ? 1 : (k64 ? 2 : 4))>(
        p, ctx, inner, kProgNames[c.prog]);
}

StaticReduceEntry g_entries[kCount];
template <int... I> void fill(std::integer_sequence<int, I...>) {
    ((g_entries[I] = StaticReduceEntry{&Tbl::progs[kCombos[I].prog], kCombos[I].binop, kCombos[I].acc_rt,
                                       kProgNames[kCombos[I].prog], &launch_one<I>}),
     ...);
}

}  // namespace

StaticReduceTable XTB_SR_TABLE() {
    static const bool once = (fill(std::make_integer_sequence<int, kCount>{}), true);
    (void) once;
    return StaticReduceTable{g_entries, kCount};
}

}  // namespace xtb
