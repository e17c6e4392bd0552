// xtb_static_f32.cu -- compile-time instantiations of the f32 expression
// programs listed in xtb_static_programs.cuh (fully unrolled elementwise kernels).
#include <utility>
#include "xtb_ew_tma.cuh"

namespace xtb {
namespace {

#define XTB_SP_ADDR(P) sprogs::P,
#define XTB_SP_NAME(P) #P,
struct Tbl {
    static constexpr sprogs::SP progs[] = { XTB_STATIC_LIST_F32(XTB_SP_ADDR) };
};
static const char* const kNames[] = { XTB_STATIC_LIST_F32(XTB_SP_NAME) };
constexpr int kCount = (int) (sizeof(Tbl::progs) / sizeof(Tbl::progs[0]));

using Slot = uint32_t;
constexpr int kV = 4;

template <int ID> int launch_one(const EwParams& p, DeviceCtx* ctx) {
    return launch_ew_nd<StaticEval<Tbl, ID>, Slot, kV>(p, ctx, kNames[ID]);
}

template <int ID> int launch_tile_one(const EwParams& p, DeviceCtx* ctx) {
    // every operand through TMA when the expression is regular (rank 2, aligned, one element size); else the
    // register-staged tile kernel
    const int r = launch_ew_tile_tma<StaticEval<Tbl, ID>, Slot>(p, ctx, kNames[ID]);
    if (r != 1) return r;
    return launch_ew_tile<StaticEval<Tbl, ID>, Slot>(p, ctx, kNames[ID]);
}

StaticEntry g_entries[kCount];
template <int... I> void fill(std::integer_sequence<int, I...>) {
    ((g_entries[I] = StaticEntry{&Tbl::progs[I], kNames[I], &launch_one<I>, &launch_tile_one<I>}), ...);
}

}  // namespace

StaticTable static_table_f32() {
    static const bool once = (fill(std::make_integer_sequence<int, kCount>{}), true);
    (void) once;
    return StaticTable{g_entries, kCount};
}

}  // namespace xtb
