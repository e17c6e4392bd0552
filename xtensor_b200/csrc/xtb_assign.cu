// xtb_assign.cu -- xtb_assign: evaluate a lowered xfunction tree into a container.
// Host side: validation, broadcast alignment (xt::broadcast_shape,
// include/xtensor/core/xstrides.hpp:737-780), iteration-space collapsing, kernel
// selection.  Kernels: xtb_ew.cuh.
#include <cstdlib>
#include "xtb_ew.cuh"
#include "xtb_jit.hpp"

namespace xtb {

template <class S, int V> static int launch_interp(const EwParams& p, DeviceCtx* ctx) {
    return launch_ew_generic<InterpEval, S, V>(p, ctx, "interp");
}

static bool ew_nd_ok_rank(int nd) { return nd <= 3; }

static int dispatch_tile(const xtb_program* prog, const EwParams& p, DeviceCtx* ctx, bool w64) {
    const bool no_static = options().no_static != 0;
    if (!no_static) {
        const StaticEntry* e = find_static(prog);
        if (e && sprogs::is64(*e->prog) == w64) return e->launch_tile(p, ctx);
    }
    int64_t tile_elems = 1;
    for (int d = 0; d < p.ndim; ++d) tile_elems *= p.shape[d];
    if (!no_static && p.idx32 && jit_program_ok(prog) && p.out.dtype == (int) p.out_rt && jit_worthwhile(tile_elems)) {
        JitSpec spec;
        spec.kind = JIT_TILE;
        spec.w64 = w64;
        void* fn = nullptr;
        if (jit_get(ctx, prog, spec, &fn) == XTB_OK) {
            int64_t batch = 1;
            for (int d = 0; d < p.ndim - 1; ++d)
                if (d != p.tile_i) batch *= p.shape[d];
            const int64_t blocks = batch * p.ntile_i * p.ntile_j;
            if (blocks < 0x7fffffffLL) {
                if (jit_launch(fn, (unsigned) blocks, 1, 256, 0, ctx->stream, &p) == XTB_OK) {
                    note_launch(w64 ? "k_ew_tile_static<jit,S64>" : "k_ew_tile_static<jit,S32>");
                    return XTB_OK;
                }   // a failed launch falls through to the interpreter kernel
            }
        }
    }
    if (w64) return launch_ew_tile<InterpEval, uint64_t>(p, ctx, "interp");
    return launch_ew_tile<InterpEval, uint32_t>(p, ctx, "interp");
}

static int dispatch_ew(const xtb_program* prog, const EwParams& p, DeviceCtx* ctx, bool w64, int V) {
    const bool no_static = options().no_static != 0;
    if (!no_static && ew_nd_ok(p)) {
        const StaticEntry* e = find_static(prog);
        if (e) {
            const bool k64 = sprogs::is64(*e->prog);
            if (k64 == w64 && V == (k64 ? 2 : 4) && p.out.dtype == sprogs::result_type(*e->prog)) return e->launch_ew(p, ctx);
        }
    }
    // no ahead-of-time instantiation: specialise the same kernel template for this program at run time
    // (V == 1: no operand can be accessed with 128-bit vectors -- broadcast expressions over odd inner extents,
    // offset views; the same kernel with coalesced scalar access instead of the interpreter)
    if (!no_static && ew_nd_ok(p) && (V == (w64 ? 2 : 4) || V == 1) && jit_program_ok(prog) && p.out.dtype == (int) p.out_rt &&
        jit_worthwhile(p.total_vec * V)) {
        JitSpec spec;
        spec.kind = JIT_EW;
        spec.w64 = w64;
        spec.V = V;
        spec.nd = p.ndim;
        void* fn = nullptr;
        if (jit_get(ctx, prog, spec, &fn) == XTB_OK) {
            const int64_t per_block = 256 * kEwItems;
            EwParams q = p;
            bool fast = (p.shape[p.ndim - 1] % V == 0) && (p.total_vec % per_block == 0) && p.out.mode != MODE_GATHER && p.out.mode != MODE_BCAST;
            for (int k = 0; k < p.n_leaves; ++k) fast = fast && p.leaf[k].mode != MODE_GATHER;
            q.fast = fast;
            char name[96];
            snprintf(name, sizeof(name), "k_ew<jit,S%d,V%d,ND%d>%s", w64 ? 64 : 32, V, p.ndim, fast ? "[fast]" : "");
            if (jit_launch(fn, (unsigned) ((p.total_vec + per_block - 1) / per_block), 1, 256, 0, ctx->stream, &q) == XTB_OK) {
                note_launch(name);
                return XTB_OK;
            }   // a failed launch falls through to the interpreter kernel
        }
    }
    if (p.idx32 && p.total_vec < (int64_t) 0x7fffffff && !options().no_staged) {
        // run-time program with 32-bit offsets: staged interpreter (cp.async operand staging)
        EwParams q = p;
        for (int k = 0; k < q.n_leaves; ++k) if (q.leaf[k].mode == MODE_LINEAR) q.leaf[k].mode = MODE_VEC;
        if (q.out.mode == MODE_LINEAR) q.out.mode = MODE_VEC;
        if (w64) return V == 2 ? launch_ew_staged<uint64_t, 2>(q, ctx) : launch_ew_staged<uint64_t, 1>(q, ctx);
        return V == 4 ? launch_ew_staged<uint32_t, 4>(q, ctx) : launch_ew_staged<uint32_t, 1>(q, ctx);
    }
    EwParams q = p;  // the generic kernel does not know MODE_LINEAR
    for (int k = 0; k < q.n_leaves; ++k) if (q.leaf[k].mode == MODE_LINEAR) q.leaf[k].mode = MODE_VEC;
    if (q.out.mode == MODE_LINEAR) q.out.mode = MODE_VEC;
    if (w64) return V == 2 ? launch_interp<uint64_t, 2>(q, ctx) : launch_interp<uint64_t, 1>(q, ctx);
    return V == 4 ? launch_interp<uint32_t, 4>(q, ctx) : launch_interp<uint32_t, 1>(q, ctx);
}

static int launch_space(const xtb_program* prog, const Space& s, const char* const* leaf_ptr, const int* leaf_dtype,
                        char* out_ptr, int out_dtype, int rt, bool w64, DeviceCtx* ctx, int depth) {
    const int OUT = prog->n_leaves;
    // Rank <= 3 kernels use 32-bit offsets: if an operand spans >= 2^31 elements, evaluate the
    // leading dimension in pieces (pointers advance, descriptors stay the same).
    if (s.ndim <= 3 && s.shape[0] >= 2 && depth < 16) {
        bool too_big = s.total >= 0x7fffffffLL;
        for (int k = 0; k <= prog->n_leaves && !too_big; ++k) {
            int64_t span = 0;
            for (int d = 0; d < s.ndim; ++d) span += (s.stride[k][d] < 0 ? -s.stride[k][d] : s.stride[k][d]) * (s.shape[d] - 1);
            too_big = span >= 0x7fffffffLL;
        }
        if (too_big) {
            const int64_t half = s.shape[0] / 2;
            for (int piece = 0; piece < 2; ++piece) {
                Space h = s;
                h.shape[0] = piece == 0 ? half : s.shape[0] - half;
                h.total = s.total / s.shape[0] * h.shape[0];
                const char* lp[XTB_MAX_LEAVES];
                for (int k = 0; k < prog->n_leaves; ++k)
                    lp[k] = leaf_ptr[k] + (piece ? half * s.stride[k][0] * dtype_size(leaf_dtype[k]) : 0);
                char* op = out_ptr + (piece ? half * s.stride[OUT][0] * dtype_size(out_dtype) : 0);
                XTB_TRY(launch_space(prog, h, lp, leaf_dtype, op, out_dtype, rt, w64, ctx, depth + 1));
            }
            return XTB_OK;
        }
    }
    EwParams p;
    memset(&p, 0, sizeof(p));
    p.prog.n_insns = prog->n_insns;
    p.prog.result_type = rt;
    memcpy(p.prog.insns, prog->insns, sizeof(xtb_insn) * prog->n_insns);
    memcpy(p.prog.imms, prog->imms, sizeof(uint64_t) * XTB_MAX_IMMS);
    p.ndim = s.ndim;
    p.n_leaves = prog->n_leaves;
    p.out_rt = rt;
    for (int d = 0; d < s.ndim; ++d) p.shape[d] = s.shape[d];

    // vector width: 16 bytes of the widest element type on the path
    int V = w64 ? 2 : 4;
    const int inner = s.ndim - 1;
    auto classify = [&](const char* base, int dtype, const int64_t* st, bool is_out) -> int {
        const int sz = dtype_size(dtype);
        if (s.shape[inner] == 1) return is_out ? MODE_GATHER : MODE_BCAST;
        if (st[inner] == 0 && !is_out) return MODE_BCAST;
        if (st[inner] != 1) return MODE_GATHER;
        const int64_t vb = (int64_t) V * sz;
        if (((uintptr_t) base) % vb != 0) return MODE_GATHER;
        for (int d = 0; d < inner; ++d)
            if ((st[d] * sz) % vb != 0) return MODE_GATHER;
        return MODE_VEC;
    };
    for (int k = 0; k < prog->n_leaves; ++k) {
        EwLeaf& L = p.leaf[k];
        L.dtype = leaf_dtype[k];
        L.ptr = leaf_ptr[k];
        for (int d = 0; d < s.ndim; ++d) L.stride[d] = s.stride[k][d];
        L.mode = classify(L.ptr, L.dtype, L.stride, false);
    }
    p.out.dtype = out_dtype;
    p.out.ptr = out_ptr;
    for (int d = 0; d < s.ndim; ++d) p.out.stride[d] = s.stride[OUT][d];
    p.out.mode = classify(p.out.ptr, p.out.dtype, p.out.stride, true);

    // If nothing can use vector access, one element per thread keeps accesses coalesced.
    bool any_vec = p.out.mode == MODE_VEC;
    for (int k = 0; k < prog->n_leaves; ++k) any_vec |= p.leaf[k].mode == MODE_VEC;
    if (!any_vec) {
        V = 1;
        for (int k = 0; k < prog->n_leaves; ++k)
            if (p.leaf[k].mode == MODE_VEC) p.leaf[k].mode = MODE_GATHER;
    }
    // Transposed leaves: an operand that cannot be read along the output's fast dim but is
    // contiguous along another one goes through the tiled kernel.
    {
        int ti = -1, n_tile = 0;
        bool tile_ok = s.ndim >= 2 && s.shape[inner] >= 16 && (p.out.stride[inner] == 1 || p.out.stride[inner] == -1);
        for (int k = 0; k < prog->n_leaves && tile_ok; ++k) {
            EwLeaf& L = p.leaf[k];
            if (L.mode != MODE_GATHER) continue;
            int best = -1;
            for (int d = 0; d < s.ndim - 1; ++d)
                if ((L.stride[d] == 1 || L.stride[d] == -1) && s.shape[d] >= 16) best = d;
            if (best < 0) continue;                       // plain strided gather: stays in the generic path
            if (ti >= 0 && best != ti) { tile_ok = false; break; }
            ti = best;
            if (n_tile >= kMaxTileLeaves) { tile_ok = false; break; }
            p.tile_slot[k] = n_tile++;
        }
        if (tile_ok && ti >= 0) {
            for (int k = 0; k < prog->n_leaves; ++k)
                if (p.leaf[k].mode == MODE_GATHER && (p.leaf[k].stride[ti] == 1 || p.leaf[k].stride[ti] == -1) && s.shape[ti] >= 16)
                    p.leaf[k].mode = MODE_TILE;
            p.tile_i = ti;
            p.ntile_i = (uint32_t) ((s.shape[ti] + kTile - 1) / kTile);
            p.ntile_j = (uint32_t) ((s.shape[inner] + kTile - 1) / kTile);
            p.div_nti = make_fastdiv(p.ntile_i);
            p.div_ntj = make_fastdiv(p.ntile_j);
            for (int d = 0; d < s.ndim; ++d) p.div_dim[d] = make_fastdiv((uint32_t) std::min<int64_t>(s.shape[d], 0x7fffffff));
            bool dims_ok = true;
            for (int d = 0; d < s.ndim; ++d) dims_ok = dims_ok && s.shape[d] < 0x7fffffffLL;
            // 32-bit offsets are fine when every operand spans < 2^31 elements
            bool ok32 = true;
            for (int k = 0; k <= prog->n_leaves; ++k) {
                const EwLeaf& L = k < prog->n_leaves ? p.leaf[k] : p.out;
                int64_t span = 0;
                for (int d = 0; d < s.ndim; ++d) span += (L.stride[d] < 0 ? -L.stride[d] : L.stride[d]) * (s.shape[d] - 1);
                ok32 = ok32 && span < 0x7fffffffLL;
            }
            p.idx32 = ok32;
            if (dims_ok) return dispatch_tile(prog, p, ctx, w64);
            for (int k = 0; k < prog->n_leaves; ++k)
                if (p.leaf[k].mode == MODE_TILE) p.leaf[k].mode = MODE_GATHER;
        }
    }
    // 32-bit offsets + linear shortcut for the rank <= 3 kernels
    {
        bool ok32 = true;
        auto fill32 = [&](EwLeaf& L, bool is_out) {
            int64_t span = 0, dense = 1;
            bool linear = (L.mode == MODE_VEC);
            for (int d = s.ndim - 1; d >= 0; --d) {
                const int64_t st = L.stride[d];
                span += (st < 0 ? -st : st) * (s.shape[d] - 1);
                if (st != dense) linear = false;
                dense *= s.shape[d];
                if (d < 4) L.s32[d] = (int32_t) st;
            }
            if (span >= 0x7fffffffLL) ok32 = false;
            if (linear && s.total < 0x7fffffffLL) L.mode = MODE_LINEAR;
            (void) is_out;
        };
        for (int k = 0; k < prog->n_leaves; ++k) fill32(p.leaf[k], false);
        fill32(p.out, true);
        for (int d = 0; d < s.ndim; ++d) ok32 = ok32 && s.shape[d] < 0x7fffffffLL;
        p.idx32 = ok32;
        if (!ok32 || !ew_nd_ok_rank(s.ndim)) {
            // the generic kernel does not know MODE_LINEAR
            for (int k = 0; k < prog->n_leaves; ++k) if (p.leaf[k].mode == MODE_LINEAR) p.leaf[k].mode = MODE_VEC;
            if (p.out.mode == MODE_LINEAR) p.out.mode = MODE_VEC;
        }
    }
    const int64_t vpr = (s.shape[inner] + V - 1) / V;
    int64_t rows = 1;
    for (int d = 0; d < inner; ++d) rows *= s.shape[d];
    p.vec_per_row = (uint32_t) vpr;
    p.total_vec = rows * vpr;
    if (vpr > 0x7fffffff) XTB_FAIL(XTB_ERR_UNSUPPORTED, "inner extent too large");
    p.div_vpr = make_fastdiv((uint32_t) vpr);
    for (int d = 0; d < s.ndim; ++d) p.div_dim[d] = make_fastdiv((uint32_t) std::min<int64_t>(s.shape[d], 0x7fffffff));
    return dispatch_ew(prog, p, ctx, w64, V);
}

}  // namespace xtb

using namespace xtb;

extern "C" int xtb_assign(const xtb_program* prog, const xtb_operand* out, const xtb_operand* leaves) {
    if (!prog || !out) XTB_FAIL(XTB_ERR_INVALID, "null argument");
    if (prog->n_leaves > 0 && !leaves) XTB_FAIL(XTB_ERR_INVALID, "null leaves");
    if (out->ndim < 0 || out->ndim > XTB_MAX_DIM) XTB_FAIL(XTB_ERR_INVALID, "out rank %d", out->ndim);
    if (out->dtype < 0 || out->dtype >= XTB_DTYPE_COUNT) XTB_FAIL(XTB_ERR_INVALID, "bad out dtype");
    int32_t leaf_dt[XTB_MAX_LEAVES];
    for (int k = 0; k < prog->n_leaves && k < XTB_MAX_LEAVES; ++k) leaf_dt[k] = leaves[k].dtype;
    int rt = 0;
    bool w64 = false;
    XTB_TRY(validate_program(prog, leaf_dt, &rt, &w64));
    if (dtype_size(out->dtype) == 8) w64 = true;

    Space s;
    s.ndim = out->ndim;
    s.n_ops = prog->n_leaves + 1;
    const int OUT = prog->n_leaves;
    for (int d = 0; d < out->ndim; ++d) {
        if (out->shape[d] < 0) XTB_FAIL(XTB_ERR_INVALID, "negative extent");
        s.shape[d] = out->shape[d];
        s.stride[OUT][d] = out->shape[d] == 1 ? 0 : out->stride[d];
    }
    for (int k = 0; k < prog->n_leaves; ++k) {
        char what[32];
        snprintf(what, sizeof(what), "leaf %d", k);
        XTB_TRY(align_operand(&leaves[k], s.ndim, s.shape, s.stride[k], what));
    }
    for (int d = 0; d < s.ndim; ++d)
        if (s.shape[d] == 0) return XTB_OK;  // nothing to assign
    DeviceCtx* ctx;
    XTB_TRY(get_ctx(&ctx));
    XTB_LAUNCH_LOCK(ctx);

    sort_space_by(&s, OUT);
    collapse_space(&s);
    if (s.ndim == 0) {  // 0-d / single element
        s.ndim = 1;
        s.shape[0] = 1;
        for (int k = 0; k < s.n_ops; ++k) s.stride[k][0] = 0;
    }

    const char* leaf_ptr[XTB_MAX_LEAVES];
    int leaf_dtype[XTB_MAX_LEAVES];
    for (int k = 0; k < prog->n_leaves; ++k) {
        leaf_dtype[k] = leaves[k].dtype;
        leaf_ptr[k] = operand_ptr(&leaves[k], dtype_size(leaves[k].dtype));
    }
    return launch_space(prog, s, leaf_ptr, leaf_dtype, operand_ptr(out, dtype_size(out->dtype)), out->dtype, rt, w64, ctx, 0);
}
