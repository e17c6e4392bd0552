// xtb_reduce.cu -- xtb_reduce: host-side planning of axis reductions.
// Follows the contract of reduce_immediate (include/xtensor/reducers/xreducer.hpp:289-565):
// axes sorted / unique / in range (:336-350), output shape with or without
// keep_dims (shape_computation :190-239), empty axes = elementwise reduce(init, x)
// (:316-325), zero-size reduced extent yields init (:1782-1785), xt::initial merged
// once at the end (:552-563).  Kernels: xtb_reduce.cuh.
#include <cfloat>
#include <climits>
#include <cstdlib>
#include "xtb_reduce.cuh"
#include "xtb_jit.hpp"

namespace xtb {

int comm_allreduce(DeviceCtx* ctx, void* buf, size_t count, int dtype, int op);  // xtb_comm.cu
int comm_world();                                                                 // xtb_comm.cu
bool comm_p2p_params(DeviceCtx* ctx, P2pParams* w);                               // xtb_comm.cu

static uint64_t identity_bits(int op, int rt) {
    union { uint64_t u; double d; float f[2]; int32_t i32[2]; uint32_t u32[2]; int64_t i64; } v;
    v.u = 0;
    // nan_min / nan_max start from NaN (XTENSOR_REDUCER_FUNCTION(nanmin, ..., std::nan("0")), xmath.hpp:2427-2442);
    // for integer accumulators they are the plain extremes
    if (op == XTB_RED_NANMIN || op == XTB_RED_NANMAX) {
        if (rt == XTB_F32) { v.u32[0] = 0x7fc00000u; return v.u; }
        if (rt == XTB_F64) { v.u = 0x7ff8000000000000ull; return v.u; }
        op = (op == XTB_RED_NANMIN) ? XTB_RED_MIN : XTB_RED_MAX;
    }
    switch (op) {
        case XTB_RED_SUM: break;
        case XTB_RED_PROD:
            switch (rt) {
                case XTB_F32: v.f[0] = 1.0f; break;
                case XTB_F64: v.d = 1.0; break;
                default: v.u = 1; break;
            }
            break;
        case XTB_RED_MAX:  // numeric_limits<T>::lowest()
            switch (rt) {
                case XTB_F32: v.f[0] = -FLT_MAX; break;
                case XTB_F64: v.d = -DBL_MAX; break;
                case XTB_I32: v.i32[0] = INT32_MIN; break;
                case XTB_I64: v.i64 = INT64_MIN; break;
                default: v.u = 0; break;
            }
            break;
        case XTB_RED_MIN:  // numeric_limits<T>::max()
            switch (rt) {
                case XTB_F32: v.f[0] = FLT_MAX; break;
                case XTB_F64: v.d = DBL_MAX; break;
                case XTB_I32: v.i32[0] = INT32_MAX; break;
                case XTB_U32: v.u32[0] = UINT32_MAX; break;
                case XTB_I64: v.i64 = INT64_MAX; break;
                default: v.u = UINT64_MAX; break;
            }
            break;
    }
    return v.u;
}

static int binop_of(int op) {
    switch (op) {
        case XTB_RED_SUM: return XTB_OP_ADD;
        case XTB_RED_PROD: return XTB_OP_MUL;
        case XTB_RED_MAX: return XTB_OP_MAXIMUM;
        case XTB_RED_MIN: return XTB_OP_MINIMUM;
        case XTB_RED_NANMIN: return XTB_OP_NANMIN;
        case XTB_RED_NANMAX: return XTB_OP_NANMAX;
        default: return -1;
    }
}

static int run_reduce_kernel(const xtb_program* prog, const RdParams& p, DeviceCtx* ctx, bool inner, bool w64, int V) {
    const bool no_static = options().no_static != 0;
    if (!no_static && p.in_rt == p.acc_rt && (V == (w64 ? 2 : 4) || V == 1) && p.K < 0x7fffffff) {
        const StaticReduceTable t = V == 1 ? static_reduce_table_v1() : static_reduce_table();
        for (int i = 0; i < t.n; ++i) {
            const StaticReduceEntry& e = t.entries[i];
            if (e.binop == p.binop && e.acc_rt == p.acc_rt && sprogs::is64(*e.prog) == w64 && sprog_matches(*e.prog, prog))
                return e.launch(p, ctx, inner);
        }
    }
    if (!no_static && p.in_rt == p.acc_rt && (V == (w64 ? 2 : 4) || V == 1) && p.K < 0x7fffffff && jit_program_ok(prog) &&
        jit_worthwhile(p.K * std::max<int64_t>(p.R, 1))) {
        // run-time specialisation of the reduction kernel for this (program, reducer, accumulator)
        const RdLaunch g = reduce_geometry(p, ctx, inner, p.out_dtype == p.acc_rt);
        static const int kinds[4] = {JIT_RED_ROWS_EXACT, JIT_RED_INNER_WARP, JIT_RED_INNER_BLOCK, JIT_RED_OUTER};
        static const char* names[4] = {"k_reduce_rows_exact<jit>", "k_reduce_inner_warp<jit>", "k_reduce_inner_block<jit>", "k_reduce_outer<jit>"};
        JitSpec spec;
        spec.kind = kinds[g.kind];
        spec.w64 = w64;
        spec.V = V;
        spec.binop = p.binop;
        spec.acc_rt = p.acc_rt;
        void* fn = nullptr;
        if (jit_get(ctx, prog, spec, &fn) == XTB_OK) {
            if (jit_launch(fn, g.gx, g.gy, 256, 0, ctx->stream, &p) == XTB_OK) {
                note_launch(names[g.kind]);
                return XTB_OK;
            }   // a failed launch falls through to the interpreter kernel
        }
    }
    if (w64) {
        if (V == 2) return launch_reduce<InterpEval, DynAcc, uint64_t, 2>(p, ctx, inner, "interp");
        return launch_reduce<InterpEval, DynAcc, uint64_t, 1>(p, ctx, inner, "interp");
    }
    if (V == 4) return launch_reduce<InterpEval, DynAcc, uint32_t, 4>(p, ctx, inner, "interp");
    return launch_reduce<InterpEval, DynAcc, uint32_t, 1>(p, ctx, inner, "interp");
}

struct ReducePlanIn {
    const xtb_program* prog;
    bool empty = false;    // a reduced extent is 0: every output is init
    int n_leaves;
    const char* leaf_ptr[XTB_MAX_LEAVES];
    int leaf_dtype[XTB_MAX_LEAVES];
    Space space;       // operands: leaves..., out (index n_leaves); reduced[] flags set
    int op;            // xtb_reduce_op
    int binop, acc_rt, in_rt;
    bool w64;
    char* out_ptr;
    int out_dtype;
    bool has_initial;
    uint64_t initial_bits, identity;
    int fin_op = 0, fin_rt = 0;
    uint64_t fin_imm = 0;
    bool want_xchg = false;   // the caller asked for the cross-GPU merge of the result
};

// second pass over partials[nsplit][K]
template <int BINOP, int ACC_RT>
static void launch_merge(const RdParams& fp, DeviceCtx* ctx, bool xchg, const P2pParams& xw) {
    if (xchg) {
        launch_pdl(k_reduce_merge<BINOP, ACC_RT, true>, (unsigned) ((fp.K + 127) / 128), kMergeWarps * 32, 0, ctx->stream, fp, xw);
    } else if (fp.K < 128) {
        launch_pdl(k_reduce_merge_few<BINOP, ACC_RT>, (unsigned) fp.K, 256, 0, ctx->stream, fp);
    } else {
        launch_pdl(k_reduce_merge<BINOP, ACC_RT, false>, (unsigned) ((fp.K + 127) / 128), kMergeWarps * 32, 0, ctx->stream, fp, xw);
    }
}
template <int BINOP>
static int launch_merge_rt(int acc_rt, const RdParams& fp, DeviceCtx* ctx, bool xchg, const P2pParams& xw) {
    switch (acc_rt) {
        case XTB_I32: launch_merge<BINOP, XTB_I32>(fp, ctx, xchg, xw); return XTB_OK;
        case XTB_U32: launch_merge<BINOP, XTB_U32>(fp, ctx, xchg, xw); return XTB_OK;
        case XTB_I64: launch_merge<BINOP, XTB_I64>(fp, ctx, xchg, xw); return XTB_OK;
        case XTB_U64: launch_merge<BINOP, XTB_U64>(fp, ctx, xchg, xw); return XTB_OK;
        case XTB_F32: launch_merge<BINOP, XTB_F32>(fp, ctx, xchg, xw); return XTB_OK;
        case XTB_F64: launch_merge<BINOP, XTB_F64>(fp, ctx, xchg, xw); return XTB_OK;
        default: XTB_FAIL(XTB_ERR_INVALID, "merge: accumulator type %d", acc_rt);
    }
}

// out[k] = finalize(initial (+) partials[0][k] (+) partials[1][k] (+) ...), optionally exchanged across GPUs
// in the same kernel (xchg)
static int merge_partials(const RdParams& fp, DeviceCtx* ctx, bool xchg, const P2pParams& xw, const char* what) {
    if (fp.K >= 0x7fffffffLL) XTB_FAIL(XTB_ERR_UNSUPPORTED, "too many outputs");
    // the reported kernel stays the first pass (the one that moves the data), with the merge appended
    char name[128];
    snprintf(name, sizeof(name), "%.90s + %s", xtb_last_kernel(), what);
    switch (fp.binop) {
        case XTB_OP_ADD: XTB_TRY(launch_merge_rt<XTB_OP_ADD>(fp.acc_rt, fp, ctx, xchg, xw)); break;
        case XTB_OP_MUL: XTB_TRY(launch_merge_rt<XTB_OP_MUL>(fp.acc_rt, fp, ctx, xchg, xw)); break;
        case XTB_OP_MAXIMUM: XTB_TRY(launch_merge_rt<XTB_OP_MAXIMUM>(fp.acc_rt, fp, ctx, xchg, xw)); break;
        case XTB_OP_MINIMUM: XTB_TRY(launch_merge_rt<XTB_OP_MINIMUM>(fp.acc_rt, fp, ctx, xchg, xw)); break;
        case XTB_OP_NANMIN: XTB_TRY(launch_merge_rt<XTB_OP_NANMIN>(fp.acc_rt, fp, ctx, xchg, xw)); break;
        case XTB_OP_NANMAX: XTB_TRY(launch_merge_rt<XTB_OP_NANMAX>(fp.acc_rt, fp, ctx, xchg, xw)); break;
        default: XTB_FAIL(XTB_ERR_INVALID, "merge: operator %d", fp.binop);
    }
    note_launch(name);
    return check_launch("k_reduce_merge");
}

static int plan_and_launch(ReducePlanIn& in, DeviceCtx* ctx) {
    Space& s = in.space;
    const int OUT = in.n_leaves;
    collapse_space(&s);
    RdParams p;
    memset(&p, 0, sizeof(p));
    p.prog.n_insns = in.prog->n_insns;
    p.prog.result_type = in.in_rt;
    memcpy(p.prog.insns, in.prog->insns, sizeof(xtb_insn) * in.prog->n_insns);
    memcpy(p.prog.imms, in.prog->imms, sizeof(uint64_t) * XTB_MAX_IMMS);
    p.binop = in.binop;
    p.acc_rt = in.acc_rt;
    p.in_rt = in.in_rt;
    p.n_leaves = in.n_leaves;
    p.identity_bits = in.identity;
    p.initial_bits = in.initial_bits;
    p.has_initial = in.has_initial;
    p.out_ptr = in.out_ptr;
    p.out_dtype = in.out_dtype;
    p.fin_op = in.fin_op;
    p.fin_rt = in.fin_rt;
    p.fin_imm = in.fin_imm;

    // split dims into kept / reduced lists (order preserved)
    int kd[XTB_MAX_DIM], rd[XTB_MAX_DIM];
    int nk = 0, nr = 0;
    for (int d = 0; d < s.ndim; ++d) (s.reduced[d] ? rd[nr++] : kd[nk++]) = d;
    const bool inner = !in.empty && s.ndim > 0 && s.reduced[s.ndim - 1];
    p.K = 1;
    p.R = 1;
    // always at least one kept and one reduced dim (dummy extent-1 dims in front)
    int ko = 0, ro = 0;
    if (nk == 0) { p.kshape[0] = 1; ko = 1; }
    if (nr == 0) { p.rshape[0] = 1; ro = 1; }
    for (int i = 0; i < nk; ++i) {
        p.kshape[ko + i] = s.shape[kd[i]];
        p.K *= s.shape[kd[i]];
        p.out_kstride[ko + i] = s.stride[OUT][kd[i]];
        for (int k = 0; k < in.n_leaves; ++k) p.leaf[k].kstride[ko + i] = s.stride[k][kd[i]];
    }
    for (int i = 0; i < nr; ++i) {
        p.rshape[ro + i] = s.shape[rd[i]];
        p.R *= s.shape[rd[i]];
        for (int k = 0; k < in.n_leaves; ++k) p.leaf[k].rstride[ro + i] = s.stride[k][rd[i]];
    }
    p.nk = nk + ko;
    p.nr = nr + ro;
    if (p.K >= 0x7fffffff) XTB_FAIL(XTB_ERR_UNSUPPORTED, "more than 2^31 reduction outputs");
    if (p.nr > 1 && p.R >= 0x7fffffff) XTB_FAIL(XTB_ERR_UNSUPPORTED, "non-mergeable reduced axes with more than 2^31 elements");
    for (int d = 0; d < p.nk; ++d) p.kdiv[d] = make_fastdiv((uint32_t) p.kshape[d]);
    for (int d = 0; d < p.nr; ++d) p.rdiv[d] = make_fastdiv((uint32_t) std::min<int64_t>(p.rshape[d], 0x7fffffff));

    int V = in.w64 ? 2 : 4;
    const int64_t vlen = inner ? p.rshape[p.nr - 1] : p.kshape[p.nk - 1];
    auto classify = [&](const char* base, int dtype, const int64_t* kst, const int64_t* rst) -> int {
        const int sz = dtype_size(dtype);
        const int64_t vs = inner ? rst[p.nr - 1] : kst[p.nk - 1];
        if (vlen == 1 || vs == 0) return MODE_BCAST;
        if (vs != 1) return MODE_GATHER;
        const int64_t vb = (int64_t) V * sz;
        if (((uintptr_t) base) % vb != 0) return MODE_GATHER;
        for (int d = 0; d < p.nk; ++d)
            if (!(d == p.nk - 1 && !inner) && (kst[d] * sz) % vb != 0) return MODE_GATHER;
        for (int d = 0; d < p.nr; ++d)
            if (!(d == p.nr - 1 && inner) && (rst[d] * sz) % vb != 0) return MODE_GATHER;
        return MODE_VEC;
    };
    bool any_vec = false;
    for (int k = 0; k < in.n_leaves; ++k) {
        RdLeaf& L = p.leaf[k];
        L.ptr = in.leaf_ptr[k];
        L.dtype = in.leaf_dtype[k];
        L.mode = classify(L.ptr, L.dtype, L.kstride, L.rstride);
        any_vec |= L.mode == MODE_VEC;
    }
    if (!any_vec && V > 1) {
        // no leaf can be read with 128-bit loads (odd pitch, offset view): scalar access, where any unit-stride
        // leaf counts as "vector" again and the unpredicated loaders apply
        V = 1;
        for (int k = 0; k < in.n_leaves; ++k) p.leaf[k].mode = classify(p.leaf[k].ptr, p.leaf[k].dtype, p.leaf[k].kstride, p.leaf[k].rstride);
    }

    if (in.empty) p.R = 0;
    // parallelisation
    const int64_t target_threads = (int64_t) ctx->sm_count * 2048;
    p.nsplit = 1;
    if (inner) {
        const int64_t RL = p.rshape[p.nr - 1];
        const int64_t vpr = (RL + V - 1) / V;
        if (vpr > 0x7fffffff) XTB_FAIL(XTB_ERR_UNSUPPORTED, "reduced extent too large");
        p.vpr = (uint32_t) vpr;
        p.vpr_div = make_fastdiv(p.vpr);
        p.rvec_total = (p.R / RL) * vpr;
        if (p.nr > 1 && p.rvec_total >= 0x7fffffff) XTB_FAIL(XTB_ERR_UNSUPPORTED, "reduction too large for split axes");
        {
            bool all_vec = p.nr == 1 && RL % V == 0;
            for (int k = 0; k < in.n_leaves; ++k) all_vec = all_vec && p.leaf[k].mode == MODE_VEC;
            p.inner_fast = all_vec ? 1 : 0;
        }
        int G = 1;
        while (G < 32 && G < p.rvec_total) G <<= 1;
        if (p.rvec_total > 32 * 8 && p.K * 32 < target_threads) G = 256;
        if (options().reduce_g == 32 && G == 256) G = 32;
        if (options().reduce_g == 256 && p.rvec_total > 32 * 8) G = 256;
        p.G = G;
        p.chunk = p.rvec_total;
        if (G == 256 && p.K * 256 < target_threads && p.rvec_total > 256 * 16) {
            // whole waves of resident CTAs (64 registers: 4 per SM), as for the row-streaming kernel below: measured
            // on a 2^26 full reduction, 592 CTAs 5.07 TB/s, 1024 CTAs 4.66, 296 CTAs 3.9 (tools/reduce_bench.py)
            const int64_t slots = (int64_t) ctx->sm_count * 4;
            const double pass_bytes = (double) p.R * (double) p.K * dtype_size(p.leaf[0].dtype);
            const int64_t waves = pass_bytes >= 1.5 * (double) slots * 768.0 * 1024.0 ? 2 : 1;
            const int q = 16 / dtype_size(p.acc_rt);
            int64_t want = std::max<int64_t>(1, slots * waves / p.K);
            if (want > q) want = want / q * q;      // partial rows stay 16-byte multiples without overshooting the wave
            if (options().reduce_split > 0) want = options().reduce_split;
            int64_t maxsplit = p.rvec_total / (256 * 8);
            int64_t ns = std::max<int64_t>(1, std::min(want, maxsplit));
            ns = std::min<int64_t>(ns, 1024);
            if (ns > 1) {
                p.chunk = (p.rvec_total + ns - 1) / ns;
                p.chunk = (p.chunk + 255) / 256 * 256;  // whole block strides
                p.nsplit = (int) ((p.rvec_total + p.chunk - 1) / p.chunk);
                p.nsplit = (p.nsplit + q - 1) / q * q;
            }
        }
    } else {
        const int64_t KL = p.kshape[p.nk - 1];
        const int64_t kvpr = (KL + V - 1) / V;
        if (kvpr > 0x7fffffff) XTB_FAIL(XTB_ERR_UNSUPPORTED, "kept extent too large");
        p.kvpr = (uint32_t) kvpr;
        p.kvpr_div = make_fastdiv(p.kvpr);
        p.kvec_total = (p.K / KL) * kvpr;
        if (p.kvec_total >= 0x7fffffff) XTB_FAIL(XTB_ERR_UNSUPPORTED, "too many outputs");
        p.chunk = p.R;
        // Row splits.  Measured on the B200 (tools/split_sweep.py, profiles/r02_split_sweep.log): a CTA of this kernel
        // streams ~15 GB/s when the GPU is saturated and 40 GB/s alone, all CTAs of a wave finish within a few percent
        // of each other, so what costs time is a last wave with only a few CTAs in it (a few CTAs over a multiple of
        // the resident count: up to +40 %), and long CTAs (a ragged end).  Hence: WHOLE waves of sm_count x 3 resident
        // CTAs -- two for a pass of about a gigabyte, four from 2 GB, eight (CTAs of ~2.4 MB at 8 GB) from 4 GB; every
        // CTA costs one 4 KB partial row that the merge kernel reads back, which is what caps the count.
        const int64_t gx = (p.kvec_total + 255) / 256;
        const int64_t slots = (int64_t) ctx->sm_count * 3;
        if (gx < slots * 2 && p.R > 32) {
            const double pass_bytes = (double) p.R * (double) p.K * dtype_size(p.leaf[0].dtype);
            const double wb = pass_bytes / ((double) slots * 768.0 * 1024.0);
            const int64_t waves = wb >= 10.0 ? 8 : (wb >= 5.0 ? 4 : (wb >= 1.5 ? 2 : 1));
            // gx column blocks rarely divide a wave: take the first wave count >= `waves` that fills >= 93 %
            int64_t want = 1;
            double best_fill = 0.0;
            for (int64_t w = waves; w <= 2 * waves + 1; ++w) {
                const int64_t n = std::max<int64_t>(1, slots * w / gx);
                const int64_t ctas = gx * n;
                const double fill = (double) ctas / (double) ((ctas + slots - 1) / slots * slots);   // of the waves it really takes
                if (fill > best_fill) { best_fill = fill; want = n; }
                if (fill >= 0.93) break;
            }
            int64_t maxsplit = p.R / 16;
            int64_t ns = std::max<int64_t>(1, std::min(want, maxsplit));
            ns = std::min<int64_t>(ns, 4096);
            if (options().reduce_split > 0) ns = std::max<int64_t>(1, std::min<int64_t>(options().reduce_split, maxsplit));
            if (ns > 1) {
                p.chunk = (p.R + ns - 1) / ns;
                p.chunk = (p.chunk + kRdFlush - 1) / kRdFlush * kRdFlush;   // whole summation blocks per split
                p.nsplit = (int) ((p.R + p.chunk - 1) / p.chunk);
            }
        }
        // Split rows are already summed out of the reference's order: use blocked summation there, which keeps
        // every fp32 chain short (<= kRdFlush terms per level).  Unsplit, the kernel adds row after row exactly
        // like reduce_immediate (xreducer.hpp:512-551) and stays bit-identical to it.
        p.two_level = p.nsplit > 1 ? 1 : 0;
        bool all_vec = p.nr == 1 && KL % V == 0, some_vec = false;
        for (int k = 0; k < in.n_leaves; ++k) {
            all_vec = all_vec && (p.leaf[k].mode == MODE_VEC || p.leaf[k].mode == MODE_BCAST);
            some_vec = some_vec || p.leaf[k].mode == MODE_VEC;
        }
        p.outer_fast = all_vec && some_vec ? 1 : 0;
    }

    // ---- cross-GPU merge of the result (the reduced axis is the sharded one) ----
    //   fused   : k_reduce_merge<XCHG> exchanges each merged output over NVLink peer memory before the one store
    //   staged  : local result -> dense accumulator-typed staging buffer -> collective -> final store.
    // Either way xt::initial and the finalize step are applied ONCE, after the cross-GPU merge.
    const int asz = dtype_size(p.acc_rt);
    P2pParams xw;
    memset(&xw, 0, sizeof(xw));
    bool fused = false, staged = false;
    if (in.want_xchg && comm_world() > 1) {
        const size_t words = (size_t) p.K * (size_t) (asz / 4);
        fused = p.nsplit > 1 && p.K >= 128 && words <= kP2pMaxWords && comm_p2p_params(ctx, &xw) && xw.world > 1;
        staged = !fused;
    }
    const size_t part_bytes = p.nsplit > 1 ? ((size_t) p.nsplit * (size_t) p.K * asz + 255) / 256 * 256 : 0;
    const size_t stage_bytes = staged ? (size_t) p.K * asz : 0;
    char* scratch = nullptr;
    if (part_bytes + stage_bytes > 0) {
        void* sp = nullptr;
        XTB_TRY(ensure_scratch(ctx, part_bytes + stage_bytes, &sp));
        scratch = (char*) sp;
    }
    if (p.nsplit > 1) p.part_ptr = scratch;
    const RdParams final_view = p;      // the caller's output, initial and finalize step
    if (staged) {
        p.out_ptr = scratch + part_bytes;
        p.out_dtype = p.acc_rt;
        int64_t st = 1;
        for (int d = p.nk - 1; d >= 0; --d) {
            p.out_kstride[d] = st;
            st *= p.kshape[d];
        }
        p.has_initial = 0;
        p.fin_op = 0;
    }
    p.out_vec_ok = 0;
    if (!inner && V > 1) {
        const int osz = dtype_size(p.out_dtype);
        bool ok = p.out_kstride[p.nk - 1] == 1 && ((uintptr_t) p.out_ptr) % ((int64_t) V * osz) == 0;
        for (int d = 0; d < p.nk - 1 && ok; ++d) ok = (p.out_kstride[d] * osz) % ((int64_t) V * osz) == 0;
        p.out_vec_ok = ok;
    }
    // preconditions of k_reduce_rows_exact
    p.exact_rows = 0;
    if (inner && p.nk == 1 && p.nr == 1 && p.G <= 32 && p.nsplit == 1 && V > 1 && p.rshape[0] % V == 0 &&
        p.rvec_total == p.G && !p.has_initial && (p.out_dtype == p.acc_rt || p.fin_op != 0) && p.in_rt == p.acc_rt &&
        (p.kshape[0] == 1 || p.out_kstride[0] == 1) && p.K < 0x7fffffffLL) {
        bool ok = true;
        for (int k = 0; k < in.n_leaves; ++k) {
            const RdLeaf& L = p.leaf[k];
            int64_t span = (L.kstride[0] < 0 ? -L.kstride[0] : L.kstride[0]) * (p.kshape[0] - 1) +
                           (L.rstride[0] < 0 ? -L.rstride[0] : L.rstride[0]) * (p.rshape[0] - 1);
            ok = ok && span < 0x7fffffffLL && L.mode != MODE_GATHER;
        }
        p.exact_rows = ok;
    }
    XTB_TRY(run_reduce_kernel(in.prog, p, ctx, inner, in.w64, V));
    if (p.nsplit > 1)
        XTB_TRY(merge_partials(p, ctx, fused, xw, fused ? "k_reduce_merge[+p2p exchange]" : p.K < 128 ? "k_reduce_merge_few" : "k_reduce_merge"));
    if (staged) {
        XTB_TRY(comm_allreduce(ctx, p.out_ptr, (size_t) p.K, p.acc_rt, in.op));
        RdParams f = final_view;
        f.nsplit = 1;
        f.part_ptr = p.out_ptr;
        XTB_TRY(merge_partials(f, ctx, false, xw, "allreduce + k_reduce_merge[final store]"));
    }
    return XTB_OK;
}

}  // namespace xtb

using namespace xtb;

// ---- decomposition of reductions the single-pass kernels serve badly ------------------------------------------
// Two shapes of problem are rewritten into two calls of xtb_reduce_fin over a small accumulator-typed temporary
// (pure descriptor work; measured before: 0.4 TB/s and 2.6 TB/s, tools/reduce_bench.py):
//   narrow : the innermost dim is kept but there are < 1024 outputs in all (sum over axis 0 of (2^20, 64), over axes
//            {0,1} of (4096,4096,16)): a thread owns 4 outputs, so a CTA of the row-streaming kernel would be mostly
//            idle.  The outermost reduced axis (extent E) is split into (E / m, m) and only its outer part is reduced
//            first, which keeps rows of m x (everything behind it) >= 32 K elements -- the shape the kernel is good
//            at; the (m, ...) temporary is reduced by the second call.  E need not be a multiple of m: the E mod m
//            trailing positions are reduced into an extra slice of the temporary.
//   mixed  : the innermost dim is reduced together with an outer one and a kept dim lies between them (axes {0,2} of
//            (4096,4096,16)): lanes sharing an output would gather 64-byte pieces a row pitch apart.  The reduced
//            axes in front of the last kept dim are reduced first (row streaming, everything behind stays), the rest
//            on the temporary.
// Both passes use the caller's reducer; xt::initial, the finalize step and the cross-GPU merge belong to the second.
// The result is a different (still fixed) summation order, as for every split reduction; integer results are exact.
static thread_local int t_decompose_depth = 0;

static int reduce_decomposed(int op, int acc_type, const xtb_program* prog, const xtb_operand* leaves, int ndim,
                             const int64_t* shape, int n_axes, const int32_t* axes, int keep_dims, const void* initial,
                             const xtb_operand* out, int allreduce, const xtb_finalize* fin, bool* handled) {
    *handled = false;
    if (options().no_decompose || t_decompose_depth >= 2 || n_axes == 0 || ndim + 1 > XTB_MAX_DIM || prog->n_leaves < 1) return XTB_OK;
    bool red[XTB_MAX_DIM] = {false};
    for (int i = 0; i < n_axes; ++i) red[axes[i]] = true;
    int64_t N = 1, K = 1;
    int li = -1, a0 = -1;
    for (int d = 0; d < ndim; ++d) {
        if (shape[d] <= 0) return XTB_OK;
        N *= shape[d];
        if (!red[d]) K *= shape[d];
        if (shape[d] > 1) {
            li = d;
            if (red[d] && a0 < 0) a0 = d;
        }
    }
    if (a0 < 0 || N < (1ll << 20)) return XTB_OK;
    bool kept_after_a0 = false;
    for (int d = a0 + 1; d < ndim; ++d) kept_after_a0 = kept_after_a0 || (!red[d] && shape[d] > 1);
    const bool mixed = red[li] && kept_after_a0;
    const bool narrow = !red[li] && K < 1024;
    if (!mixed && !narrow) return XTB_OK;

    // axes of the first pass (indices in the ORIGINAL space) and the split of a0
    bool peel[XTB_MAX_DIM] = {false};
    int64_t m = 1;
    if (mixed) {
        int last_kept = -1;
        for (int d = 0; d < ndim; ++d)
            if (!red[d] && shape[d] > 1) last_kept = d;
        int64_t shrink = 1;
        for (int d = 0; d < last_kept; ++d)
            if (red[d] && shape[d] > 1) { peel[d] = true; shrink *= shape[d]; }
        if (shrink < 4) return XTB_OK;
    } else {
        int64_t after = 1;
        bool other_reduced = false;
        for (int d = a0 + 1; d < ndim; ++d) {
            after *= shape[d];
            other_reduced = other_reduced || (red[d] && shape[d] > 1);
        }
        peel[a0] = true;
        if (after < 16384) {
            m = 32768 / after;
            const int64_t E = shape[a0];
            if (E < 4 * m) m = E / 4;
            if (m < 2) {
                if (!other_reduced) return XTB_OK;
                m = 1;
            } else {
                // a nearby divisor of E avoids the remainder slice
                for (int64_t c = m; c >= std::max<int64_t>(2, m / 2); --c)
                    if (E % c == 0) { m = c; break; }
            }
        } else if (!other_reduced) {
            return XTB_OK;
        }
    }
    const int64_t E = shape[a0];
    const int64_t q = m > 1 ? E / m : E, rem = m > 1 ? E - q * m : 0;
    const bool split = m > 1;
    const int nd1 = ndim + (split ? 1 : 0);
    auto map_dim = [&](int d) { return (split && d > a0) ? d + 1 : d; };   // original dim -> dim of the split space (a0 -> its outer part)

    // every leaf at full rank with explicit strides (0 where it broadcasts)
    xtb_operand l1[XTB_MAX_LEAVES], l1r[XTB_MAX_LEAVES];
    int64_t shape1[XTB_MAX_DIM], tshape[XTB_MAX_DIM];
    for (int d = 0; d < ndim; ++d) shape1[map_dim(d)] = shape[d];
    if (split) { shape1[a0] = q; shape1[a0 + 1] = m; }
    for (int k = 0; k < prog->n_leaves; ++k) {
        int64_t st[XTB_MAX_DIM];
        char what[32];
        snprintf(what, sizeof(what), "leaf %d", k);
        XTB_TRY(align_operand(&leaves[k], ndim, shape, st, what));
        xtb_operand o = leaves[k];
        o.ndim = nd1;
        for (int d = 0; d < ndim; ++d) {
            const int e = map_dim(d);
            o.shape[e] = st[d] == 0 ? 1 : shape[d];
            o.stride[e] = st[d];
        }
        if (split) {
            const bool b = st[a0] == 0;
            o.shape[a0] = b ? 1 : q;
            o.stride[a0] = b ? 0 : st[a0] * m;
            o.shape[a0 + 1] = b ? 1 : m;
            o.stride[a0 + 1] = b ? 0 : st[a0];
        }
        l1[k] = o;
        if (rem > 0) {
            // the E mod m trailing positions of a0: outer extent rem, inner extent 1
            xtb_operand r = o;
            const bool b = st[a0] == 0;
            r.offset += b ? 0 : q * m * st[a0];
            r.shape[a0] = b ? 1 : rem;
            r.stride[a0] = b ? 0 : st[a0];
            r.shape[a0 + 1] = 1;
            r.stride[a0 + 1] = 0;
            l1r[k] = r;
        }
    }
    // temporary: the split space with the peeled axes reduced to 1 (keep_dims), dense, accumulator-typed
    int32_t axes1[XTB_MAX_DIM];
    int na1 = 0;
    for (int d = 0; d < ndim; ++d)
        if (peel[d]) axes1[na1++] = map_dim(d);
    for (int d = 0; d < nd1; ++d) tshape[d] = shape1[d];
    for (int i = 0; i < na1; ++i) tshape[axes1[i]] = 1;
    const int64_t mslices = split ? m + (rem > 0 ? 1 : 0) : 1;
    if (split) tshape[a0 + 1] = mslices;
    int64_t tn = 1;
    for (int d = 0; d < nd1; ++d) tn *= tshape[d];
    const int asz = dtype_size(acc_type);
    struct Tmp { void* p = nullptr; ~Tmp() { if (p) xtb_free(p); } } tmp;
    XTB_TRY(xtb_malloc((size_t) tn * asz, &tmp.p));
    xtb_operand t{};
    t.base = tmp.p;
    t.dtype = acc_type;
    t.ndim = nd1;
    {
        int64_t stv = 1;
        for (int d = nd1 - 1; d >= 0; --d) {
            t.shape[d] = tshape[d];
            t.stride[d] = tshape[d] == 1 ? 0 : stv;
            stv *= tshape[d];
        }
    }
    struct Depth { Depth() { ++t_decompose_depth; } ~Depth() { --t_decompose_depth; } } depth_guard;
    // pass 1 (and the remainder slice)
    {
        xtb_operand t1 = t;
        if (split) t1.shape[a0 + 1] = m;
        XTB_TRY(xtb_reduce_fin(op, acc_type, prog, l1, nd1, shape1, na1, axes1, 1, nullptr, &t1, 0, nullptr));
        if (rem > 0) {
            int64_t shr[XTB_MAX_DIM];
            for (int d = 0; d < nd1; ++d) shr[d] = shape1[d];
            shr[a0] = rem;
            shr[a0 + 1] = 1;
            xtb_operand tr = t;
            tr.offset += m * t.stride[a0 + 1];
            tr.shape[a0 + 1] = 1;
            tr.stride[a0 + 1] = 0;
            XTB_TRY(xtb_reduce_fin(op, acc_type, prog, l1r, nd1, shr, na1, axes1, 1, nullptr, &tr, 0, nullptr));
        }
    }
    // pass 2: the caller's reducer over the temporary (identity map), every original reduced axis plus the inner part of a0
    xtb_program p2{};
    p2.n_leaves = 1;
    p2.insns[p2.n_insns++] = xtb_insn{(uint8_t) XTB_OP_PUSH, (uint8_t) acc_type, (uint8_t) XTB_SRC_LEAF, 0};
    int32_t axes2[XTB_MAX_DIM];
    int na2 = 0;
    for (int d = 0; d < nd1; ++d) {
        bool r2 = split && d == a0 + 1;
        for (int i = 0; i < n_axes && !r2; ++i) r2 = map_dim(axes[i]) == d;
        if (r2) axes2[na2++] = d;
    }
    xtb_operand o2 = *out;
    if (split && keep_dims) {
        // the caller's output has no dim for the inner part of a0: insert an extent-1 dim
        if (out->ndim != ndim) XTB_FAIL(XTB_ERR_SHAPE, "reducer output has rank %d, expected %d", out->ndim, ndim);
        o2.ndim = nd1;
        for (int d = 0; d < ndim; ++d) {
            o2.shape[map_dim(d)] = out->shape[d];
            o2.stride[map_dim(d)] = out->stride[d];
        }
        o2.shape[a0] = out->shape[a0];
        o2.stride[a0] = out->stride[a0];
        o2.shape[a0 + 1] = 1;
        o2.stride[a0 + 1] = 0;
    }
    XTB_TRY(xtb_reduce_fin(op, acc_type, &p2, &t, nd1, tshape, na2, axes2, keep_dims, initial, &o2, allreduce, fin));
    *handled = true;
    return XTB_OK;
}

extern "C" int xtb_reduce(int op, int acc_type, const xtb_program* prog, const xtb_operand* leaves, int ndim,
                          const int64_t* shape, int n_axes, const int32_t* axes, int keep_dims, const void* initial,
                          const xtb_operand* out, int allreduce) {
    return xtb_reduce_fin(op, acc_type, prog, leaves, ndim, shape, n_axes, axes, keep_dims, initial, out, allreduce, nullptr);
}

extern "C" int xtb_reduce_fin(int op, int acc_type, const xtb_program* prog, const xtb_operand* leaves, int ndim,
                              const int64_t* shape, int n_axes, const int32_t* axes, int keep_dims, const void* initial,
                              const xtb_operand* out, int allreduce, const xtb_finalize* fin) {
    if (!prog || !out || (ndim > 0 && !shape)) XTB_FAIL(XTB_ERR_INVALID, "null argument");
    if (ndim < 0 || ndim > XTB_MAX_DIM) XTB_FAIL(XTB_ERR_INVALID, "rank %d out of range", ndim);
    if (n_axes < 0 || n_axes > ndim) XTB_FAIL(XTB_ERR_AXIS, "%d axes for rank %d", n_axes, ndim);
    if (n_axes > 0 && !axes) XTB_FAIL(XTB_ERR_INVALID, "null axes");
    const int binop = binop_of(op);
    if (binop < 0) XTB_FAIL(XTB_ERR_INVALID, "unknown reducer %d", op);
    if (acc_type < XTB_I32 || acc_type > XTB_F64) XTB_FAIL(XTB_ERR_INVALID, "accumulator must be a register type");
    // xreducer.hpp:336-350
    for (int i = 1; i < n_axes; ++i) {
        if (axes[i] < axes[i - 1]) XTB_FAIL(XTB_ERR_AXIS, "Reducing axes should be sorted.");
        if (axes[i] == axes[i - 1]) XTB_FAIL(XTB_ERR_AXIS, "Reducing axes should not contain duplicates.");
    }
    if (n_axes > 0 && (axes[0] < 0 || axes[n_axes - 1] > ndim - 1))
        XTB_FAIL(XTB_ERR_AXIS, "Axis %d out of bounds for reduction.", axes[n_axes - 1]);
    int32_t leaf_dt[XTB_MAX_LEAVES];
    for (int k = 0; k < prog->n_leaves && k < XTB_MAX_LEAVES; ++k) leaf_dt[k] = leaves[k].dtype;
    int rt = 0;
    bool w64 = false;
    XTB_TRY(validate_program(prog, leaf_dt, &rt, &w64));
    if (out->dtype < 0 || out->dtype >= XTB_DTYPE_COUNT) XTB_FAIL(XTB_ERR_INVALID, "bad out dtype");
    if (fin && fin->op != XTB_FIN_NONE) {
        if (fin->op != XTB_FIN_DIV && fin->op != XTB_FIN_DIV_SQRT) XTB_FAIL(XTB_ERR_INVALID, "unknown finalize step %d", fin->op);
        if (fin->type != XTB_F32 && fin->type != XTB_F64) XTB_FAIL(XTB_ERR_INVALID, "finalize runs in f32 or f64");
        // the finalize step converts by itself: a 32-bit accumulator keeps 32-bit slots whatever the output dtype
        if (dtype_size(acc_type) == 8) w64 = true;
    } else if (dtype_size(acc_type) == 8 || dtype_size(out->dtype) == 8) {
        w64 = true;
    }

    {
        bool handled = false;
        XTB_TRY(reduce_decomposed(op, acc_type, prog, leaves, ndim, shape, n_axes, axes, keep_dims, initial, out, allreduce, fin, &handled));
        if (handled) return XTB_OK;
    }
    ReducePlanIn in;
    in.prog = prog;
    in.n_leaves = prog->n_leaves;
    Space& s = in.space;
    s = Space();
    s.ndim = ndim;
    s.n_ops = prog->n_leaves + 1;
    const int OUT = prog->n_leaves;
    bool red[XTB_MAX_DIM] = {false};
    for (int i = 0; i < n_axes; ++i) red[axes[i]] = true;
    // output shape check (shape_computation, xreducer.hpp:190-239)
    const int out_rank = keep_dims ? ndim : ndim - n_axes;
    if (out->ndim != out_rank) XTB_FAIL(XTB_ERR_SHAPE, "reducer output has rank %d, expected %d", out->ndim, out_rank);
    int64_t K = 1, R = 1;
    for (int d = 0, od = 0; d < ndim; ++d) {
        if (shape[d] < 0) XTB_FAIL(XTB_ERR_INVALID, "negative extent");
        s.shape[d] = shape[d];
        s.reduced[d] = red[d];
        if (red[d]) {
            R *= shape[d];
            if (keep_dims) {
                if (out->shape[od] != 1) XTB_FAIL(XTB_ERR_SHAPE, "keep_dims output must have extent 1 on axis %d", d);
                ++od;
            }
            s.stride[OUT][d] = 0;
        } else {
            K *= shape[d];
            if (out->shape[od] != shape[d])
                XTB_FAIL(XTB_ERR_SHAPE, "reducer output extent %lld on dim %d, expected %lld", (long long) out->shape[od], od,
                         (long long) shape[d]);
            s.stride[OUT][d] = shape[d] == 1 ? 0 : out->stride[od];
            ++od;
        }
    }
    for (int k = 0; k < prog->n_leaves; ++k) {
        char what[32];
        snprintf(what, sizeof(what), "leaf %d", k);
        XTB_TRY(align_operand(&leaves[k], ndim, s.shape, s.stride[k], what));
        in.leaf_ptr[k] = operand_ptr(&leaves[k], dtype_size(leaves[k].dtype));
        in.leaf_dtype[k] = leaves[k].dtype;
    }
    if (K == 0) return XTB_OK;
    DeviceCtx* ctx;
    XTB_TRY(get_ctx(&ctx));
    XTB_LAUNCH_LOCK(ctx);
    in.op = op;
    if (fin && fin->op != XTB_FIN_NONE) {
        in.fin_op = fin->op;
        in.fin_rt = fin->type;
        in.fin_imm = fin->imm;
    }
    in.binop = binop;
    in.acc_rt = acc_type;
    in.in_rt = rt;
    in.w64 = w64;
    in.out_ptr = operand_ptr(out, dtype_size(out->dtype));
    in.out_dtype = out->dtype;
    in.has_initial = initial != nullptr;
    in.initial_bits = 0;
    if (initial) memcpy(&in.initial_bits, initial, dtype_size(acc_type));
    in.identity = identity_bits(op, acc_type);
    if (R == 0) {
        // zero-size reduction: every output is init (xreducer.hpp:1782-1785); drop the
        // empty axes so that the kernel loops zero times.
        for (int d = 0; d < ndim; ++d)
            if (red[d]) s.shape[d] = 1;
        in.empty = true;
    }
    in.want_xchg = allreduce != 0;
    return plan_and_launch(in, ctx);
}
