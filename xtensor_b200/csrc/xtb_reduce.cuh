// xtb_reduce.cuh -- axis-reduction kernels behind xtb_reduce.
//
// Replaces reduce_immediate (include/xtensor/reducers/xreducer.hpp:289-565) and
// the per-output nested loops of xreducer_stepper::aggregate_impl (:1778-1868).
// The collapsed iteration space is split into kept dims (K outputs) and reduced
// dims (R inputs per output):
//   k_reduce_inner : innermost dim is reduced  (xreducer.hpp:482-511, "inner_stride == 1")
//                    G lanes share one output: 128-bit loads along the reduced
//                    dim, warp-shuffle tree (G <= 32) or block tree (G = 256)
//   k_reduce_outer : innermost dim is kept     (xreducer.hpp:512-551, row streaming)
//                    one thread owns V adjacent outputs and walks the reduced
//                    dims in the reference's order; rows may be split over
//                    blockIdx.y, partials are merged by a second launch.
// Any expression can be fused in front of the reduction (xreducer over xfunction).
#pragma once
#include "xtb_ew.cuh"

namespace xtb {

struct RdLeaf {
    const char* ptr;
    int64_t kstride[XTB_MAX_DIM];
    int64_t rstride[XTB_MAX_DIM];
    int32_t dtype;
    int32_t mode;  // LeafMode along the vector dim
};

struct RdParams {
    DevProgram prog;
    int32_t nk, nr;
    int64_t kshape[XTB_MAX_DIM], rshape[XTB_MAX_DIM];
    FastDiv kdiv[XTB_MAX_DIM], rdiv[XTB_MAX_DIM];
    int64_t K, R;
    int32_t binop;      // XTB_OP_ADD / MUL / MAXIMUM / MINIMUM
    int32_t acc_rt;     // accumulator register type
    int32_t in_rt;      // register type of the program result
    int32_t n_leaves;
    RdLeaf leaf[XTB_MAX_LEAVES];
    // destination: either the final output (strided over kept dims) or partials[K][nsplit]
    char* out_ptr;
    int64_t out_kstride[XTB_MAX_DIM];
    int32_t out_dtype;
    int32_t out_vec_ok;     // outer kernel: V results can be stored with one vector store
    int32_t nsplit;         // >1: write acc-typed partials to part_ptr[ko * nsplit + split]
    int32_t has_initial;
    char* part_ptr;
    int64_t chunk;          // reduced positions (outer) / r-vectors (inner) per split
    uint64_t identity_bits;
    uint64_t initial_bits;
    // inner kernel
    int32_t G;              // lanes per output: 1..32 or 256
    uint32_t vpr;           // vectors per innermost reduced row
    FastDiv vpr_div;
    int64_t rvec_total;     // r-vectors per output
    // outer kernel
    uint32_t kvpr;          // vectors per innermost kept row
    FastDiv kvpr_div;
    int64_t kvec_total;
};

// position of one thread: kept coords + reduced coords (+ column along the vector dim)
struct RdFetch {
    const RdParams& p;
    int64_t koff_idx[XTB_MAX_DIM];  // kept coordinates
    int64_t ridx[XTB_MAX_DIM];      // reduced coordinates
    int64_t col;                    // first element along the vector dim
    int nvalid;
    bool vec_is_reduced;

    XTB_DEV int64_t offset_of(const RdLeaf& L) const {
        int64_t off = 0;
        for (int d = 0; d < p.nk; ++d) off += koff_idx[d] * L.kstride[d];
        for (int d = 0; d < p.nr; ++d) off += ridx[d] * L.rstride[d];
        return off;
    }
    template <class S, int V> XTB_DEV void load(int k, int dt, S (&x)[V]) const {
        const RdLeaf& L = p.leaf[k];
        const int sz = dtype_size(dt);
        const int64_t vstride = vec_is_reduced ? L.rstride[p.nr - 1] : L.kstride[p.nk - 1];
        const char* ptr = L.ptr + (offset_of(L) + col * vstride) * sz;
        if (L.mode == MODE_BCAST) {
            S s = load_elem<S>(ptr, dt);
#pragma unroll
            for (int v = 0; v < V; ++v) x[v] = s;
        } else if (L.mode == MODE_VEC && nvalid == V) {
            load_vec<S, V>(ptr, dt, x);
        } else {
            const int64_t step = vstride * sz;
#pragma unroll
            for (int v = 0; v < V; ++v) x[v] = (v < nvalid) ? load_elem<S>(ptr + v * step, dt) : S(0);
        }
    }
};

// accumulate policies: run-time (interpreter) and compile-time (static programs)
struct DynAcc {
    template <class S, int V> static XTB_DEV void cast_in(const RdParams& p, S (&x)[V]) {
        if (p.in_rt != p.acc_rt) exec_unary<S, V, false>(XTB_OP_CAST, p.in_rt, p.acc_rt, x);
    }
    template <class S, int V> static XTB_DEV void step(const RdParams& p, S (&acc)[V], const S (&x)[V]) {
        exec_binary<S, V, false>(p.binop, p.acc_rt, acc, x);
    }
};
template <int BINOP, int ACC_RT> struct StaticAcc {
    template <class S, int V> static XTB_DEV void cast_in(const RdParams&, S (&)[V]) {}
    template <class S, int V> static XTB_DEV void step(const RdParams&, S (&acc)[V], const S (&x)[V]) {
        exec_binary_c<BINOP, ACC_RT, S, V>(acc, x);
    }
};

template <class S> XTB_DEV void rd_decompose(uint32_t lin, int n, const int64_t* shape, const FastDiv* div, int64_t* idx) {
    for (int d = n - 1; d > 0; --d) {
        const uint32_t q = fd_div(lin, div[d]);
        idx[d] = lin - q * (uint32_t) shape[d];
        lin = q;
    }
    idx[0] = lin;
}

// final value -> (merge initial) -> cast -> store
template <class S> XTB_DEV void rd_finish_store(const RdParams& p, char* dst, S v) {
    S a[1] = {v};
    if (p.has_initial) {
        S b[1] = {(S) p.initial_bits};
        exec_binary<S, 1, false>(p.binop, p.acc_rt, a, b);
    }
    store_elem<S>(dst, p.out_dtype, p.acc_rt, a[0]);
}

// ---- innermost dim reduced ------------------------------------------------------
template <class Eval, class Acc, class S, int V>
__global__ void __launch_bounds__(256) k_reduce_inner(const __grid_constant__ RdParams p) {
    __shared__ S smem[8];
    const int G = p.G;
    const int tid = threadIdx.x;
    const int lane_in_group = (G >= 256) ? tid : (tid & (G - 1));
    const int groups_per_block = (G >= 256) ? 1 : 256 / G;
    const int group = (G >= 256) ? 0 : tid / G;
    const int64_t jbeg = (int64_t) blockIdx.y * p.chunk;
    int64_t jend = jbeg + p.chunk;
    if (jend > p.rvec_total) jend = p.rvec_total;
    const int64_t out_groups = (p.K + groups_per_block - 1) / groups_per_block;
    for (int64_t gb = blockIdx.x; gb < out_groups; gb += gridDim.x) {
        const int64_t ko = gb * groups_per_block + group;
        const bool active = ko < p.K;
        RdFetch f{p, {0}, {0}, 0, V, true};
        S acc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = (S) p.identity_bits;
        if (active) {
            rd_decompose<S>((uint32_t) ko, p.nk, p.kshape, p.kdiv, f.koff_idx);
            const int64_t RL = p.rshape[p.nr - 1];
            for (int64_t j = jbeg + lane_in_group; j < jend; j += G) {
                int64_t cv = j;
                if (p.nr > 1) {
                    const uint32_t ro = fd_div((uint32_t) j, p.vpr_div);
                    cv = j - (int64_t) ro * p.vpr;
                    rd_decompose<S>(ro, p.nr - 1, p.rshape, p.rdiv, f.ridx);
                }
                f.col = cv * V;
                const int64_t rem = RL - f.col;
                f.nvalid = rem < V ? (int) rem : V;
                S x[V];
                Eval::template run<S, V>(p.prog, f, x);
                Acc::template cast_in<S, V>(p, x);
                if (f.nvalid < V) {
#pragma unroll
                    for (int v = 0; v < V; ++v)
                        if (v >= f.nvalid) x[v] = (S) p.identity_bits;
                }
                Acc::template step<S, V>(p, acc, x);
            }
        }
        // combine the V lanes of the thread, then the G lanes of the group
        S r[1] = {acc[0]};
#pragma unroll
        for (int v = 1; v < V; ++v) {
            S y[1] = {acc[v]};
            Acc::template step<S, 1>(p, r, y);
        }
        const int W = G >= 32 ? 32 : G;
        for (int o = W >> 1; o > 0; o >>= 1) {
            S y[1];
            if constexpr (sizeof(S) == 8) y[0] = __shfl_xor_sync(0xffffffffu, (unsigned long long) r[0], o);
            else y[0] = __shfl_xor_sync(0xffffffffu, r[0], o);
            Acc::template step<S, 1>(p, r, y);
        }
        if (G >= 256) {
            __syncthreads();
            if ((tid & 31) == 0) smem[tid >> 5] = r[0];
            __syncthreads();
            if (tid < 32) {
                r[0] = tid < 8 ? smem[tid] : (S) p.identity_bits;
                for (int o = 4; o > 0; o >>= 1) {
                    S y[1];
                    if constexpr (sizeof(S) == 8) y[0] = __shfl_xor_sync(0xffffffffu, (unsigned long long) r[0], o);
                    else y[0] = __shfl_xor_sync(0xffffffffu, r[0], o);
                    Acc::template step<S, 1>(p, r, y);
                }
            }
        }
        if (active && lane_in_group == 0) {
            if (p.nsplit > 1) {
                store_elem<S>(p.part_ptr + (ko * p.nsplit + blockIdx.y) * dtype_size(p.acc_rt), p.acc_rt, p.acc_rt, r[0]);
            } else {
                int64_t off = 0;
                for (int d = 0; d < p.nk; ++d) off += f.koff_idx[d] * p.out_kstride[d];
                rd_finish_store<S>(p, p.out_ptr + off * dtype_size(p.out_dtype), r[0]);
            }
        }
    }
}

// ---- innermost dim kept ----------------------------------------------------------
template <class Eval, class Acc, class S, int V>
__global__ void __launch_bounds__(256) k_reduce_outer(const __grid_constant__ RdParams p) {
    const int64_t rbeg = (int64_t) blockIdx.y * p.chunk;
    int64_t rend = rbeg + p.chunk;
    if (rend > p.R) rend = p.R;
    for (int64_t t = (int64_t) blockIdx.x * 256 + threadIdx.x; t < p.kvec_total; t += (int64_t) gridDim.x * 256) {
        RdFetch f{p, {0}, {0}, 0, V, false};
        const uint32_t krow = fd_div((uint32_t) t, p.kvpr_div);
        const uint32_t cv = (uint32_t) t - krow * p.kvpr;
        // kept coordinates: all but the innermost from krow; innermost coordinate = col
        if (p.nk > 1) rd_decompose<S>(krow, p.nk - 1, p.kshape, p.kdiv, f.koff_idx);
        f.koff_idx[p.nk - 1] = 0;
        f.col = (int64_t) cv * V;
        const int64_t rem = p.kshape[p.nk - 1] - f.col;
        f.nvalid = rem < V ? (int) rem : V;
        S acc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = (S) p.identity_bits;
        if (p.nr == 1) {
#pragma unroll(Eval::kUnroll)
            for (int64_t r = rbeg; r < rend; ++r) {
                f.ridx[0] = r;
                S x[V];
                Eval::template run<S, V>(p.prog, f, x);
                Acc::template cast_in<S, V>(p, x);
                Acc::template step<S, V>(p, acc, x);
            }
        } else {
            for (int64_t r = rbeg; r < rend; ++r) {
                rd_decompose<S>((uint32_t) r, p.nr, p.rshape, p.rdiv, f.ridx);
                S x[V];
                Eval::template run<S, V>(p.prog, f, x);
                Acc::template cast_in<S, V>(p, x);
                Acc::template step<S, V>(p, acc, x);
            }
        }
        if (p.nsplit > 1) {
            const int asz = dtype_size(p.acc_rt);
            // partials[K][nsplit]: the merge pass reads each output's partials contiguously
            const int64_t ko0 = (int64_t) krow * p.kshape[p.nk - 1] + f.col;
            char* dst = p.part_ptr + (ko0 * p.nsplit + blockIdx.y) * asz;
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (v < f.nvalid) store_elem<S>(dst + (int64_t) v * p.nsplit * asz, p.acc_rt, p.acc_rt, acc[v]);
        } else {
            int64_t off = f.col * p.out_kstride[p.nk - 1];
            for (int d = 0; d < p.nk - 1; ++d) off += f.koff_idx[d] * p.out_kstride[d];
            const int osz = dtype_size(p.out_dtype);
            char* dst = p.out_ptr + off * osz;
            if (p.has_initial) {
                S b[V];
#pragma unroll
                for (int v = 0; v < V; ++v) b[v] = (S) p.initial_bits;
                Acc::template step<S, V>(p, acc, b);
            }
            if (p.out_vec_ok && f.nvalid == V) {
                store_vec<S, V>(dst, p.out_dtype, p.acc_rt, acc);
            } else {
                const int64_t step = p.out_kstride[p.nk - 1] * osz;
#pragma unroll
                for (int v = 0; v < V; ++v)
                    if (v < f.nvalid) store_elem<S>(dst + v * step, p.out_dtype, p.acc_rt, acc[v]);
            }
        }
    }
}

template <class Eval, class Acc, class S, int V>
static int launch_reduce(const RdParams& p, DeviceCtx* ctx, bool inner, const char* evname) {
    char name[96];
    if (inner) {
        const int gpb = p.G >= 256 ? 1 : 256 / p.G;
        const int64_t out_groups = (p.K + gpb - 1) / gpb;
        dim3 grid((unsigned) std::min<int64_t>(out_groups, (int64_t) ctx->sm_count * 64), (unsigned) p.nsplit);
        snprintf(name, sizeof(name), "k_reduce_inner<%s,S%d,V%d>[G=%d,split=%d]", evname, (int) sizeof(S) * 8, V, p.G, p.nsplit);
        k_reduce_inner<Eval, Acc, S, V><<<grid, 256, 0, ctx->stream>>>(p);
    } else {
        const int64_t blocks = (p.kvec_total + 255) / 256;
        dim3 grid((unsigned) std::min<int64_t>(blocks, (int64_t) ctx->sm_count * 64), (unsigned) p.nsplit);
        snprintf(name, sizeof(name), "k_reduce_outer<%s,S%d,V%d>[split=%d]", evname, (int) sizeof(S) * 8, V, p.nsplit);
        k_reduce_outer<Eval, Acc, S, V><<<grid, 256, 0, ctx->stream>>>(p);
    }
    note_launch(name);
    return check_launch(name);
}

// registry of compile-time (program, reducer, accumulator type) combinations
struct StaticReduceEntry {
    const sprogs::SP* prog;
    int binop;
    int acc_rt;
    const char* name;
    int (*launch)(const RdParams&, DeviceCtx*, bool inner);
};
struct StaticReduceTable {
    const StaticReduceEntry* entries;
    int n;
};
StaticReduceTable static_reduce_table();

}  // namespace xtb
