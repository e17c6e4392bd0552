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
#include "xtb_p2p.cuh"
#ifndef XTB_RTC
#include <cstdlib>
#endif

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
    int32_t nsplit;         // >1: write acc-typed partials to part_ptr[split * K + ko]
    int32_t has_initial;
    char* part_ptr;
    int64_t chunk;          // reduced positions (outer) / r-vectors (inner) per split
    uint64_t identity_bits;
    uint64_t initial_bits;
    // inner kernel
    int32_t exact_rows;     // k_reduce_rows_exact applies (host-verified preconditions)
    int32_t G;              // lanes per output: 1..32 or 256
    uint32_t vpr;           // vectors per innermost reduced row
    FastDiv vpr_div;
    int64_t rvec_total;     // r-vectors per output
    // outer kernel
    uint32_t kvpr;          // vectors per innermost kept row
    FastDiv kvpr_div;
    int64_t kvec_total;
    // accuracy: blocked summation.  The running accumulator is flushed into a second-level total after every
    // block of kRdFlush rows (outer kernel) / kRdInnerFlush vectors per lane (inner kernels); for fp32 sums the
    // second level is fp64 (RdTotal), so the only fp32 chains are the <= 8-term blocks and the merge trees.
    // Set whenever the result is not the reference's sequential order anyway (split rows, lanes sharing an
    // output); an unsplit strided-axis reduction keeps the reference's order bit for bit.
    int32_t two_level;
    int32_t outer_fast;     // outer kernel: whole batches may use the unpredicated loader (RdFastLoader)
    int32_t inner_fast;     // inner kernels (one reduced dim): the same for whole batches of a lane's vectors
    // finalize step fused into the last store (mean_functor::finalize, xblockwise_reducer_functors.hpp:146-186):
    //   0 none, 1: out = T(acc) / imm, 2: out = sqrt(T(acc) / imm), T = fin_rt (XTB_F32 / XTB_F64)
    int32_t fin_op;
    int32_t fin_rt;
    uint64_t fin_imm;
};
constexpr int kRdFlush = 8;        // = the staged batch of the outer kernel
constexpr int kRdInnerFlush = 4;   // = the staged batch of the inner kernels

// Leaf access of one thread.  The kept part of every leaf's offset is folded into a
// per-leaf base pointer once per output (or once per thread in the outer kernel); the
// loops over the reduced dims then only add `reduced index * stride`.
// NL = number of leaves known at compile time (static programs) or XTB_MAX_LEAVES.
template <int NL> struct RdFetch {
    const RdParams& p;
    const char* cur[NL];   // address of the current vector of each leaf
    int nvalid;
    bool vec_is_reduced;

    template <class S, int V> XTB_DEV void load(int k, int dt, S (&x)[V]) const {
        const RdLeaf& L = p.leaf[k];
        const int sz = dtype_size(dt);
        const char* ptr = cur[k];
        if (L.mode == MODE_VEC && nvalid == V) {
            load_vec<S, V>(ptr, dt, x);
        } else if (L.mode == MODE_BCAST) {
            S s = load_elem<S>(ptr, dt);
#pragma unroll
            for (int v = 0; v < V; ++v) x[v] = s;
        } else {
            const int64_t step = (vec_is_reduced ? L.rstride[p.nr - 1] : L.kstride[p.nk - 1]) * sz;
#pragma unroll
            for (int v = 0; v < V; ++v) x[v] = (v < nvalid) ? load_elem<S>(ptr + v * step, dt) : S(0);
        }
    }
};

// element offset of kept position `ko` (row-major over the kept dims) for one stride table
XTB_DEV int64_t rd_kept_offset(const RdParams& p, uint32_t ko, const int64_t* kstride) {
    int64_t off = 0;
    for (int d = p.nk - 1; d > 0; --d) {
        const uint32_t q = fd_div(ko, p.kdiv[d]);
        off += (int64_t) (ko - q * (uint32_t) p.kshape[d]) * kstride[d];
        ko = q;
    }
    return off + (int64_t) ko * kstride[0];
}
// element offset of reduced position `r` over the first n reduced dims
XTB_DEV int64_t rd_reduced_offset(const RdParams& p, uint32_t r, int n, const int64_t* rstride) {
    int64_t off = 0;
    for (int d = n - 1; d > 0; --d) {
        const uint32_t q = fd_div(r, p.rdiv[d]);
        off += (int64_t) (r - q * (uint32_t) p.rshape[d]) * rstride[d];
        r = q;
    }
    return off + (int64_t) r * rstride[0];
}

// accumulate policies: run-time (interpreter) and compile-time (static programs)
struct DynAcc {
    template <class S, int V> static XTB_DEV void cast_in(const RdParams& p, S (&x)[V]) {
        if (p.in_rt != p.acc_rt) exec_unary<S, V, false>(XTB_OP_CAST, p.in_rt, p.acc_rt, x);
    }
    template <class S, int V> static XTB_DEV void step(const RdParams& p, S (&acc)[V], const S (&x)[V]) {
        exec_binary<S, V, false>(p.binop, p.acc_rt, acc, x);
    }
    static XTB_DEV bool wide(const RdParams& p) { return p.binop == XTB_OP_ADD && p.acc_rt == XTB_F32; }
};
template <int BINOP, int ACC_RT> struct StaticAcc {
    template <class S, int V> static XTB_DEV void cast_in(const RdParams&, S (&)[V]) {}
    template <class S, int V> static XTB_DEV void step(const RdParams&, S (&acc)[V], const S (&x)[V]) {
        exec_binary_c<BINOP, ACC_RT, S, V>(acc, x);
    }
    static XTB_DEV constexpr bool wide(const RdParams&) { return BINOP == XTB_OP_ADD && ACC_RT == XTB_F32; }
};

// second level of the blocked summation: block results are merged into `tot` in the accumulator type, except
// for fp32 sums, whose blocks are added in fp64 (one conversion + one add per block) and rounded once at the end
template <class Acc, class S, int V> struct RdTotal {
    S tot[V];
    double totd[V];
    XTB_DEV void init(const RdParams& p) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
            tot[v] = (S) p.identity_bits;
            totd[v] = 0.0;
        }
    }
    // tot (+)= acc; acc = identity
    XTB_DEV void flush(const RdParams& p, S (&acc)[V]) {
        if (Acc::wide(p)) {
#pragma unroll
            for (int v = 0; v < V; ++v) totd[v] += (double) get<float>(acc[v]);
        } else {
            Acc::template step<S, V>(p, tot, acc);
        }
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = (S) p.identity_bits;
    }
    // acc = tot (+) acc
    XTB_DEV void finish(const RdParams& p, S (&acc)[V]) {
        if (Acc::wide(p)) {
#pragma unroll
            for (int v = 0; v < V; ++v) acc[v] = put<S>((float) (totd[v] + (double) get<float>(acc[v])));
        } else {
            Acc::template step<S, V>(p, tot, acc);
#pragma unroll
            for (int v = 0; v < V; ++v) acc[v] = tot[v];
        }
    }
};

template <class S> XTB_DEV S shfl_xor(S v, int o) {
    if constexpr (sizeof(S) == 8) return (S) __shfl_xor_sync(0xffffffffu, (unsigned long long) v, o);
    else return (S) __shfl_xor_sync(0xffffffffu, v, o);
}

// accumulator value -> finalize (mean: / count, stddev: sqrt) -> cast -> store
template <class T> XTB_DEV void rd_store_final_t(const RdParams& p, char* dst, T v) {
    if (p.fin_op == 0) {
        store_as<T>(dst, p.out_dtype, v);
    } else if (p.fin_rt == XTB_F32) {
        float x = (float) v / __uint_as_float((uint32_t) p.fin_imm);
        if (p.fin_op == 2) x = sqrtf(x);
        store_as<float>(dst, p.out_dtype, x);
    } else {
        double x = (double) v / __longlong_as_double((long long) p.fin_imm);
        if (p.fin_op == 2) x = sqrt(x);
        store_as<double>(dst, p.out_dtype, x);
    }
}
template <class S> XTB_DEV void rd_store_final(const RdParams& p, char* dst, S v) {
    XTB_TYPE_SWITCH(p.acc_rt, S, { rd_store_final_t<T>(p, dst, get<T>(v)); })
}

// final value -> (merge initial) -> finalize -> cast -> store
template <class S> XTB_DEV void rd_finish_store(const RdParams& p, char* dst, S v) {
    S a[1] = {v};
    if (p.has_initial) {
        S b[1] = {(S) p.initial_bits};
        exec_binary<S, 1, false>(p.binop, p.acc_rt, a, b);
    }
    rd_store_final<S>(p, dst, a[0]);
}

// final value of a typed merge -> (merge initial) -> cast -> store
template <int BINOP, int ACC_RT, class S> XTB_DEV void rd_finish_store_c(const RdParams& p, char* dst, S v) {
    S a[1] = {v};
    if (p.has_initial) {
        S b[1] = {(S) p.initial_bits};
        exec_binary_c<BINOP, ACC_RT, S, 1>(a, b);
    }
    using T = reg_t<ACC_RT>;
    rd_store_final_t<T>(p, dst, get<T>(a[0]));
}

// Second pass of a split reduction: out[k] = partials[0][k] (+) partials[1][k] (+) ... over part[nsplit][K].
// The merge operator and the accumulator type are compile-time (4 x 6 small instantiations): the kernel is
// a few hundred instructions, which matters because it runs for microseconds between two bandwidth-bound
// kernels (the run-time-typed version spent its time fetching instructions).
// A block owns 128 outputs (4 per lane, 128-bit loads); warp w of its 16 warps adds splits w, w+16, ...,
// eight loads in flight per lane, then the warps are combined in order through shared memory: a fixed
// order and a couple of memory round trips, however few outputs there are (the case that needs
// splitting at all).
//
// XCHG: the reduction runs over the axis that is sharded across GPUs, so the merged value is only this
// rank's partial.  The thread that finishes output k then exchanges it with the other ranks over NVLink
// peer memory (xtb_p2p.cuh) and combines the R partials in rank order before the one store: the local
// merge and the cross-GPU merge are one kernel, and the partial never travels through HBM in between.
// Word layout = k_allreduce_p2p on the accumulator type.
constexpr int kMergeWarps = 16;
template <int ACC_RT> struct MergeSlot { using type = std::conditional_t<dtype_size(ACC_RT) == 8, uint64_t, uint32_t>; };

template <int BINOP, int ACC_RT, bool XCHG>
__global__ void __launch_bounds__(kMergeWarps * 32) k_reduce_merge(const __grid_constant__ RdParams p, const __grid_constant__ P2pParams xw) {
    using S = typename MergeSlot<ACC_RT>::type;
    constexpr int asz = dtype_size(ACC_RT);
    __shared__ S sm[kMergeWarps][128];
    pdl_enter();
    uint32_t epoch = 0;
    if constexpr (XCHG) epoch = p2p_epoch(xw);
    constexpr int U = 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t k0 = ((int64_t) blockIdx.x * 32 + lane) * 4;
    S acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = (S) p.identity_bits;
    if (k0 < p.K) {
        const bool vec = p.K % 4 == 0;
        for (int s0 = warp; s0 < p.nsplit; s0 += kMergeWarps * U) {
            S x[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int sp = s0 + kMergeWarps * u;
                const char* src = p.part_ptr + ((int64_t) sp * p.K + k0) * asz;
                if (sp < p.nsplit && vec) {
                    load_vec<S, 4>(src, ACC_RT, x[u]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[u][i] = (sp < p.nsplit && k0 + i < p.K) ? load_elem<S>(src + i * asz, ACC_RT) : (S) p.identity_bits;
                }
            }
            // pairwise tree over the batch (absent splits hold the identity), then one add into the running value
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!(s0 + kMergeWarps * u < p.nsplit)) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[u][i] = (S) p.identity_bits;
                }
            }
#pragma unroll
            for (int st = 1; st < U; st <<= 1)
#pragma unroll
                for (int u = 0; u + st < U; u += 2 * st) exec_binary_c<BINOP, ACC_RT, S, 4>(x[u], x[u + st]);
            exec_binary_c<BINOP, ACC_RT, S, 4>(acc, x[0]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) sm[warp][lane * 4 + i] = acc[i];
    __syncthreads();
    // warp 0..3 finish one of the lane's four outputs each (coalesced stores along k)
    if (warp < 4) {
        const int64_t k = (int64_t) blockIdx.x * 128 + warp * 32 + lane;
        if (k < p.K) {
            const int c = warp * 32 + lane;
            // the warps' values are combined in a fixed pairwise tree
            S t[kMergeWarps][1];
#pragma unroll
            for (int w = 0; w < kMergeWarps; ++w) t[w][0] = sm[w][c];
#pragma unroll
            for (int st = 1; st < kMergeWarps; st <<= 1)
#pragma unroll
                for (int w = 0; w + st < kMergeWarps; w += 2 * st) exec_binary_c<BINOP, ACC_RT, S, 1>(t[w], t[w + st]);
            S r[1] = {t[0][0]};
            if constexpr (XCHG) {
                constexpr int W = sizeof(S) / 4;
                const S mine = r[0];
                uint32_t mw[W], got[kP2pMaxWorld][W];
                memcpy(mw, &mine, sizeof(S));
                p2p_exchange<W>(xw, epoch, (size_t) k * W, mw, got);
#pragma unroll
                for (int q = 0; q < kP2pMaxWorld; ++q) {
                    if (q < xw.world) {
                        S y[1] = {mine};
                        if (q != xw.rank) memcpy(&y[0], got[q], sizeof(S));
                        if (q == 0) r[0] = y[0];
                        else exec_binary_c<BINOP, ACC_RT, S, 1>(r, y);
                    }
                }
            }
            const int64_t off = rd_kept_offset(p, (uint32_t) k, p.out_kstride);
            rd_finish_store_c<BINOP, ACC_RT, S>(p, p.out_ptr + off * dtype_size(p.out_dtype), r[0]);
        }
    }
    if constexpr (XCHG) {
        __syncthreads();
        if (threadIdx.x == 0) p2p_finish(xw, epoch);
    }
}

// Few outputs (full reductions, K < 128): one block per output, thread t adds splits t, t+256, ...,
// then a fixed-shape tree over the block.
template <int BINOP, int ACC_RT>
__global__ void __launch_bounds__(256) k_reduce_merge_few(const __grid_constant__ RdParams p) {
    using S = typename MergeSlot<ACC_RT>::type;
    constexpr int asz = dtype_size(ACC_RT);
    __shared__ S sm[256];
    pdl_enter();
    const int t = threadIdx.x;
    const int64_t k = blockIdx.x;
    S acc[1] = {(S) p.identity_bits};
    for (int s0 = t; s0 < p.nsplit; s0 += 256 * 4) {
        S x[4][1];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int sp = s0 + 256 * u;
            x[u][0] = sp < p.nsplit ? load_elem<S>(p.part_ptr + ((int64_t) sp * p.K + k) * asz, ACC_RT) : (S) p.identity_bits;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (s0 + 256 * u < p.nsplit) exec_binary_c<BINOP, ACC_RT, S, 1>(acc, x[u]);
    }
    sm[t] = acc[0];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (t < o) {
            S a[1] = {sm[t]}, b[1] = {sm[t + o]};
            exec_binary_c<BINOP, ACC_RT, S, 1>(a, b);
            sm[t] = a[0];
        }
        __syncthreads();
    }
    if (t == 0) {
        const int64_t off = rd_kept_offset(p, (uint32_t) k, p.out_kstride);
        rd_finish_store_c<BINOP, ACC_RT, S>(p, p.out_ptr + off * dtype_size(p.out_dtype), sm[0]);
    }
}

// Stage U vectors of leaf K (and recursively of the following leaves) of a compile-time
// program.  dtype / element size are constants; the access mode is one uniform branch per leaf.
template <class Eval, class S, int V, int U, int K, int INV = 0> struct RdLeafLoader {
    template <class PF, class AddrFn>
    static XTB_DEV void run(const RdParams& p, const int (&nvalid)[U], int64_t gather_stride_elems_sel, PF& pf, AddrFn addr_of,
                            bool skip_invariant = false) {
        if constexpr (K < Eval::kLeaves) {
            constexpr int dt = Eval::template leaf_dtype<K>();
            constexpr int sz = dtype_size(dt);
            const RdLeaf& L = p.leaf[K];
            if constexpr (((INV >> K) & 1) != 0) {
                // compile-time invariant leaf (the host checked rstride == 0, vector access, whole vectors):
                // ONE vector, loaded by the first iteration only
                if (!skip_invariant) load_vec<S, V>(addr_of(L, sz, 0), dt, pf.inv[K]);
                RdLeafLoader<Eval, S, V, U, K + 1, INV>::run(p, nvalid, gather_stride_elems_sel, pf, addr_of, skip_invariant);
                return;
            }
            if (skip_invariant && L.rstride[0] == 0 && p.nr == 1) {
                RdLeafLoader<Eval, S, V, U, K + 1, INV>::run(p, nvalid, gather_stride_elems_sel, pf, addr_of, skip_invariant);
                return;
            }
            const char* addr[U];
#pragma unroll
            for (int u = 0; u < U; ++u) addr[u] = addr_of(L, sz, u);
            if (L.mode == MODE_VEC) {
                bool full = true;
#pragma unroll
                for (int u = 0; u < U; ++u) full = full && nvalid[u] == V;
                if (full) {
#pragma unroll
                    for (int u = 0; u < U; ++u) load_vec<S, V>(addr[u], dt, pf.pre[K][u]);
                } else {
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int v = 0; v < V; ++v) pf.pre[K][u][v] = (v < nvalid[u]) ? load_elem<S>(addr[u] + v * sz, dt) : S(0);
                }
            } else if (L.mode == MODE_BCAST) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const S v0 = nvalid[u] > 0 ? load_elem<S>(addr[u], dt) : S(0);
#pragma unroll
                    for (int v = 0; v < V; ++v) pf.pre[K][u][v] = v0;
                }
            } else {
                const int64_t step = (gather_stride_elems_sel ? L.rstride[p.nr - 1] : L.kstride[p.nk - 1]) * sz;
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int v = 0; v < V; ++v) pf.pre[K][u][v] = (v < nvalid[u]) ? load_elem<S>(addr[u] + v * step, dt) : S(0);
            }
            RdLeafLoader<Eval, S, V, U, K + 1, INV>::run(p, nvalid, gather_stride_elems_sel, pf, addr_of, skip_invariant);
        }
    }
};

// Whole batches of the outer kernel (host-verified: every leaf vector-accessed, every thread owns whole vectors,
// one reduced dim): U unpredicated 128-bit loads per leaf from `row0 + u * rstep`, nothing else -- the general
// loader above spends ~35 instructions per load on liveness and address arithmetic, which is what limited the
// fused map-reduce kernels (variance, amax) below the plain sum.
template <class Eval, class S, int V, int U, int K, int INV = 0> struct RdFastLoader {
    template <class PF, int NL>
    static XTB_DEV void run(const RdParams& p, const char* const (&row0)[NL], const int64_t (&rstep)[NL], PF& pf, bool skip_invariant) {
        if constexpr (K < Eval::kLeaves) {
            constexpr int dt = Eval::template leaf_dtype<K>();
            if constexpr (((INV >> K) & 1) != 0) {
                if (!skip_invariant) load_vec<S, V>(row0[K], dt, pf.inv[K]);
            } else if (p.leaf[K].mode == MODE_BCAST) {
                // one element per row, the same for the thread's V outputs (the index vector of argmin / argmax)
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const S v0 = load_elem<S>(row0[K] + u * rstep[K], dt);
#pragma unroll
                    for (int v = 0; v < V; ++v) pf.pre[K][u][v] = v0;
                }
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) load_vec<S, V>(row0[K] + u * rstep[K], dt, pf.pre[K][u]);
            }
            RdFastLoader<Eval, S, V, U, K + 1, INV>::template run<PF, NL>(p, row0, rstep, pf, skip_invariant);
        }
    }
};

// one lane's share of the reduced elements of output `ko`: vectors j = jbeg + lane, += stride
template <class Eval, class Acc, class S, int V, int NL>
XTB_DEV S rd_inner_partial(const RdParams& p, uint32_t ko, int64_t jbeg, int64_t jend, int lane, int stride) {
    RdFetch<NL> f{p, {nullptr}, V, true};
    const char* base[NL];
    int64_t vstep[NL];
#pragma unroll
    for (int k = 0; k < NL; ++k) {
        if (k < p.n_leaves) {
            const RdLeaf& L = p.leaf[k];
            const int sz = dtype_size(L.dtype);
            base[k] = L.ptr + rd_kept_offset(p, ko, L.kstride) * sz;
            vstep[k] = L.rstride[p.nr - 1] * (int64_t) (V * sz);
        }
    }
    S acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = (S) p.identity_bits;
    const int64_t RL = p.rshape[p.nr - 1];
    if constexpr (Eval::kPrefetch) {
        if (p.nr == 1) {
            // U vectors of every leaf in flight before any arithmetic
            constexpr int U = 4;
            PreFetch<NL, U, S, V> pf;
            // blocked summation: lanes interleave, so the order is never the reference's; keep chains short
            static_assert(U == kRdInnerFlush, "one block per staged batch");
            RdTotal<Acc, S, V> total;
            total.init(p);
            int64_t j = jbeg + lane;
            if (p.inner_fast) {
                // whole batches: unpredicated 128-bit loads (same batches, same order as the general loop below)
                const char* row0[NL];
                int64_t rs[NL];
#pragma unroll
                for (int k = 0; k < NL; ++k) {
                    rs[k] = (int64_t) stride * vstep[k];
                    row0[k] = base[k] + j * vstep[k];
                }
                if constexpr (V == 1) {
                    // scalar access (odd pitches): a lane's load is 4 or 8 bytes, so four times as many stay in flight;
                    // the summation blocks are the same kRdInnerFlush elements, hence the same result
                    constexpr int UB = 4 * U;
                    PreFetch<NL, UB, S, V> pfw;
                    for (; j + (int64_t) (UB - 1) * stride < jend; j += (int64_t) stride * UB) {
                        RdFastLoader<Eval, S, V, UB, 0>::template run<decltype(pfw), NL>(p, row0, rs, pfw, false);
#pragma unroll
                        for (int k = 0; k < NL; ++k) row0[k] += UB * rs[k];
#pragma unroll
                        for (int u = 0; u < UB; ++u) {
                            pfw.u = u;
                            S x[V];
                            Eval::template run<S, V>(p.prog, pfw, x);
                            Acc::template cast_in<S, V>(p, x);
                            Acc::template step<S, V>(p, acc, x);
                            if ((u + 1) % U == 0) total.flush(p, acc);
                        }
                    }
                }
                for (; j + (int64_t) (U - 1) * stride < jend; j += (int64_t) stride * U) {
                    RdFastLoader<Eval, S, V, U, 0>::template run<decltype(pf), NL>(p, row0, rs, pf, false);
#pragma unroll
                    for (int k = 0; k < NL; ++k) row0[k] += U * rs[k];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        pf.u = u;
                        S x[V];
                        Eval::template run<S, V>(p.prog, pf, x);
                        Acc::template cast_in<S, V>(p, x);
                        Acc::template step<S, V>(p, acc, x);
                    }
                    total.flush(p, acc);
                }
            }
            for (; j < jend; j += (int64_t) stride * U) {
                int nvalid[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int64_t ju = j + (int64_t) u * stride;
                    const int64_t rem = RL - ju * V;
                    nvalid[u] = ju < jend ? (rem < V ? (int) rem : V) : 0;
                }
                RdLeafLoader<Eval, S, V, U, 0>::run(p, nvalid, 1, pf, [&](const RdLeaf& L, int sz, int u) -> const char* {
                    const int64_t koff = rd_kept_offset(p, ko, L.kstride);
                    const int64_t ju = nvalid[u] > 0 ? j + (int64_t) u * stride : 0;
                    return L.ptr + (koff + ju * V * L.rstride[0]) * sz;
                });
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (nvalid[u] > 0) {
                        pf.u = u;
                        S x[V];
                        Eval::template run<S, V>(p.prog, pf, x);
                        Acc::template cast_in<S, V>(p, x);
                        if (nvalid[u] < V) {
#pragma unroll
                            for (int v = 0; v < V; ++v)
                                if (v >= nvalid[u]) x[v] = (S) p.identity_bits;
                        }
                        Acc::template step<S, V>(p, acc, x);
                    }
                }
                total.flush(p, acc);
            }
            total.finish(p, acc);
            S r0[1] = {acc[0]};
#pragma unroll
            for (int v = 1; v < V; ++v) {
                S y[1] = {acc[v]};
                Acc::template step<S, 1>(p, r0, y);
            }
            return r0[0];
        }
    }
    RdTotal<Acc, S, V> total2;
    total2.init(p);
    int in_block2 = 0;
    for (int64_t j = jbeg + lane; j < jend; j += stride) {
        int64_t cv = j;
        if (p.nr > 1) {
            const uint32_t ro = fd_div((uint32_t) j, p.vpr_div);
            cv = j - (int64_t) ro * p.vpr;
#pragma unroll
            for (int k = 0; k < NL; ++k)
                if (k < p.n_leaves)
                    f.cur[k] = base[k] + rd_reduced_offset(p, ro, p.nr - 1, p.leaf[k].rstride) * dtype_size(p.leaf[k].dtype) + cv * vstep[k];
        } else {
#pragma unroll
            for (int k = 0; k < NL; ++k)
                if (k < p.n_leaves) f.cur[k] = base[k] + cv * vstep[k];
        }
        const int64_t rem = RL - cv * V;
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
        if (++in_block2 >= kRdInnerFlush) {     // same block boundaries as the staged loop above
            total2.flush(p, acc);
            in_block2 = 0;
        }
    }
    total2.finish(p, acc);
    S r[1] = {acc[0]};
#pragma unroll
    for (int v = 1; v < V; ++v) {
        S y[1] = {acc[v]};
        Acc::template step<S, 1>(p, r, y);
    }
    return r[0];
}

// ---- innermost dim reduced, G <= 32 lanes per output ---------------------------------
// A warp produces 32 consecutive outputs per round: its 32/G groups each reduce G outputs
// one after the other (output g*G + it in iteration it), lane g*G + it keeps that result,
// so the round ends with ONE coalesced store of 32 results.
template <class Eval, class Acc, class S, int V>
__global__ void __launch_bounds__(256) k_reduce_inner_warp(const __grid_constant__ RdParams p) {
    constexpr int NL = Eval::kLeaves;
    const int G = p.G;
    const int lane = threadIdx.x & 31;
    const int li = lane & (G - 1);      // lane within its group
    const int g = lane / G;             // group within the warp
    const int64_t rounds = (p.K + 31) / 32;
    const int64_t warp0 = (int64_t) blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t jbeg = (int64_t) blockIdx.y * p.chunk;
    int64_t jend = jbeg + p.chunk;
    if (jend > p.rvec_total) jend = p.rvec_total;
    for (int64_t round = warp0; round < rounds; round += (int64_t) gridDim.x * 8) {
        const int64_t out_base = round * 32;
        S keep = (S) p.identity_bits;
        bool done = false;
        if constexpr (Eval::kPrefetch) {
            if (p.nr == 1 && p.rvec_total <= G && p.nsplit == 1) {
                // short rows: every lane owns at most one vector per output; stage the vectors of
                // U consecutive outputs of the group before reducing them
                constexpr int U = 4;
                const int64_t RL = p.rshape[0];
                const int64_t rem = RL - (int64_t) li * V;
                const int my_valid = li < p.rvec_total ? (rem < V ? (int) rem : V) : 0;
                PreFetch<NL, U, S, V> pf;
                for (int it0 = 0; it0 < G; it0 += U) {
                    int nvalid[U];
                    int64_t kos[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        kos[u] = out_base + g * G + it0 + u;
                        nvalid[u] = (it0 + u < G && kos[u] < p.K) ? my_valid : 0;
                    }
                    RdLeafLoader<Eval, S, V, U, 0>::run(p, nvalid, 1, pf, [&](const RdLeaf& L, int sz, int u) -> const char* {
                        const int64_t koff = nvalid[u] > 0 ? rd_kept_offset(p, (uint32_t) kos[u], L.kstride) + (int64_t) li * V * L.rstride[0] : 0;
                        return L.ptr + koff * sz;
                    });
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (it0 + u < G) {     // uniform over the warp
                            S x[V];
#pragma unroll
                            for (int v = 0; v < V; ++v) x[v] = (S) p.identity_bits;
                            if (nvalid[u] > 0) {
                                pf.u = u;
                                Eval::template run<S, V>(p.prog, pf, x);
                                Acc::template cast_in<S, V>(p, x);
#pragma unroll
                                for (int v = 0; v < V; ++v)
                                    if (v >= nvalid[u]) x[v] = (S) p.identity_bits;
                            }
                            S r[1] = {x[0]};
#pragma unroll
                            for (int v = 1; v < V; ++v) {
                                S y[1] = {x[v]};
                                Acc::template step<S, 1>(p, r, y);
                            }
                            for (int o = G >> 1; o > 0; o >>= 1) {
                                S y[1] = {shfl_xor<S>(r[0], o)};
                                Acc::template step<S, 1>(p, r, y);
                            }
                            if (li == it0 + u) keep = r[0];
                        }
                    }
                }
                done = true;
            }
        }
        for (int it = 0; it < G && !done; ++it) {
            const int64_t ko = out_base + g * G + it;
            S r[1] = {(S) p.identity_bits};
            if (ko < p.K) r[0] = rd_inner_partial<Eval, Acc, S, V, NL>(p, (uint32_t) ko, jbeg, jend, li, G);
            for (int o = G >> 1; o > 0; o >>= 1) {
                S y[1] = {shfl_xor<S>(r[0], o)};
                Acc::template step<S, 1>(p, r, y);
            }
            if (li == it) keep = r[0];
        }
        const int64_t ko = out_base + lane;
        if (ko < p.K) {
            if (p.nsplit > 1) {
                store_elem<S>(p.part_ptr + ((int64_t) blockIdx.y * p.K + ko) * dtype_size(p.acc_rt), p.acc_rt, p.acc_rt, keep);
            } else {
                const int64_t off = rd_kept_offset(p, (uint32_t) ko, p.out_kstride);
                rd_finish_store<S>(p, p.out_ptr + off * dtype_size(p.out_dtype), keep);
            }
        }
    }
}

// ---- innermost dim reduced, short rows, exact fit ---------------------------------------------
// The common "reduce the last axis" shape (cfg3 axis 2: 16.7 M rows of 16 floats): one kept dim,
// one reduced dim whose V-wide vectors exactly fill the G lanes of a group, dense output of the
// accumulator type.  Everything the general kernel decides per element is decided once on the
// host, so a warp round (32 outputs) is: G batched 128-bit loads per lane, G tiny shuffle trees,
// one coalesced 128-byte store.  32-bit index math throughout.
template <class Eval, class Acc, class S, int V>
__global__ void __launch_bounds__(256) k_reduce_rows_exact(const __grid_constant__ RdParams p) {
    constexpr int NL = Eval::kLeaves;
    constexpr int U = 4;
    const int G = p.G;
    const int lane = threadIdx.x & 31;
    const int li = lane & (G - 1);
    const int g = lane / G;
    const uint32_t K = (uint32_t) p.K;
    const uint32_t rounds = (K + 31) >> 5;
    const uint32_t warp0 = blockIdx.x * 8 + (threadIdx.x >> 5);
    constexpr int RT = Eval::kResultType;
    constexpr int osz = dtype_size(RT);
    PreFetch<NL, U, S, V> pf;
    for (uint32_t round = warp0; round < rounds; round += gridDim.x * 8) {
        const uint32_t out_base = round << 5;
        S keep = (S) p.identity_bits;
        for (int it0 = 0; it0 < G; it0 += U) {
            int nvalid[U];
            uint32_t kos[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                kos[u] = out_base + g * G + it0 + u;
                nvalid[u] = (it0 + u < G && kos[u] < K) ? V : 0;
            }
            RdLeafLoader<Eval, S, V, U, 0>::run(p, nvalid, 1, pf, [&](const RdLeaf& L, int sz, int u) -> const char* {
                const uint32_t ko = nvalid[u] > 0 ? kos[u] : 0u;
                return L.ptr + (int64_t) ((int32_t) ko * (int32_t) L.kstride[0] + (int32_t) (li * V) * (int32_t) L.rstride[0]) * sz;
            });
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (it0 + u < G) {     // uniform over the warp
                    pf.u = u;
                    S x[V];
                    Eval::template run<S, V>(p.prog, pf, x);
                    S r[1] = {x[0]};
#pragma unroll
                    for (int v = 1; v < V; ++v) {
                        S y[1] = {x[v]};
                        Acc::template step<S, 1>(p, r, y);
                    }
                    if (nvalid[u] == 0) r[0] = (S) p.identity_bits;
                    for (int o = G >> 1; o > 0; o >>= 1) {
                        S y[1] = {shfl_xor<S>(r[0], o)};
                        Acc::template step<S, 1>(p, r, y);
                    }
                    if (li == it0 + u) keep = r[0];
                }
            }
        }
        const uint32_t ko = out_base + lane;
        if (ko < K) {
            if (p.fin_op == 0) store_elem<S>(p.out_ptr + (int64_t) ko * osz, RT, RT, keep);
            else rd_store_final<S>(p, p.out_ptr + (int64_t) ko * dtype_size(p.out_dtype), keep);
        }
    }
}

// ---- innermost dim reduced, one block (256 lanes) per output ----------------------------
template <class Eval, class Acc, class S, int V>
__global__ void __launch_bounds__(256) k_reduce_inner_block(const __grid_constant__ RdParams p) {
    constexpr int NL = Eval::kLeaves;
    __shared__ S smem[8];
    const int tid = threadIdx.x;
    const int64_t jbeg = (int64_t) blockIdx.y * p.chunk;
    int64_t jend = jbeg + p.chunk;
    if (jend > p.rvec_total) jend = p.rvec_total;
    for (int64_t ko = blockIdx.x; ko < p.K; ko += gridDim.x) {
        S r[1] = {rd_inner_partial<Eval, Acc, S, V, NL>(p, (uint32_t) ko, jbeg, jend, tid, 256)};
        for (int o = 16; o > 0; o >>= 1) {
            S y[1] = {shfl_xor<S>(r[0], o)};
            Acc::template step<S, 1>(p, r, y);
        }
        __syncthreads();
        if ((tid & 31) == 0) smem[tid >> 5] = r[0];
        __syncthreads();
        if (tid < 32) {
            r[0] = tid < 8 ? smem[tid] : (S) p.identity_bits;
            for (int o = 4; o > 0; o >>= 1) {
                S y[1] = {shfl_xor<S>(r[0], o)};
                Acc::template step<S, 1>(p, r, y);
            }
            if (tid == 0) {
                if (p.nsplit > 1) {
                    store_elem<S>(p.part_ptr + ((int64_t) blockIdx.y * p.K + ko) * dtype_size(p.acc_rt), p.acc_rt, p.acc_rt, r[0]);
                } else {
                    const int64_t off = rd_kept_offset(p, (uint32_t) ko, p.out_kstride);
                    rd_finish_store<S>(p, p.out_ptr + off * dtype_size(p.out_dtype), r[0]);
                }
            }
        }
    }
}

// ---- innermost dim kept --------------------------------------------------------------------
// INV: bit mask of leaves the host found invariant along the reduced dim (k_reduce_outer_inv)
template <class Eval, class Acc, class S, int V, int INV>
XTB_DEV void reduce_outer_body(const RdParams& p) {
    constexpr int NL = Eval::kLeaves;
    const int64_t rbeg = (int64_t) blockIdx.y * p.chunk;
    int64_t rend = rbeg + p.chunk;
    if (rend > p.R) rend = p.R;
    const uint32_t KL = (uint32_t) p.kshape[p.nk - 1];
    for (int64_t t = (int64_t) blockIdx.x * 256 + threadIdx.x; t < p.kvec_total; t += (int64_t) gridDim.x * 256) {
        const uint32_t krow = fd_div((uint32_t) t, p.kvpr_div);
        const uint32_t cv = (uint32_t) t - krow * p.kvpr;
        const uint32_t col = cv * V;
        const uint32_t ko0 = krow * KL + col;      // linear kept index of the thread's first output
        RdFetch<NL> f{p, {nullptr}, V, false};
        const uint32_t rem = KL - col;
        f.nvalid = rem < (uint32_t) V ? (int) rem : V;
        const char* base[NL];
        int64_t rstep[NL];
#pragma unroll
        for (int k = 0; k < NL; ++k) {
            if (k < p.n_leaves) {
                const RdLeaf& L = p.leaf[k];
                const int sz = dtype_size(L.dtype);
                base[k] = L.ptr + rd_kept_offset(p, ko0, L.kstride) * sz;
                rstep[k] = L.rstride[p.nr - 1] * (int64_t) sz;
            }
        }
        S acc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = (S) p.identity_bits;
        if (Eval::kPrefetch && p.nr == 1) {
            if constexpr (Eval::kPrefetch) {
                // U rows of every leaf in flight before any arithmetic; rows are then accumulated in
                // the reference's order (row r before row r + 1)
                constexpr int U = 8;
                std::conditional_t<INV != 0, PreFetchInv<NL, U, S, V, INV>, PreFetch<NL, U, S, V>> pf;
                // blocked summation (two_level): `acc` holds one batch of kRdFlush rows, `total` the flushed blocks
                static_assert(U == kRdFlush, "one block per staged batch");
                RdTotal<Acc, S, V> total;
                total.init(p);
                int64_t r = rbeg;
                if (p.outer_fast) {
                    const char* row0[NL];
                    int64_t rs[NL];
#pragma unroll
                    for (int k = 0; k < NL; ++k) {
                        rs[k] = rstep[k];
                        row0[k] = base[k] + r * rs[k];
                    }
                    for (; r + U <= rend; r += U) {
                        RdFastLoader<Eval, S, V, U, 0, INV>::template run<decltype(pf), NL>(p, row0, rs, pf, r != rbeg);
#pragma unroll
                        for (int k = 0; k < NL; ++k) row0[k] += U * rs[k];
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            pf.u = u;
                            S x[V];
                            Eval::template run<S, V>(p.prog, pf, x);
                            Acc::template cast_in<S, V>(p, x);
                            Acc::template step<S, V>(p, acc, x);
                        }
                        if (p.two_level) total.flush(p, acc);
                    }
                }
                const int64_t rslow = r;   // the invariant leaves are staged by whichever loop runs first
                for (; r < rend; r += U) {
                    int nvalid[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) nvalid[u] = (r + u < rend) ? f.nvalid : 0;
                    // leaves that do not depend on the reduced index (e.g. the mean in square(a - mean))
                    // keep the registers staged by the first iteration
                    RdLeafLoader<Eval, S, V, U, 0, INV>::run(p, nvalid, 0, pf, [&](const RdLeaf& L, int sz, int u) -> const char* {
                        const int64_t koff = rd_kept_offset(p, ko0, L.kstride);
                        const int64_t ru = nvalid[u] > 0 ? r + u : rbeg;
                        return L.ptr + (koff + ru * L.rstride[0]) * sz;
                    }, r != rbeg || rslow != rbeg);
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (nvalid[u] > 0) {
                            pf.u = u;
                            S x[V];
                            Eval::template run<S, V>(p.prog, pf, x);
                            Acc::template cast_in<S, V>(p, x);
                            Acc::template step<S, V>(p, acc, x);
                        }
                    }
                    if (p.two_level) total.flush(p, acc);
                }
                if (p.two_level) total.finish(p, acc);
            }
        } else {
            // same block boundaries as the staged loop above (results do not depend on the evaluator)
            RdTotal<Acc, S, V> total;
            total.init(p);
            int in_block = 0;
            for (int64_t r = rbeg; r < rend; ++r) {
#pragma unroll
                for (int k = 0; k < NL; ++k) {
                    if (k < p.n_leaves) {
                        if (p.nr == 1) f.cur[k] = base[k] + r * rstep[k];
                        else f.cur[k] = base[k] + rd_reduced_offset(p, (uint32_t) r, p.nr, p.leaf[k].rstride) * dtype_size(p.leaf[k].dtype);
                    }
                }
                S x[V];
                Eval::template run<S, V>(p.prog, f, x);
                Acc::template cast_in<S, V>(p, x);
                Acc::template step<S, V>(p, acc, x);
                if (p.two_level && ++in_block >= kRdFlush) {
                    total.flush(p, acc);
                    in_block = 0;
                }
            }
            if (p.two_level) total.finish(p, acc);
        }
        if (p.nsplit > 1) {
            // partials[nsplit][K]: k_reduce_merge streams them with coalesced 128-bit loads
            const int asz = dtype_size(p.acc_rt);
            char* dst = p.part_ptr + ((int64_t) blockIdx.y * p.K + ko0) * asz;
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (v < f.nvalid) store_elem<S>(dst + (int64_t) v * asz, p.acc_rt, p.acc_rt, acc[v]);
        } else {
            const int osz = dtype_size(p.out_dtype);
            char* dst = p.out_ptr + rd_kept_offset(p, ko0, p.out_kstride) * osz;
            if (p.has_initial) {
                S b[V];
#pragma unroll
                for (int v = 0; v < V; ++v) b[v] = (S) p.initial_bits;
                Acc::template step<S, V>(p, acc, b);
            }
            if (p.out_vec_ok && f.nvalid == V && p.fin_op == 0) {
                store_vec<S, V>(dst, p.out_dtype, p.acc_rt, acc);
            } else {
                const int64_t step = p.out_kstride[p.nk - 1] * osz;
#pragma unroll
                for (int v = 0; v < V; ++v)
                    if (v < f.nvalid) rd_store_final<S>(p, dst + v * step, acc[v]);
            }
        }
    }
}

// compile-time one-leaf programs: 3 CTAs per SM (<= 80 registers) keep 96 KB of rows in flight per SM
template <class Eval, class Acc, class S, int V>
__global__ void __launch_bounds__(256, Eval::kPrefetch ? (Eval::kLeaves >= 2 ? 2 : 3) : 1) k_reduce_outer(const __grid_constant__ RdParams p) {
    pdl_enter();
    reduce_outer_body<Eval, Acc, S, V, 0>(p);
}
// invariant leaves (the mean in square(a - mean)) staged once; 80 registers -> 3 CTAs per SM
template <class Eval, class Acc, class S, int V, int INV>
__global__ void __launch_bounds__(256, 3) k_reduce_outer_inv(const __grid_constant__ RdParams p) {
    pdl_enter();
    reduce_outer_body<Eval, Acc, S, V, INV>(p);
}

#ifndef XTB_RTC
// launch geometry shared by the ahead-of-time and the run-time specialised kernels
struct RdLaunch {
    int kind;          // 0 rows_exact, 1 inner_warp, 2 inner_block, 3 outer
    unsigned gx, gy;
};
static inline RdLaunch reduce_geometry(const RdParams& p, const DeviceCtx* ctx, bool inner, bool allow_exact) {
    RdLaunch g{3, 1, (unsigned) p.nsplit};
    if (inner && p.exact_rows && allow_exact) {
        const int64_t rounds = (p.K + 31) / 32;
        g.kind = 0;
        g.gx = (unsigned) std::min<int64_t>((rounds + 7) / 8, (int64_t) ctx->sm_count * 64);
        g.gy = 1;
    } else if (inner && p.G <= 32) {
        const int64_t rounds = (p.K + 31) / 32;
        g.kind = 1;
        g.gx = (unsigned) std::min<int64_t>((rounds + 7) / 8, (int64_t) ctx->sm_count * 64);
    } else if (inner) {
        g.kind = 2;
        g.gx = (unsigned) std::min<int64_t>(p.K, (int64_t) ctx->sm_count * 64);
    } else {
        const int64_t blocks = (p.kvec_total + 255) / 256;
        g.kind = 3;
        g.gx = (unsigned) std::min<int64_t>(blocks, (int64_t) ctx->sm_count * 64);
    }
    return g;
}

template <class Eval, class Acc, class S, int V>
static int launch_reduce(const RdParams& p, DeviceCtx* ctx, bool inner, const char* evname) {
    char name[96];
    if constexpr (Eval::kPrefetch) {
        if (inner && p.exact_rows) {
            const int64_t rounds = (p.K + 31) / 32;
            const unsigned grid = (unsigned) std::min<int64_t>((rounds + 7) / 8, (int64_t) ctx->sm_count * 64);
            snprintf(name, sizeof(name), "k_reduce_rows_exact<%s,S%d,V%d>[G=%d]", evname, (int) sizeof(S) * 8, V, p.G);
            k_reduce_rows_exact<Eval, Acc, S, V><<<grid, 256, 0, ctx->stream>>>(p);
            note_launch(name);
            return check_launch(name);
        }
    }
    if (inner && p.G <= 32) {
        const int64_t rounds = (p.K + 31) / 32;
        dim3 grid((unsigned) std::min<int64_t>((rounds + 7) / 8, (int64_t) ctx->sm_count * 64), (unsigned) p.nsplit);
        snprintf(name, sizeof(name), "k_reduce_inner_warp<%s,S%d,V%d>[G=%d,split=%d]", evname, (int) sizeof(S) * 8, V, p.G, p.nsplit);
        k_reduce_inner_warp<Eval, Acc, S, V><<<grid, 256, 0, ctx->stream>>>(p);
    } else if (inner) {
        dim3 grid((unsigned) std::min<int64_t>(p.K, (int64_t) ctx->sm_count * 64), (unsigned) p.nsplit);
        snprintf(name, sizeof(name), "k_reduce_inner_block<%s,S%d,V%d>[split=%d]", evname, (int) sizeof(S) * 8, V, p.nsplit);
        k_reduce_inner_block<Eval, Acc, S, V><<<grid, 256, 0, ctx->stream>>>(p);
    } else {
        const int64_t blocks = (p.kvec_total + 255) / 256;
        dim3 grid((unsigned) std::min<int64_t>(blocks, (int64_t) ctx->sm_count * 64), (unsigned) p.nsplit);
        snprintf(name, sizeof(name), "k_reduce_outer<%s,S%d,V%d>[split=%d]", evname, (int) sizeof(S) * 8, V, p.nsplit);
        bool inv1 = false;
        if constexpr (Eval::kPrefetch && Eval::kLeaves == 2 && V > 1) {
            // leaf 1 staged once when it does not move along the reduced dim and every thread owns whole vectors
            const RdLeaf& L = p.leaf[1];
            inv1 = p.nr == 1 && p.n_leaves == 2 && L.rstride[0] == 0 && L.mode == MODE_VEC &&
                   p.leaf[0].rstride[0] != 0 && p.kshape[p.nk - 1] % V == 0;
            if (inv1) {
                snprintf(name, sizeof(name), "k_reduce_outer<%s,S%d,V%d,inv1>[split=%d]", evname, (int) sizeof(S) * 8, V, p.nsplit);
                launch_pdl(k_reduce_outer_inv<Eval, Acc, S, V, 2>, grid, 256, 0, ctx->stream, p);
            }
        }
        if (!inv1) launch_pdl(k_reduce_outer<Eval, Acc, S, V>, grid, 256, 0, ctx->stream, p);
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
StaticReduceTable static_reduce_table();      // V = 4 (32-bit slots) / 2 (64-bit slots)
StaticReduceTable static_reduce_table_v1();   // V = 1

#endif  // XTB_RTC

}  // namespace xtb
