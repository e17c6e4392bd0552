// xtb_ew.cuh -- the fused N-ary broadcast elementwise kernels behind xtb_assign.
//
// Replaces the three CPU loops selected by
// xexpression_assigner_base<xtensor_expression_tag>::assign_data
// (include/xtensor/core/xassign.hpp:439-478):
//   linear_assigner::run        xassign.hpp:701-849   -> k_ew<ND=1>   (contiguous, 128-bit)
//   strided_loop_assigner::run  xassign.hpp:1100-1344 -> k_ew<ND=2,3> (broadcast rows)
//   stepper_assigner::run       xassign.hpp:644-695   -> k_ew_generic (any rank / strides)
// The stepper odometer of xiterator.hpp:589-631 becomes closed-form index math:
//   addr_k(i) = base_k + sum_d i_d * stride_k[d],  stride 0 on broadcast dims.
#pragma once
#include "xtb_ops.cuh"
#ifndef XTB_RTC
#include <algorithm>
#include "xtb_common.hpp"
#endif
#include "xtb_static_programs.cuh"

namespace xtb {

// MODE_LINEAR: MODE_VEC and additionally dense over the whole collapsed space, so the
// element offset is just the linear index (no coordinate math).
// MODE_TILE (k_ew_tile only): the leaf is contiguous along another dim than the output; it is
// read in 32x32 tiles along ITS contiguous dim and transposed through shared memory.
enum LeafMode : int32_t { MODE_VEC = 0, MODE_BCAST = 1, MODE_GATHER = 2, MODE_LINEAR = 3, MODE_TILE = 4 };

struct EwLeaf {
    const char* ptr;
    int64_t stride[XTB_MAX_DIM];  // elements, per collapsed dim
    int32_t s32[4];               // the same for the rank <= 3 kernels (valid when EwParams::idx32)
    int32_t dtype;
    int32_t mode;
};

struct EwParams {
    DevProgram prog;
    int32_t ndim;
    int32_t n_leaves;
    int64_t shape[XTB_MAX_DIM];
    int64_t total_vec;      // number of V-wide vectors (rows * vec_per_row)
    uint32_t vec_per_row;
    uint32_t out_rt;        // register type of the value to store
    int32_t idx32;          // every operand offset fits in int32: rank <= 3 kernels may run
    int32_t fast;           // full vectors, full blocks, no gather operand (rank <= 3 kernels)
    // k_ew_tile: tile over (dim tile_i, inner dim); batch = all other dims
    int32_t tile_i;
    int32_t tile_slot[XTB_MAX_LEAVES];   // shared-memory slot of each MODE_TILE leaf
    uint32_t ntile_i, ntile_j;
    FastDiv div_ntj, div_nti;
    FastDiv div_vpr;
    FastDiv div_dim[XTB_MAX_DIM];
    EwLeaf leaf[XTB_MAX_LEAVES];
    EwLeaf out;
};

// Position of one thread's vector in the rank <= 3 kernels: 32-bit coordinates and offsets
// (host guarantees every operand spans < 2^31 elements).
template <int ND> struct EwFetch {
    uint32_t idx[ND > 1 ? ND - 1 : 1];  // outer coordinates
    uint32_t col;                       // first inner-dim element of this thread's vector
    uint32_t lin;                       // linear element index of that element
    int nvalid;

    XTB_DEV int32_t offset_of(const EwLeaf& L) const {
        if (L.mode == MODE_LINEAR) return (int32_t) lin;
        int32_t off = (int32_t) col * L.s32[ND - 1];
#pragma unroll
        for (int d = 0; d < ND - 1; ++d) off += (int32_t) idx[d] * L.s32[d];
        return off;
    }
};

// generic rank: run-time ndim, 64-bit coordinates
struct EwFetchN {
    const EwParams& p;
    int64_t idx[XTB_MAX_DIM];
    int64_t col;
    int nvalid;
    XTB_DEV int64_t offset_of(const EwLeaf& L) const {
        const int nd = p.ndim;
        int64_t off = col * L.stride[nd - 1];
        for (int d = 0; d < nd - 1; ++d) off += idx[d] * L.stride[d];
        return off;
    }
    template <class S, int V> XTB_DEV void load(int k, int dt, S (&x)[V]) const {
        const EwLeaf& L = p.leaf[k];
        const int sz = dtype_size(dt);
        const char* ptr = L.ptr + offset_of(L) * sz;
        if (L.mode == MODE_BCAST) {
            S s = load_elem<S>(ptr, dt);
#pragma unroll
            for (int v = 0; v < V; ++v) x[v] = s;
        } else if (L.mode == MODE_VEC && nvalid == V) {
            load_vec<S, V>(ptr, dt, x);
        } else {
            const int64_t step = L.stride[p.ndim - 1] * sz;
#pragma unroll
            for (int v = 0; v < V; ++v) x[v] = (v < nvalid) ? load_elem<S>(ptr + v * step, dt) : S(0);
        }
    }
};

struct InterpEval {
    static constexpr int kUnroll = 1;
    static constexpr int kResultType = -1;  // run-time (EwParams::out_rt)
    static constexpr int kLeaves = XTB_MAX_LEAVES;
    static constexpr bool kPrefetch = false;
    template <class S, int V, class Fetch> static XTB_DEV void run(const DevProgram& prog, Fetch& f, S (&r)[V]) {
        interpret<S, V>(prog, f, r);
    }
};
// Tbl: struct with a static constexpr array `progs` of SProg; ID indexes it.
template <class Tbl, int ID> struct StaticEval {
    static constexpr int kUnroll = 4;
    static constexpr int kResultType = sprogs::result_type(Tbl::progs[ID]);
    static constexpr int kLeaves = Tbl::progs[ID].n_leaves;
    static constexpr bool kPrefetch = true;
    template <int K> static constexpr int leaf_dtype() { return sprogs::leaf_dtype(Tbl::progs[ID], K); }
    template <class S, int V, class Fetch> static XTB_DEV void run(const DevProgram& prog, Fetch& f, S (&r)[V]) {
        eval_static<Tbl, ID, S, V>(prog.imms, f, r);
    }
};

template <class S, int V, class Fetch>
XTB_DEV void ew_store(const EwParams& p, const Fetch& f, const S (&r)[V]) {
    const EwLeaf& O = p.out;
    const int sz = dtype_size(O.dtype);
    char* ptr = (char*) O.ptr + f.offset_of(O) * sz;
    if (O.mode == MODE_VEC && f.nvalid == V) {
        store_vec<S, V>(ptr, O.dtype, (int) p.out_rt, r);
    } else {
        const int nd = p.ndim;
        const int64_t step = O.stride[nd - 1] * sz;
#pragma unroll
        for (int v = 0; v < V; ++v)
            if (v < f.nvalid) store_elem<S>(ptr + v * step, O.dtype, (int) p.out_rt, r[v]);
    }
}

// Rank <= 3 kernel for compile-time programs: the output dtype equals the program's
// result type (host checks), so the store is a plain vector store.
template <int RT, class S, int V, int ND>
XTB_DEV void ew_store_static(const EwParams& p, const EwFetch<ND>& f, const S (&r)[V]) {
    const EwLeaf& O = p.out;
    constexpr int sz = dtype_size(RT);
    char* ptr = (char*) O.ptr + (int64_t) f.offset_of(O) * sz;
    if ((O.mode == MODE_VEC || O.mode == MODE_LINEAR) && f.nvalid == V) {
        store_vec<S, V>(ptr, RT, RT, r);
    } else {
        const int64_t step = (int64_t) O.s32[ND - 1] * sz;
#pragma unroll
        for (int v = 0; v < V; ++v)
            if (v < f.nvalid) store_elem<S>(ptr + v * step, RT, RT, r[v]);
    }
}

// Stage leaf K (and, recursively, the following leaves) of a compile-time program: the
// storage dtype is a constant, so element size, address arithmetic and the load instruction
// are all resolved at compile time; only the access mode is a (uniform) run-time branch.
template <class Eval, class S, int V, int ND, int ITEMS, bool FAST, int K> struct EwLeafLoader {
    template <class PF>
    static XTB_DEV void run(const EwParams& p, const EwFetch<ND> (&f)[ITEMS], const int (&nvalid)[ITEMS], PF& pf) {
        if constexpr (K < Eval::kLeaves) {
            constexpr int dt = Eval::template leaf_dtype<K>();
            constexpr int sz = dtype_size(dt);
            const EwLeaf& L = p.leaf[K];
            const char* addr[ITEMS];
            if (L.mode == MODE_LINEAR) {
#pragma unroll
                for (int it = 0; it < ITEMS; ++it) addr[it] = L.ptr + (uint64_t) f[it].lin * (uint32_t) sz;
            } else {
#pragma unroll
                for (int it = 0; it < ITEMS; ++it) {
                    int32_t off = (int32_t) f[it].col * L.s32[ND - 1];
#pragma unroll
                    for (int d = 0; d < ND - 1; ++d) off += (int32_t) f[it].idx[d] * L.s32[d];
                    addr[it] = L.ptr + (int64_t) off * sz;
                }
            }
            if (L.mode == MODE_BCAST) {
#pragma unroll
                for (int it = 0; it < ITEMS; ++it) {
                    const S v0 = (FAST || nvalid[it] > 0) ? load_elem<S>(addr[it], dt) : S(0);
#pragma unroll
                    for (int v = 0; v < V; ++v) pf.pre[K][it][v] = v0;
                }
            } else if (FAST || L.mode != MODE_GATHER) {
                bool full = true;
                if constexpr (!FAST) {
#pragma unroll
                    for (int it = 0; it < ITEMS; ++it) full = full && nvalid[it] == V;
                }
                if (full) {
#pragma unroll
                    for (int it = 0; it < ITEMS; ++it) load_vec<S, V>(addr[it], dt, pf.pre[K][it]);
                } else {
#pragma unroll
                    for (int it = 0; it < ITEMS; ++it)
#pragma unroll
                        for (int v = 0; v < V; ++v) pf.pre[K][it][v] = (v < nvalid[it]) ? load_elem<S>(addr[it] + v * sz, dt) : S(0);
                }
            } else {
                const int64_t step = (int64_t) L.s32[ND - 1] * sz;
#pragma unroll
                for (int it = 0; it < ITEMS; ++it)
#pragma unroll
                    for (int v = 0; v < V; ++v) pf.pre[K][it][v] = (v < nvalid[it]) ? load_elem<S>(addr[it] + v * step, dt) : S(0);
            }
            EwLeafLoader<Eval, S, V, ND, ITEMS, FAST, K + 1>::run(p, f, nvalid, pf);
        }
    }
};

// vectors per thread of the rank <= 3 kernels (ahead-of-time and run-time specialised alike)
#ifndef XTB_EW_ITEMS
#define XTB_EW_ITEMS 2
#endif
constexpr int kEwItems = XTB_EW_ITEMS;

// One thread evaluates ITEMS vectors of V consecutive inner-dim elements; within a block
// consecutive threads take consecutive vectors (coalesced 128-bit access).  Compile-time
// programs only: coordinates for all items first, then the batched loads of every leaf
// (ITEMS x leaves 128-bit loads in flight per thread), then evaluation and stores.
// FAST: the host verified that every vector is full (inner % V == 0), every block is full and
// no operand needs the strided-gather path, so all tail / liveness predicates vanish.
template <class Eval, class S, int V, int ND, int ITEMS, bool FAST>
XTB_DEV void ew_body(const EwParams& p) {
    constexpr int NL = Eval::kLeaves > 0 ? Eval::kLeaves : 1;
    const uint32_t inner = (uint32_t) p.shape[ND - 1];
    const uint32_t total = (uint32_t) p.total_vec;
    const uint32_t base = blockIdx.x * (256u * ITEMS) + threadIdx.x;
    EwFetch<ND> f[ITEMS] = {};
    int nvalid[ITEMS];
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        uint32_t vec = base + it * 256u;
        bool live = true;
        if constexpr (!FAST) {
            live = vec < total;
            if (!live) vec = 0;
        }
        uint32_t cv = vec;
        if constexpr (ND > 1) {
            uint32_t row = fd_div(vec, p.div_vpr);
            cv = vec - row * p.vec_per_row;
            f[it].lin = row * inner + cv * V;
#pragma unroll
            for (int d = ND - 2; d >= 0; --d) {
                if (d == 0) {
                    f[it].idx[0] = row;
                } else {
                    uint32_t q = fd_div(row, p.div_dim[d]);
                    f[it].idx[d] = row - q * (uint32_t) p.shape[d];
                    row = q;
                }
            }
        } else {
            f[it].lin = vec * V;
        }
        f[it].col = cv * V;
        if constexpr (FAST) {
            f[it].nvalid = V;
        } else {
            const uint32_t rem = inner - f[it].col;
            f[it].nvalid = live ? (rem < (uint32_t) V ? (int) rem : V) : 0;
        }
        nvalid[it] = f[it].nvalid;
    }
    PreFetch<NL, ITEMS, S, V> pf;
    EwLeafLoader<Eval, S, V, ND, ITEMS, FAST, 0>::run(p, f, nvalid, pf);
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        if (FAST || nvalid[it] > 0) {
            pf.u = it;
            S r[V];
            Eval::template run<S, V>(p.prog, pf, r);
            ew_store_static<Eval::kResultType, S, V, ND>(p, f[it], r);
        }
    }
}

template <class Eval, class S, int V, int ND, int ITEMS>
__global__ void __launch_bounds__(256) k_ew(const __grid_constant__ EwParams p) {
    pdl_enter();
    if (p.fast) ew_body<Eval, S, V, ND, ITEMS, true>(p);
    else ew_body<Eval, S, V, ND, ITEMS, false>(p);
}

// Any rank, 64-bit indices (slow path: rank > 3 after collapsing, or >= 2^31 vectors).
template <class Eval, class S, int V>
__global__ void __launch_bounds__(256) k_ew_generic(const __grid_constant__ EwParams p) {
    const int nd = p.ndim;
    const int64_t inner = p.shape[nd - 1];
    for (int64_t vec = (int64_t) blockIdx.x * 256 + threadIdx.x; vec < p.total_vec; vec += (int64_t) gridDim.x * 256) {
        EwFetchN f{p, {0}, 0, V};
        int64_t row = vec / p.vec_per_row;
        f.col = (vec - row * p.vec_per_row) * V;
        for (int d = nd - 2; d >= 0; --d) {
            const int64_t q = row / p.shape[d];
            f.idx[d] = row - q * p.shape[d];
            row = q;
        }
        const int64_t rem = inner - f.col;
        f.nvalid = rem < V ? (int) rem : V;
        S r[V];
        Eval::template run<S, V>(p.prog, f, r);
        ew_store<S, V>(p, f, r);
    }
}


#ifndef XTB_RTC
// ---- launch ---------------------------------------------------------------------
// Rank-specialised kernels (compile-time programs): 32-bit index math, ND <= 3.
template <class Eval, class S, int V>
static int launch_ew_nd(const EwParams& p, DeviceCtx* ctx, const char* evname) {
    const int nd = p.ndim;
    constexpr int ITEMS = kEwItems;
    const int64_t per_block = 256 * ITEMS;
    const unsigned grid = (unsigned) ((p.total_vec + per_block - 1) / per_block);
    char name[96];
    EwParams q = p;
    bool fast = (p.shape[nd - 1] % V == 0) && (p.total_vec % per_block == 0) && p.out.mode != MODE_GATHER && p.out.mode != MODE_BCAST;
    for (int k = 0; k < p.n_leaves; ++k) fast = fast && p.leaf[k].mode != MODE_GATHER;
    q.fast = fast;
    snprintf(name, sizeof(name), "k_ew<%s,S%d,V%d,ND%d>%s", evname, (int) sizeof(S) * 8, V, nd, fast ? "[fast]" : "");
    switch (nd) {
        case 1: launch_pdl(k_ew<Eval, S, V, 1, ITEMS>, grid, 256, 0, ctx->stream, q); break;
        case 2: launch_pdl(k_ew<Eval, S, V, 2, ITEMS>, grid, 256, 0, ctx->stream, q); break;
        default: launch_pdl(k_ew<Eval, S, V, 3, ITEMS>, grid, 256, 0, ctx->stream, q); break;
    }
    note_launch(name);
    return check_launch(name);
}
static inline bool ew_nd_ok(const EwParams& p) { return p.ndim <= 3 && p.idx32 && p.total_vec < (int64_t) 0x7fffffff; }

// Any rank / any size.
template <class Eval, class S, int V>
static int launch_ew_generic(const EwParams& p, DeviceCtx* ctx, const char* evname) {
    const int64_t blocks = (p.total_vec + 255) / 256;
    const int64_t cap = (int64_t) ctx->sm_count * 64;
    const unsigned grid = (unsigned) std::min(blocks, cap);
    char name[96];
    snprintf(name, sizeof(name), "k_ew_generic<%s,S%d,V%d>", evname, (int) sizeof(S) * 8, V);
    k_ew_generic<Eval, S, V><<<grid, 256, 0, ctx->stream>>>(p);
    note_launch(name);
    return check_launch(name);
}

#endif  // XTB_RTC
// ---- tiled kernel: transposed leaves -------------------------------------------------------
// Replaces stepper_assigner::run (xassign.hpp:644-695) for the case the CPU handles worst: an
// operand whose fast dim is not the output's (xt::transpose, column-major leaves).  A block
// owns a 32(i) x 32(j) tile: MODE_TILE leaves are read coalesced along i into padded shared
// memory, then every thread evaluates 4 outputs coalesced along j, reading those leaves
// transposed from shared memory and all other leaves directly.
constexpr int kTile = 32;
constexpr int kMaxTileLeaves = 3;

struct TileFetch {
    const EwParams& p;
    const void* smem;
    int64_t idx[XTB_MAX_DIM];   // full coordinates of the current element
    int tx, ii;
    template <class S, int V> XTB_DEV void load(int k, int dt, S (&x)[V]) const {
        const EwLeaf& L = p.leaf[k];
        if (L.mode == MODE_TILE) {
            const S(*t)[kTile][kTile + 1] = (const S(*)[kTile][kTile + 1]) smem;
            x[0] = t[p.tile_slot[k]][tx][ii];
        } else {
            int64_t off = 0;
            for (int d = 0; d < p.ndim; ++d) off += idx[d] * L.stride[d];
            x[0] = load_elem<S>(L.ptr + off * dtype_size(dt), dt);
        }
    }
};

template <class Eval, class S>
__global__ void __launch_bounds__(256) k_ew_tile(const __grid_constant__ EwParams p) {
    __shared__ S tile[kMaxTileLeaves][kTile][kTile + 1];
    const int nd = p.ndim, di = p.tile_i, dj = nd - 1;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    // block -> (batch, tile_i, tile_j)
    uint32_t b = blockIdx.x;
    uint32_t q = fd_div(b, p.div_ntj);
    const uint32_t tj = b - q * p.ntile_j;
    b = q;
    q = fd_div(b, p.div_nti);
    const uint32_t ti = b - q * p.ntile_i;
    uint32_t batch = q;
    TileFetch f{p, tile, {0}, tx, 0};
    for (int d = nd - 2; d >= 0; --d) {
        if (d == di) continue;
        const uint32_t qq = fd_div(batch, p.div_dim[d]);
        f.idx[d] = batch - qq * (uint32_t) p.shape[d];
        batch = qq;
    }
    const int64_t Ni = p.shape[di], Nj = p.shape[dj];
    const int64_t i0 = (int64_t) ti * kTile, j0 = (int64_t) tj * kTile;
    // phase 1: transposed leaves, coalesced along i
    for (int k = 0; k < p.n_leaves; ++k) {
        const EwLeaf& L = p.leaf[k];
        if (L.mode != MODE_TILE) continue;
        const int sz = dtype_size(L.dtype);
        int64_t boff = 0;
        for (int d = 0; d < nd - 1; ++d)
            if (d != di) boff += f.idx[d] * L.stride[d];
        const int slot = p.tile_slot[k];
#pragma unroll
        for (int r = 0; r < kTile / 8; ++r) {
            const int jj = ty + 8 * r;
            const int64_t i = i0 + tx, j = j0 + jj;
            if (i < Ni && j < Nj) tile[slot][jj][tx] = load_elem<S>(L.ptr + (boff + i * L.stride[di] + j * L.stride[dj]) * sz, L.dtype);
        }
    }
    __syncthreads();
    // phase 2: evaluate, coalesced along j
    const EwLeaf& O = p.out;
    const int osz = dtype_size(O.dtype);
#pragma unroll
    for (int r = 0; r < kTile / 8; ++r) {
        const int ii = ty + 8 * r;
        const int64_t i = i0 + ii, j = j0 + tx;
        if (i < Ni && j < Nj) {
            f.ii = ii;
            f.idx[di] = i;
            f.idx[dj] = j;
            S x[1];
            Eval::template run<S, 1>(p.prog, f, x);
            int64_t off = 0;
            for (int d = 0; d < nd; ++d) off += f.idx[d] * O.stride[d];
            store_elem<S>((char*) O.ptr + off * osz, O.dtype, (int) p.out_rt, x[0]);
        }
    }
}

// Compile-time-program version: dtype / element size constant per leaf, 32-bit offsets, and all
// global loads of the block (tile leaves AND the direct leaves of the 4 outputs a thread owns)
// are issued before the barrier, so their latencies overlap.
template <class Eval, class S, int K> struct TileLeafStage {
    // phase 1: issue loads.  tile leaves -> shared memory, direct leaves -> registers
    template <class PF>
    static XTB_DEV void load(const EwParams& p, S (*tile)[kTile][kTile + 1], const int32_t (&bidx)[XTB_MAX_DIM], int32_t i0, int32_t j0,
                             int tx, int ty, PF& pf) {
        if constexpr (K < Eval::kLeaves) {
            constexpr int dt = Eval::template leaf_dtype<K>();
            constexpr int sz = dtype_size(dt);
            const EwLeaf& L = p.leaf[K];
            const int nd = p.ndim, di = p.tile_i, dj = nd - 1;
            const int32_t Ni = (int32_t) p.shape[di], Nj = (int32_t) p.shape[dj];
            int32_t boff = 0;
            for (int d = 0; d < nd - 1; ++d)
                if (d != di) boff += bidx[d] * (int32_t) L.stride[d];
            const int32_t si = (int32_t) L.stride[di], sj = (int32_t) L.stride[dj];
            if (L.mode == MODE_TILE) {
                S tmp[kTile / 8];
#pragma unroll
                for (int r = 0; r < kTile / 8; ++r) {
                    const int32_t i = i0 + tx, j = j0 + ty + 8 * r;
                    tmp[r] = (i < Ni && j < Nj) ? load_elem<S>(L.ptr + (int64_t) (boff + i * si + j * sj) * sz, dt) : S(0);
                }
                const int slot = p.tile_slot[K];
#pragma unroll
                for (int r = 0; r < kTile / 8; ++r) tile[slot][ty + 8 * r][tx] = tmp[r];
            } else {
#pragma unroll
                for (int r = 0; r < kTile / 8; ++r) {
                    const int32_t i = i0 + ty + 8 * r, j = j0 + tx;
                    pf.pre[K][r][0] = (i < Ni && j < Nj) ? load_elem<S>(L.ptr + (int64_t) (boff + i * si + j * sj) * sz, dt) : S(0);
                }
            }
            TileLeafStage<Eval, S, K + 1>::load(p, tile, bidx, i0, j0, tx, ty, pf);
        }
    }
    // phase 2: tile leaves from shared memory (transposed read)
    template <class PF>
    static XTB_DEV void gather(const EwParams& p, const S (*tile)[kTile][kTile + 1], int tx, int ty, PF& pf) {
        if constexpr (K < Eval::kLeaves) {
            if (p.leaf[K].mode == MODE_TILE) {
                const int slot = p.tile_slot[K];
#pragma unroll
                for (int r = 0; r < kTile / 8; ++r) pf.pre[K][r][0] = tile[slot][tx][ty + 8 * r];
            }
            TileLeafStage<Eval, S, K + 1>::gather(p, tile, tx, ty, pf);
        }
    }
};

template <class Eval, class S>
__global__ void __launch_bounds__(256) k_ew_tile_static(const __grid_constant__ EwParams p) {
    constexpr int NL = Eval::kLeaves > 0 ? Eval::kLeaves : 1;
    constexpr int RT = Eval::kResultType;
    __shared__ S tile[kMaxTileLeaves][kTile][kTile + 1];
    const int nd = p.ndim, di = p.tile_i, dj = nd - 1;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    uint32_t b = blockIdx.x;
    uint32_t q = fd_div(b, p.div_ntj);
    const uint32_t tj = b - q * p.ntile_j;
    b = q;
    q = fd_div(b, p.div_nti);
    const uint32_t ti = b - q * p.ntile_i;
    uint32_t batch = q;
    int32_t bidx[XTB_MAX_DIM] = {0};
    for (int d = nd - 2; d >= 0; --d) {
        if (d == di) continue;
        const uint32_t qq = fd_div(batch, p.div_dim[d]);
        bidx[d] = (int32_t) (batch - qq * (uint32_t) p.shape[d]);
        batch = qq;
    }
    const int32_t i0 = (int32_t) (ti * kTile), j0 = (int32_t) (tj * kTile);
    PreFetch<NL, kTile / 8, S, 1> pf;
    TileLeafStage<Eval, S, 0>::load(p, tile, bidx, i0, j0, tx, ty, pf);
    __syncthreads();
    TileLeafStage<Eval, S, 0>::gather(p, tile, tx, ty, pf);
    const EwLeaf& O = p.out;
    constexpr int osz = dtype_size(RT);
    const int32_t Ni = (int32_t) p.shape[di], Nj = (int32_t) p.shape[dj];
    int32_t boff = 0;
    for (int d = 0; d < nd - 1; ++d)
        if (d != di) boff += bidx[d] * (int32_t) O.stride[d];
#pragma unroll
    for (int r = 0; r < kTile / 8; ++r) {
        const int32_t i = i0 + ty + 8 * r, j = j0 + tx;
        if (i < Ni && j < Nj) {
            pf.u = r;
            S x[1];
            Eval::template run<S, 1>(p.prog, pf, x);
            store_elem<S>((char*) O.ptr + (int64_t) (boff + i * (int32_t) O.stride[di] + j * (int32_t) O.stride[dj]) * osz, RT, RT, x[0]);
        }
    }
}

#ifndef XTB_RTC
template <class Eval, class S>
static int launch_ew_tile(const EwParams& p, DeviceCtx* ctx, const char* evname) {
    int64_t batch = 1;
    for (int d = 0; d < p.ndim - 1; ++d)
        if (d != p.tile_i) batch *= p.shape[d];
    const int64_t blocks = batch * p.ntile_i * p.ntile_j;
    if (blocks >= 0x7fffffffLL) return set_error(XTB_ERR_UNSUPPORTED, "too many tiles");
    char name[96];
    if constexpr (Eval::kPrefetch) {
        if (p.idx32 && p.out.dtype == Eval::kResultType) {
            snprintf(name, sizeof(name), "k_ew_tile_static<%s,S%d>", evname, (int) sizeof(S) * 8);
            k_ew_tile_static<Eval, S><<<(unsigned) blocks, 256, 0, ctx->stream>>>(p);
            note_launch(name);
            return check_launch(name);
        }
    }
    snprintf(name, sizeof(name), "k_ew_tile<%s,S%d>", evname, (int) sizeof(S) * 8);
    k_ew_tile<Eval, S><<<(unsigned) blocks, 256, 0, ctx->stream>>>(p);
    note_launch(name);
    return check_launch(name);
}

#endif  // XTB_RTC
// ---- staged interpreter kernel -----------------------------------------------------------------
// Run-time programs cannot rely on the compiler to hoist loads out of the interpreter loop, so
// memory-level parallelism is built explicitly: every thread first issues cp.async copies of all
// the vectors it will need (leaves x ITEMS, 16 bytes each) into its private shared-memory slots,
// waits once, and only then interprets.  PUSH then reads shared memory (dynamic leaf index is
// free there), and the opcode switch runs once per V-wide vector (vec_unary / vec_binary).
constexpr int kStageItems = 2;

XTB_DEV void cp_async_16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t) __cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
XTB_DEV void cp_async_8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t) __cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
XTB_DEV void cp_async_4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t) __cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
XTB_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// thread-private staging slot of (leaf k, item it): 16 bytes holding V elements in storage dtype
struct StagedFetch {
    const EwParams& p;
    const uint4* stage;   // [n_leaves][kStageItems][256]
    int it;
    template <class S, int V> XTB_DEV void load(int k, int dt, S (&x)[V]) const {
        const char* s = (const char*) &stage[(k * kStageItems + it) * 256 + threadIdx.x];
        const int sz = dtype_size(dt);
        if (p.leaf[k].mode == MODE_BCAST) {
            const S e = load_elem<S>(s, dt);
#pragma unroll
            for (int v = 0; v < V; ++v) x[v] = e;
        } else if (sz == 4 && V == 4) {
            const uint4 r = *(const uint4*) s;
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int v = 0; v < V; ++v) x[v] = (S) w[v & 3];
        } else if (sz == 8 && V == 2 && sizeof(S) == 8) {
            const uint4 r = *(const uint4*) s;
            x[0] = (S) (((uint64_t) r.y << 32) | r.x);
            if (V > 1) x[V > 1 ? 1 : 0] = (S) (((uint64_t) r.w << 32) | r.z);
        } else {
#pragma unroll
            for (int v = 0; v < V; ++v) x[v] = load_elem<S>(s + v * sz, dt);
        }
    }
};

template <class S, int V>
__global__ void __launch_bounds__(256) k_ew_staged(const __grid_constant__ EwParams p) {
    extern __shared__ uint4 stage[];
    const int nd = p.ndim;
    const uint32_t inner = (uint32_t) p.shape[nd - 1];
    const uint32_t total = (uint32_t) p.total_vec;
    const uint32_t base = blockIdx.x * (256u * kStageItems) + threadIdx.x;
    int32_t idx[kStageItems][XTB_MAX_DIM];
    uint32_t col[kStageItems];
    int nvalid[kStageItems];
#pragma unroll
    for (int it = 0; it < kStageItems; ++it) {
        uint32_t vec = base + it * 256u;
        const bool live = vec < total;
        if (!live) vec = 0;
        uint32_t row = fd_div(vec, p.div_vpr);
        col[it] = (vec - row * p.vec_per_row) * V;
        for (int d = nd - 2; d >= 0; --d) {
            const uint32_t q = d == 0 ? 0u : fd_div(row, p.div_dim[d]);
            idx[it][d] = (int32_t) (d == 0 ? row : row - q * (uint32_t) p.shape[d]);
            row = q;
        }
        const uint32_t rem = inner - col[it];
        nvalid[it] = live ? (rem < (uint32_t) V ? (int) rem : V) : 0;
    }
    // phase 1: asynchronous copies of every operand vector into this thread's slots
    for (int k = 0; k < p.n_leaves; ++k) {
        const EwLeaf& L = p.leaf[k];
        const int sz = dtype_size(L.dtype);
#pragma unroll
        for (int it = 0; it < kStageItems; ++it) {
            if (nvalid[it] == 0) continue;
            int32_t off = (int32_t) col[it] * (int32_t) L.stride[nd - 1];
            for (int d = 0; d < nd - 1; ++d) off += idx[it][d] * (int32_t) L.stride[d];
            const char* g = L.ptr + (int64_t) off * sz;
            char* sdst = (char*) &stage[(k * kStageItems + it) * 256 + threadIdx.x];
            const int bytes = V * sz;
            if (L.mode == MODE_VEC && nvalid[it] == V && bytes >= 4) {
                if (bytes == 16) cp_async_16(sdst, g);
                else if (bytes == 8) cp_async_8(sdst, g);
                else if (bytes == 4) cp_async_4(sdst, g);
                else { cp_async_16(sdst, g); cp_async_16(sdst + 16, g + 16); }
            } else if (L.mode == MODE_BCAST) {
                if (sz == 4) cp_async_4(sdst, g);
                else if (sz == 8) cp_async_8(sdst, g);
                else if (sz == 2) *(uint16_t*) sdst = *(const uint16_t*) g;
                else *(uint8_t*) sdst = *(const uint8_t*) g;
            } else {
                // tails, misaligned rows, strided gathers: element by element
                const int64_t step = (L.mode == MODE_VEC ? 1 : L.stride[nd - 1]) * sz;
                for (int v = 0; v < nvalid[it]; ++v) {
                    const char* ge = g + v * step;
                    char* se = sdst + v * sz;
                    if (sz == 4) cp_async_4(se, ge);
                    else if (sz == 8) cp_async_8(se, ge);
                    else if (sz == 2) *(uint16_t*) se = *(const uint16_t*) ge;
                    else *(uint8_t*) se = *(const uint8_t*) ge;
                }
            }
        }
    }
    cp_async_wait_all();
    // phase 2: interpret
#pragma unroll
    for (int it = 0; it < kStageItems; ++it) {
        if (nvalid[it] == 0) continue;
        StagedFetch f{p, stage, it};
        S r[V];
        interpret<S, V>(p.prog, f, r);
        const EwLeaf& O = p.out;
        const int osz = dtype_size(O.dtype);
        int32_t off = (int32_t) col[it] * (int32_t) O.stride[nd - 1];
        for (int d = 0; d < nd - 1; ++d) off += idx[it][d] * (int32_t) O.stride[d];
        char* ptr = (char*) O.ptr + (int64_t) off * osz;
        if (O.mode == MODE_VEC && nvalid[it] == V) {
            store_vec<S, V>(ptr, O.dtype, (int) p.out_rt, r);
        } else {
            const int64_t step = O.stride[nd - 1] * osz;
            for (int v = 0; v < nvalid[it]; ++v) store_elem<S>(ptr + v * step, O.dtype, (int) p.out_rt, r[v]);
        }
    }
}

#ifndef XTB_RTC
template <class S, int V>
static int launch_ew_staged(const EwParams& p, DeviceCtx* ctx) {
    const int64_t per_block = 256 * kStageItems;
    const unsigned grid = (unsigned) ((p.total_vec + per_block - 1) / per_block);
    const size_t smem = (size_t) std::max(p.n_leaves, 1) * kStageItems * 256 * sizeof(uint4);
    // per launch, not once per process: the attribute belongs to the current device's context
    cudaFuncSetAttribute(k_ew_staged<S, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, XTB_MAX_LEAVES * kStageItems * 256 * (int) sizeof(uint4));
    char name[96];
    snprintf(name, sizeof(name), "k_ew_staged<interp,S%d,V%d>", (int) sizeof(S) * 8, V);
    k_ew_staged<S, V><<<grid, 256, smem, ctx->stream>>>(p);
    note_launch(name);
    return check_launch(name);
}

#endif  // XTB_RTC
#ifndef XTB_RTC
// ---- registry of compile-time programs ---------------------------------------------
// Programs whose instruction stream equals a pre-instantiated SProg run a fully
// unrolled kernel; anything else runs the interpreter.  Both paths share every
// line of functor code (xtb_ops.cuh).
struct StaticEntry {
    const sprogs::SP* prog;
    const char* name;
    int (*launch_ew)(const EwParams&, DeviceCtx*);
    int (*launch_tile)(const EwParams&, DeviceCtx*);
};
struct StaticTable {
    const StaticEntry* entries;
    int n;
};
StaticTable static_table_f32();
StaticTable static_table_f64();
StaticTable static_table_i32();

inline bool sprog_matches(const sprogs::SP& sp, const xtb_program* prog) {
    if (prog->n_insns != sp.n) return false;
    for (int i = 0; i < sp.n; ++i) {
        const xtb_insn in = prog->insns[i];
        if (in.op != sp.ins[i].op || in.type != sp.ins[i].type || in.src != sp.ins[i].src || in.arg != sp.ins[i].arg)
            return false;
    }
    return true;
}
inline const StaticEntry* find_static(const xtb_program* prog) {
    const StaticTable tables[3] = {static_table_f32(), static_table_f64(), static_table_i32()};
    for (const StaticTable& t : tables)
        for (int i = 0; i < t.n; ++i)
            if (sprog_matches(*t.entries[i].prog, prog)) return &t.entries[i];
    return nullptr;
}

#endif  // XTB_RTC

}  // namespace xtb
