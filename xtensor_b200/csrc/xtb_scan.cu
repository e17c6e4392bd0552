// xtb_scan.cu -- xtb_scan: inclusive scan along an axis (xt::cumsum / xt::cumprod).
//
// Replaces detail::accumulator_impl (include/xtensor/reducers/xaccumulator.hpp:215-341):
// the reference copies the input into a dense result (promoting the value type,
// :224-234) and scans it in place with a serial loop along the axis (:282-294), or
// over the flattened row-major traversal when no axis is given (:299-341).
//
//   k_scan_lookback : scan along a contiguous run (last axis, or the flat scan).
//                     Single pass, decoupled look-back (Merrill & Garland): a tile of
//                     2048 elements publishes its aggregate, then a warp looks back over
//                     its predecessors' aggregates / inclusive prefixes.  Tile prefixes are
//                     always accumulated in ascending tile order, so floating-point results
//                     do not depend on timing (run-to-run deterministic).
//   k_scan_columns  : scan along a strided axis: one thread per (outer, inner) column walks
//                     the axis in the reference's order (bit-exact, also for floating point),
//                     coalesced across the inner dims.
#include <algorithm>
#include <cstdlib>
#include <cuda.h>              // CUtensorMap (the driver entry point is looked up at run time, libcuda is not linked)
#include <cudaTypedefs.h>
#include "xtb_common.hpp"
#include "xtb_ops.cuh"

namespace xtb {

constexpr int kScanThreads = 256;

struct ScanParams {
    const char* in;
    char* out;
    int32_t in_dtype;
    int32_t op;                   // XTB_RED_SUM / XTB_RED_PROD
    int64_t n;                    // scan length
    int64_t rows;                 // independent contiguous scans (lookback) / outer count (columns)
    int64_t inner;                // columns kernel: product of dims after the axis
    int64_t in_axis_stride;       // elements
    // outer (and, for the columns kernel, inner) coordinates -> input offset
    int32_t n_outer;
    int64_t outer_shape[XTB_MAX_DIM];
    int64_t outer_stride[XTB_MAX_DIM];
    FastDiv outer_div[XTB_MAX_DIM];
    int32_t n_inner;
    int64_t inner_shape[XTB_MAX_DIM];
    int64_t inner_stride[XTB_MAX_DIM];
    FastDiv inner_div[XTB_MAX_DIM];
    // lookback state
    uint32_t tiles_per_row;
    uint32_t total_tiles;
    char* aggregate;              // {flag, value} slot per tile
    char* prefix;                 // {flag, value} slot per block of kScanWindow tiles: the block total
    int32_t vec_io;               // input is the accumulator dtype, unit stride, 16-byte aligned rows (in and out)
    int32_t packed;               // short dense rows: a warp's window holds 128 / seg_vecs whole rows
    int32_t seg_vecs;             // 16-byte vectors per row (power of two <= 128) when packed
    int32_t st_elems;             // k_scan_stile: elements per super-tile
};

template <class T> XTB_DEV T load_cast(const char* p, int dt) {
    switch (dt) {
        case XTB_BOOL: return (T) (*(const uint8_t*) p != 0);
        case XTB_I8: return (T) * (const int8_t*) p;
        case XTB_U8: return (T) * (const uint8_t*) p;
        case XTB_I16: return (T) * (const int16_t*) p;
        case XTB_U16: return (T) * (const uint16_t*) p;
        case XTB_I32: return (T) * (const int32_t*) p;
        case XTB_U32: return (T) * (const uint32_t*) p;
        case XTB_I64: return (T) * (const long long*) p;
        case XTB_U64: return (T) * (const unsigned long long*) p;
        case XTB_F32: return (T) * (const float*) p;
        default: return (T) * (const double*) p;
    }
}

template <class T> XTB_DEV T scan_op(int op, T a, T b) { return op == XTB_RED_PROD ? (T) (a * b) : (T) (a + b); }
template <class T> XTB_DEV T scan_identity(int op) { return op == XTB_RED_PROD ? T(1) : T(0); }
template <class T> XTB_DEV T shfl_xor_t(T v, int m);
// compile-time operator (the bandwidth-critical kernels are instantiated per operator: no select per element)
template <int OP, class T> XTB_DEV T sop(T a, T b) {
    if constexpr (OP == XTB_RED_PROD) return (T) (a * b);
    else return (T) (a + b);
}
template <int OP, class T> XTB_DEV constexpr T sident() { return OP == XTB_RED_PROD ? T(1) : T(0); }
template <int OP, class T> XTB_DEV T warp_total_c(T v) {
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) v = sop<OP, T>(v, shfl_xor_t<T>(v, m));
    return v;
}

XTB_DEV int64_t scan_offset(uint32_t lin, int n, const int64_t* shape, const int64_t* stride, const FastDiv* div) {
    int64_t off = 0;
    for (int d = n - 1; d > 0; --d) {
        const uint32_t q = fd_div(lin, div[d]);
        off += (int64_t) (lin - q * (uint32_t) shape[d]) * stride[d];
        lin = q;
    }
    return n > 0 ? off + (int64_t) lin * stride[0] : 0;
}

template <class T> XTB_DEV T shfl_up_t(T v, int d) {
    if constexpr (sizeof(T) == 8) {
        unsigned long long u;
        memcpy(&u, &v, 8);
        u = __shfl_up_sync(0xffffffffu, u, d);
        T r;
        memcpy(&r, &u, 8);
        return r;
    } else {
        unsigned u;
        memcpy(&u, &v, 4);
        u = __shfl_up_sync(0xffffffffu, u, d);
        T r;
        memcpy(&r, &u, 4);
        return r;
    }
}

// streaming store of N bytes (4 / 8 / 16) from registers.  No "memory" clobber: the kernels that use it never
// read back what they store, and the compiler stays free to hoist the next loads above the store.
template <int N> XTB_DEV void memcpy_stream(void* dst, const void* src) {
    if constexpr (N == 16) {
        uint4 v;
        memcpy(&v, src, 16);
        asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
    } else if constexpr (N == 8) {
        uint2 v;
        memcpy(&v, src, 8);
        asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(dst), "r"(v.x), "r"(v.y));
    } else {
        static_assert(N == 4, "");
        uint32_t v;
        memcpy(&v, src, 4);
        asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(dst), "r"(v));
    }
}

// ---- contiguous scan: single pass, two-level decoupled look-back ------------------------------
// Tile = 512 threads x 64 bytes.  Inside a tile the data stays in the *striped* arrangement it is
// loaded in (warp w owns 32*ITEMS consecutive elements; its q-th 128-bit access covers vectors
// q*32 + lane), so every global access is a fully coalesced 512-byte warp transaction; the scan is
// done on that arrangement with shuffles (vector-local scan, one warp scan per q, carry over q).
//
// Across tiles: tiles of a row are grouped in blocks of kScanWindow.  A tile publishes its aggregate
// as soon as it is known; the last tile of a block publishes the block total.  The exclusive prefix of
// tile t in block k is
//        tree(block totals 0..k-1)  (+)  tree(aggregates of the tiles of block k before t)
// where tree() is a fixed-shape reduction (per-lane ascending partial sums, then a butterfly), so a
// floating-point result depends only on the tile's position, never on timing: run-to-run
// deterministic, no serial chain (dependency depth 2), and one read + one write of the data.
// Tiles are processed in launch order (blockIdx.x), so everything a tile waits for is already running.
constexpr int kScanWindow = 256;   // tiles per block = 32 lanes x 8
constexpr int kScanWarpsPerTile = kScanThreads / 32;

template <class T> struct ScanTile {
    static constexpr int VEC = 16 / (int) sizeof(T);
    static constexpr int NV = 4;
    static constexpr int ITEMS = VEC * NV;
    static constexpr int WARP_ELEMS = 32 * ITEMS;
    static constexpr int TILE = kScanThreads * ITEMS;
    static constexpr int SLOT = 2 * (int) sizeof(T);   // {flag, value}: 8 or 16 bytes, one memory transaction
};

// {flag, value} slots: written and read with ONE 64- / 128-bit access, so a reader that sees the flag
// sees the value (no fence); 0 = empty (the state is zeroed before the launch)
template <class T> XTB_DEV void slot_publish(char* slots, uint32_t i, T v) {
    if constexpr (sizeof(T) == 4) {
        uint32_t bits;
        memcpy(&bits, &v, 4);
        const unsigned long long w = (1ull << 32) | bits;
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(slots + (size_t) i * 8), "l"(w) : "memory");
    } else {
        unsigned long long bits;
        memcpy(&bits, &v, 8);
        asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(slots + (size_t) i * 16), "l"(1ull), "l"(bits) : "memory");
    }
}
template <class T> XTB_DEV bool slot_try(const char* slots, uint32_t i, T& v) {
    if constexpr (sizeof(T) == 4) {
        unsigned long long w;
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(slots + (size_t) i * 8) : "memory");
        const uint32_t bits = (uint32_t) w;
        memcpy(&v, &bits, 4);
        return (w >> 32) != 0ull;
    } else {
        unsigned long long f, bits;
        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(f), "=l"(bits) : "l"(slots + (size_t) i * 16) : "memory");
        memcpy(&v, &bits, 8);
        return f != 0ull;
    }
}

template <class T> XTB_DEV T shfl_idx_t(T v, int src) {
    if constexpr (sizeof(T) == 8) {
        unsigned long long u;
        memcpy(&u, &v, 8);
        u = __shfl_sync(0xffffffffu, u, src);
        T r;
        memcpy(&r, &u, 8);
        return r;
    } else {
        unsigned u;
        memcpy(&u, &v, 4);
        u = __shfl_sync(0xffffffffu, u, src);
        T r;
        memcpy(&r, &u, 4);
        return r;
    }
}
template <class T> XTB_DEV T shfl_xor_t(T v, int m) {
    if constexpr (sizeof(T) == 8) {
        unsigned long long u;
        memcpy(&u, &v, 8);
        u = __shfl_xor_sync(0xffffffffu, u, m);
        T r;
        memcpy(&r, &u, 8);
        return r;
    } else {
        unsigned u;
        memcpy(&u, &v, 4);
        u = __shfl_xor_sync(0xffffffffu, u, m);
        T r;
        memcpy(&r, &u, 4);
        return r;
    }
}
// butterfly total (the operators are commutative, so every lane ends with the same bits)
template <class T> XTB_DEV T warp_total(int op, T v) {
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) v = scan_op<T>(op, v, shfl_xor_t<T>(v, m));
    return v;
}

// Executed by one whole warp of a tile: gathers what precedes the tile in its row.  Returns (in every
// lane) the exclusive prefix of the tile; *block_part = the part contributed by the tile's own block
// (what the last tile of a block adds its aggregate to when it publishes the block total).
// The tile's own aggregate is not needed here, so this can run while the tile's data is still in flight.
template <class T>
XTB_DEV T tile_lookback(const ScanParams& p, int op, uint32_t row, uint32_t trow, int lane, T* block_part) {
    const T ident = scan_identity<T>(op);
    const uint32_t blocks_per_row = (p.tiles_per_row + kScanWindow - 1) / kScanWindow;
    const uint32_t kb = trow / kScanWindow;            // complete blocks before mine
    const uint32_t cnt = trow - kb * kScanWindow;      // tiles of my block before me
    const uint32_t first = row * p.tiles_per_row + kb * kScanWindow;
    // aggregates of my block: lane l owns tiles l, l+32, .. (ascending); all loads (and the first block
    // total) are issued before the first one is inspected
    T v[8];
    bool ok[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t idx = (uint32_t) j * 32 + lane;
        v[j] = ident;
        ok[j] = idx >= cnt || slot_try<T>(p.aggregate, first + idx, v[j]);
    }
    T bv = ident;
    bool bok = (uint32_t) lane >= kb || slot_try<T>(p.prefix, row * blocks_per_row + lane, bv);
    T a = ident;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t idx = (uint32_t) j * 32 + lane;
        while (!ok[j]) {
            __nanosleep(40);
            ok[j] = slot_try<T>(p.aggregate, first + idx, v[j]);
        }
        if (idx < cnt) a = scan_op<T>(op, a, v[j]);
    }
    a = warp_total<T>(op, a);
    *block_part = a;
    T e = a;
    if (kb > 0) {
        T b = ident;
        for (uint32_t m = lane; m < kb; m += 32) {
            if (m >= 32) bok = slot_try<T>(p.prefix, row * blocks_per_row + m, bv);
            while (!bok) {
                __nanosleep(40);
                bok = slot_try<T>(p.prefix, row * blocks_per_row + m, bv);
            }
            b = scan_op<T>(op, b, bv);
        }
        b = warp_total<T>(op, b);
        e = cnt > 0 ? scan_op<T>(op, b, a) : b;
    }
    return e;
}

// The window-local half of tile_lookback: tree(aggregates of the tiles of my window before me), in every lane.
// Same shape as in tile_lookback, so "part (+) my aggregate" is the same block total whoever publishes it.
template <class T>
XTB_DEV T window_part(const ScanParams& p, int op, uint32_t row, uint32_t trow, int lane) {
    const T ident = scan_identity<T>(op);
    const uint32_t kb = trow / kScanWindow;
    const uint32_t cnt = trow - kb * kScanWindow;
    const uint32_t first = row * p.tiles_per_row + kb * kScanWindow;
    T v[8];
    bool ok[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t idx = (uint32_t) j * 32 + lane;
        v[j] = ident;
        ok[j] = idx >= cnt || slot_try<T>(p.aggregate, first + idx, v[j]);
    }
    T a = ident;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t idx = (uint32_t) j * 32 + lane;
        while (!ok[j]) {
            __nanosleep(40);
            ok[j] = slot_try<T>(p.aggregate, first + idx, v[j]);
        }
        if (idx < cnt) a = scan_op<T>(op, a, v[j]);
    }
    return warp_total<T>(op, a);
}

// WARP_ROWS: rows of at most 32*ITEMS elements, one warp per row (no block or tile combine)
template <class T, bool WARP_ROWS>
__global__ void __launch_bounds__(kScanThreads, 4) k_scan_tiles(const __grid_constant__ ScanParams p) {
    using C = ScanTile<T>;
    constexpr int VEC = C::VEC, NV = C::NV;
    __shared__ T s_warp[kScanWarpsPerTile];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int op = p.op;
    const T ident = scan_identity<T>(op);
    const uint32_t tile = blockIdx.x;
    uint32_t row, trow;
    int64_t wbase;                       // first element (within the row) of this warp's chunk
    int64_t limit = p.n;                 // elements addressable from in_row / out_row
    int seg = 128;                       // vectors per scan segment inside the warp's window
    const int isz = dtype_size(p.in_dtype);
    const char* in_row;
    T* out_row;
    if constexpr (WARP_ROWS) {
        const int64_t r = (int64_t) tile * kScanWarpsPerTile + warp;
        trow = 0;
        if (p.packed) {
            // the array is one dense run of short rows: windows of 128 vectors over the flat data
            wbase = r * C::WARP_ELEMS;
            limit = p.rows * p.n;
            if (wbase >= limit) return;
            seg = p.seg_vecs;
            row = 0;
            in_row = p.in;
            out_row = (T*) p.out;
        } else {
            if (r >= p.rows) return;
            row = (uint32_t) r;
            wbase = 0;
            in_row = p.in + scan_offset(row, p.n_outer, p.outer_shape, p.outer_stride, p.outer_div) * isz;
            out_row = (T*) p.out + (int64_t) row * p.n;
        }
    } else {
        row = tile / p.tiles_per_row;
        trow = tile - row * p.tiles_per_row;
        wbase = (int64_t) trow * C::TILE + (int64_t) warp * C::WARP_ELEMS;
        in_row = p.in + scan_offset(row, p.n_outer, p.outer_shape, p.outer_stride, p.outer_div) * isz;
        out_row = (T*) p.out + (int64_t) row * p.n;
    }

    // ---- load (striped) ----
    T x[NV][VEC];
    if (p.vec_io) {
        // rows are 16-byte aligned; a vector is loaded whole when it lies inside the row
        const char* src = in_row + wbase * (int64_t) sizeof(T);
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int vi = q * 32 + lane;
            const int64_t j0 = wbase + (int64_t) vi * VEC;
            if (j0 + VEC <= limit) {
                const uint4 r = ldg_stream_16(src + (size_t) vi * 16);
                memcpy(&x[q][0], &r, 16);
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) x[q][i] = j0 + i < limit ? ((const T*) in_row)[j0 + i] : ident;
            }
        }
    } else {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const int64_t j = wbase + (int64_t) (q * 32 + lane) * VEC + i;
                x[q][i] = j < limit ? load_cast<T>(in_row + j * p.in_axis_stride * isz, p.in_dtype) : ident;
            }
        }
    }
    // ---- warp-level scan on the striped arrangement ----
    T inc[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) {
#pragma unroll
        for (int i = 1; i < VEC; ++i) x[q][i] = scan_op<T>(op, x[q][i - 1], x[q][i]);
        inc[q] = x[q][VEC - 1];
    }
    const int segl = seg < 32 ? seg : 32;            // lanes per segment within one q-row
    const int lis = lane & (segl - 1);               // lane position inside its segment
    const int qper = seg >= 32 ? seg >> 5 : 1;       // q-rows per segment
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const T y = shfl_up_t<T>(inc[q], d);
            if (lis >= d) inc[q] = scan_op<T>(op, y, inc[q]);
        }
    }
    T off[NV];                           // exclusive offset of vector (q, lane) within its segment
    T carry = ident;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const T ex = shfl_up_t<T>(inc[q], 1);
        const T rowtot = shfl_idx_t<T>(inc[q], 31);
        const bool seg_start = (q & (qper - 1)) == 0;
        const T c = seg_start ? ident : carry;
        off[q] = lis == 0 ? c : scan_op<T>(op, c, ex);
        carry = seg_start ? rowtot : scan_op<T>(op, carry, rowtot);
    }
    T base_off = ident;                  // everything before this warp's chunk
    if constexpr (!WARP_ROWS) {
        // ---- block-level: offsets of the warps, tile aggregate ----
        if (lane == 0) s_warp[warp] = carry;
        __syncthreads();
        T wv = lane < kScanWarpsPerTile ? s_warp[lane] : ident;
#pragma unroll
        for (int d = 1; d < kScanWarpsPerTile; d <<= 1) {
            const T y = shfl_up_t<T>(wv, d);
            if (lane >= d) wv = scan_op<T>(op, y, wv);
        }
        const T tile_total = shfl_idx_t<T>(wv, kScanWarpsPerTile - 1);
        const T warp_off = shfl_idx_t<T>(wv, warp > 0 ? warp - 1 : 0);
        base_off = warp > 0 ? warp_off : ident;
    }
    // ---- apply offsets, store (striped) ----
    const bool first_warp_of_row = (WARP_ROWS && !p.packed) || (!WARP_ROWS && trow == 0 && warp == 0);
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const T o = first_warp_of_row ? off[q] : scan_op<T>(op, base_off, off[q]);
        if (!(first_warp_of_row && q == 0 && lane == 0)) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) x[q][i] = scan_op<T>(op, o, x[q][i]);
        }
    }
    if (p.vec_io) {
        char* dst = (char*) (out_row + wbase);
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int vi = q * 32 + lane;
            const int64_t j0 = wbase + (int64_t) vi * VEC;
            if (j0 + VEC <= limit) {
                uint4 r;
                memcpy(&r, &x[q][0], 16);
                stg_stream_16(dst + (size_t) vi * 16, r);
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i)
                    if (j0 + i < limit) out_row[j0 + i] = x[q][i];
            }
        }
    } else {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const int64_t j = wbase + (int64_t) (q * 32 + lane) * VEC + i;
                if (j < limit) out_row[j] = x[q][i];
            }
        }
    }
}

// ---- contiguous scan, rows longer than one tile: reduce ahead, scan from L2 --------------------------
// A single-pass look-back scan (Merrill & Garland; CUB's DeviceScan) retires tiles in order: a tile cannot
// store before every older tile has LOADED, so under HBM3e latencies its registers / shared memory idle while
// the slowest older load completes (measured on this B200: CUB InclusiveSum and the round-1 staged look-back
// kernel both stop at 0.58 of the copy bandwidth).  Here no tile ever waits for another one:
//   CTA k  (1) loads tile k from HBM, reduces it and publishes its aggregate            ("reduce ahead"),
//          (2) loads tile k - D, which CTA k - D pulled through the 126 MB L2 a few microseconds ago, scans it
//              with the prefix gathered from aggregates that were published D tiles earlier, and stores it,
//          (3) (one warp) folds complete groups of 32 units of the aggregate tree that finished D / 2 tiles ago.
// Both tile loads are issued before anything is consumed.  HBM traffic stays one read + one write per element
// (D x tile = 24 MB of look-ahead lives in L2); the L2 serves one extra read.  Tiles are taken in launch order.
// Across tiles a fan-32 tree: level 0 holds every tile's aggregate, a unit of level l + 1 the total of 32 units
// of level l.  The exclusive prefix of tile t is
//        part[L-1] (+) ... (+) part[1] (+) part[0],   part[l] = butterfly(totals of the siblings before t's unit at level l)
// which depends only on t's position: floating-point results are run-to-run deterministic.
// Every slot is one {flag, value} word (one 64- / 128-bit access: no fences); a reader still checks the flag.
constexpr int kLbFan = 32;
constexpr int kLbMaxLevels = 5;

struct LbTree {
    int32_t levels;
    uint32_t ahead;                     // D: tiles between the reduce visit and the scan visit (multiple of 64)
    uint32_t units[kLbMaxLevels];       // units per row at each level = ceil(tiles_per_row / 32^l)
    uint32_t level_off[kLbMaxLevels];   // first slot of each level; slot = level_off[l] + row * units[l] + unit
    char* slots;                        // zeroed before the launch
};

XTB_DEV uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
XTB_DEV void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}

template <class T> XTB_DEV T slot_wait(const char* slots, uint32_t i) {
    T v;
    while (!slot_try<T>(slots, i, v)) __nanosleep(20);
    return v;
}

template <class T, int NV, int OP, int TH>
__global__ void __launch_bounds__(TH, (NV <= 4 ? 5 : 3) * (256 / TH)) k_scan_ahead(const __grid_constant__ ScanParams p, const __grid_constant__ LbTree tr) {
    constexpr int VEC = 16 / (int) sizeof(T);
    constexpr int WARP_ELEMS = 32 * NV * VEC;
    constexpr int TILE = TH * NV * VEC;
    constexpr T ident = sident<OP, T>();
    __shared__ __align__(128) T s_tile[TILE];              // the reduce-visit tile lands here (bulk copy: no registers in flight)
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ T s_red[(TH / 32)];
    __shared__ T s_warp[(TH / 32)];
    __shared__ T s_prefix;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t k = blockIdx.x;
    const bool do_red = k < p.total_tiles;                 // visit 1: reduce tile k
    const bool do_scan = k >= tr.ahead;                    // visit 2: scan tile k - D
    const uint32_t j = k - tr.ahead;
    const int isz = dtype_size(p.in_dtype);
    const bool one_row = p.rows == 1;                      // the flat scan: no row arithmetic at all

    auto tile_pos = [&](uint32_t tile, uint32_t& row, uint32_t& trow) {
        if (one_row) {
            row = 0;
            trow = tile;
        } else {
            row = tile / p.tiles_per_row;
            trow = tile - row * p.tiles_per_row;
        }
    };
    auto row_ptr = [&](uint32_t row) -> const char* {
        return one_row ? p.in : p.in + scan_offset(row, p.n_outer, p.outer_shape, p.outer_stride, p.outer_div) * isz;
    };
    auto load_tile = [&](const char* in_row, uint32_t trow, T (&x)[NV][VEC]) {
        const int64_t wbase = (int64_t) trow * TILE + (int64_t) (warp * WARP_ELEMS);
        if (p.vec_io) {
            const char* src = in_row + (wbase + lane * VEC) * (int64_t) sizeof(T);
            if (wbase + WARP_ELEMS <= p.n) {               // whole warp chunk inside the row: no per-vector checks
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    const uint4 r = ldg_stream_16(src + q * 512);
                    memcpy(&x[q][0], &r, 16);
                }
            } else {
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    const int64_t j0 = wbase + (int64_t) (q * 32 + lane) * VEC;
                    if (j0 + VEC <= p.n) {
                        const uint4 r = ldg_stream_16(src + q * 512);
                        memcpy(&x[q][0], &r, 16);
                    } else {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) x[q][i] = j0 + i < p.n ? ((const T*) in_row)[j0 + i] : ident;
                    }
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < NV; ++q) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const int64_t e = wbase + (int64_t) (q * 32 + lane) * VEC + i;
                    x[q][i] = e < p.n ? load_cast<T>(in_row + e * p.in_axis_stride * isz, p.in_dtype) : ident;
                }
            }
        }
    };

    // ---- (1a) request tile k: one bulk copy HBM -> shared memory when the tile is whole and copyable ----
    uint32_t krow = 0, ktrow = 0, jrow = 0, jtrow = 0;
    bool red_bulk = false;
    const char* krow_ptr = nullptr;
    if (do_red) {
        tile_pos(k, krow, ktrow);
        krow_ptr = row_ptr(krow);
        red_bulk = p.vec_io && (int64_t) (ktrow + 1) * TILE <= p.n;
        if (red_bulk && tid == 0) {
            const uint32_t bar = smem_u32(&s_bar);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t) (TILE * sizeof(T))) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(s_tile)), "l"(krow_ptr + (int64_t) ktrow * TILE * (int64_t) sizeof(T)), "r"((uint32_t) (TILE * sizeof(T))), "r"(bar)
                         : "memory");
        }
    }
    // ---- (2) tile j = k - D: requested from L2 into registers ----
    T x[NV][VEC];
    if (do_scan) {
        tile_pos(j, jrow, jtrow);
        load_tile(row_ptr(jrow), jtrow, x);
    }
    // ---- (3) tree maintenance for tile m = k - D / 2: everything it reads was published D / 2 tiles ago ----
    if (warp == (TH / 32) - 1 && k >= tr.ahead / 2 && k - tr.ahead / 2 < p.total_tiles) {
        uint32_t row, trow;
        tile_pos(k - tr.ahead / 2, row, trow);
        if ((trow + 1) % kLbFan == 0) {                    // (most tiles stop here)
            uint32_t span = kLbFan;
            for (int l = 0; l + 1 < tr.levels; ++l) {
                // m closes a complete level-(l+1) unit that is not the last one of its row?
                if ((trow + 1) % span != 0 || trow + 1 >= p.tiles_per_row) break;
                const uint32_t unit = trow / span;                               // index of that unit at level l + 1
                const T v = slot_wait<T>(tr.slots, tr.level_off[l] + row * tr.units[l] + unit * kLbFan + lane);
                const T total = warp_total_c<OP, T>(v);
                if (lane == 0) slot_publish<T>(tr.slots, tr.level_off[l + 1] + row * tr.units[l + 1] + unit, total);
                __syncwarp();
                span *= kLbFan;
            }
        }
    }
    // ---- (2a) prefix of tile j, gathered by warp 0 while the tile loads are in flight: the slots of all levels are
    //      requested together (one L2 round trip), then combined from the highest level down ----
    if (warp == 0 && do_scan && jtrow > 0) {
        T v[kLbMaxLevels];
        bool need[kLbMaxLevels], ok[kLbMaxLevels];
        uint32_t idx[kLbMaxLevels];
        uint32_t u = jtrow;
#pragma unroll
        for (int l = 0; l < kLbMaxLevels; ++l) {
            const uint32_t cnt = u % kLbFan;
            need[l] = l < tr.levels && (uint32_t) lane < cnt;
            idx[l] = l < tr.levels ? tr.level_off[l] + jrow * tr.units[l] + (u - cnt) + lane : 0u;
            v[l] = ident;
            ok[l] = !need[l] || slot_try<T>(tr.slots, idx[l], v[l]);
            u /= kLbFan;
        }
        T acc = ident;
#pragma unroll
        for (int l = kLbMaxLevels - 1; l >= 0; --l) {
            if (l < tr.levels) {                           // warp-uniform
                while (!ok[l]) {
                    __nanosleep(20);
                    ok[l] = slot_try<T>(tr.slots, idx[l], v[l]);
                }
                const T part = warp_total_c<OP, T>(need[l] ? v[l] : ident);
                acc = sop<OP, T>(acc, part);               // higher levels lie before lower ones; ident (+) x == x
            }
        }
        if (lane == 0) s_prefix = acc;
    }
    // ---- (2b) warp-level scan of tile j on the striped arrangement ----
    T off[NV];
    if (do_scan) {
        T inc[NV];
#pragma unroll
        for (int q = 0; q < NV; ++q) {
#pragma unroll
            for (int i = 1; i < VEC; ++i) x[q][i] = sop<OP, T>(x[q][i - 1], x[q][i]);
            inc[q] = x[q][VEC - 1];
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const T y = shfl_up_t<T>(inc[q], d);
                if (lane >= d) inc[q] = sop<OP, T>(y, inc[q]);
            }
        }
        T carry = ident;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const T ex = shfl_up_t<T>(inc[q], 1);
            const T rowtot = shfl_idx_t<T>(inc[q], 31);
            off[q] = lane == 0 ? carry : sop<OP, T>(carry, ex);
            carry = q == 0 ? rowtot : sop<OP, T>(carry, rowtot);
        }
        if (lane == 0) s_warp[warp] = carry;
    }
    __syncthreads();                                       // s_warp, s_prefix; also: the mbarrier is initialised
    if (do_scan) {
        // ---- (2c) offsets of the warps, apply, store (striped): leaves before tile k has even arrived ----
        T wv = lane < (TH / 32) ? s_warp[lane] : ident;
#pragma unroll
        for (int d = 1; d < (TH / 32); d <<= 1) {
            const T y = shfl_up_t<T>(wv, d);
            if (lane >= d) wv = sop<OP, T>(y, wv);
        }
        const T warp_off = shfl_idx_t<T>(wv, warp > 0 ? warp - 1 : 0);
        T base_off = warp > 0 ? warp_off : ident;
        bool have_base = warp > 0;
        if (jtrow > 0) {
            const T e = s_prefix;
            base_off = warp > 0 ? sop<OP, T>(e, warp_off) : e;
            have_base = true;
        }
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const bool first_vec = !have_base && q == 0 && lane == 0;       // the very first vector of a row keeps its values
            if (!first_vec) {
                const T o = have_base ? ((q == 0 && lane == 0) ? base_off : sop<OP, T>(base_off, off[q])) : off[q];
#pragma unroll
                for (int i = 0; i < VEC; ++i) x[q][i] = sop<OP, T>(o, x[q][i]);
            }
        }
        const int64_t wbase = (int64_t) jtrow * TILE + (int64_t) (warp * WARP_ELEMS);
        T* out_row = (T*) p.out + (int64_t) jrow * p.n;
        if (p.vec_io && wbase + WARP_ELEMS <= p.n) {
            char* dst = (char*) (out_row + wbase + lane * VEC);
#pragma unroll
            for (int q = 0; q < NV; ++q) memcpy_stream<16>(dst + q * 512, &x[q][0]);
        } else {
#pragma unroll
            for (int q = 0; q < NV; ++q) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const int64_t e = wbase + (int64_t) (q * 32 + lane) * VEC + i;
                    if (e < p.n) out_row[e] = x[q][i];
                }
            }
        }
    }
    if (!do_red) return;
    // ---- (1b) reduce tile k: fixed order inside the thread, butterfly over the warp, fixed tree over the warps ----
    T xa[NV][VEC];
    if (red_bulk) {
        mbar_wait(smem_u32(&s_bar), 0);
        const T* mine = s_tile + warp * WARP_ELEMS + lane * VEC;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const uint4 r = *(const uint4*) (mine + q * 32 * VEC);
            memcpy(&xa[q][0], &r, 16);
        }
    } else {
        load_tile(krow_ptr, ktrow, xa);
    }
    T t = ident;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        T v = xa[q][0];
#pragma unroll
        for (int i = 1; i < VEC; ++i) v = sop<OP, T>(v, xa[q][i]);
        t = q == 0 ? v : sop<OP, T>(t, v);
    }
    t = warp_total_c<OP, T>(t);
    if (lane == 0) s_red[warp] = t;
    __syncthreads();
    if (warp == 0) {
        T v = lane < (TH / 32) ? s_red[lane] : ident;
#pragma unroll
        for (int mk = 1; mk < (TH / 32); mk <<= 1) v = sop<OP, T>(v, shfl_xor_t<T>(v, mk));
        if (lane == 0 && ktrow + 1 < p.tiles_per_row) slot_publish<T>(tr.slots, tr.level_off[0] + krow * tr.units[0] + ktrow, v);
    }
}

// ---- contiguous scan, long rows: shared-memory staged super-tiles -------------------------------
// At HBM3e latencies the bytes a register-resident tile keeps in flight (64 B per thread) cannot cover
// the look-back wait.  Here a CTA owns a super-tile of up to 64 KB: each warp fetches its slice with one
// bulk asynchronous copy (cp.async.bulk -> mbarrier; no registers held while the data is in flight, 3
// CTAs = 192 KB in flight per SM), scans it in place in shared memory as soon as it lands, and the CTA
// then does ONE look-back for the whole super-tile before streaming the result out.  Inputs that cannot
// be bulk-copied (other dtype, strided, unaligned) are read through registers into the same pipeline.
constexpr int kStMaxBytesLimit = 64 * 1024;


template <class T, int kStThreads, int kStCtas, bool CHAINED>
__global__ void __launch_bounds__(kStThreads + (CHAINED ? 32 : 0), kStCtas) k_scan_stile(const __grid_constant__ ScanParams p) {
    constexpr int kStWarps = kStThreads / 32;               // scan warps; a chained tile has one more warp: the look-back warp
    constexpr int VEC = 16 / (int) sizeof(T), NV = 4, SUB = 128 * VEC;     // elements per warp sub-chunk
    extern __shared__ __align__(128) unsigned char st_smem[];
    T* sm = (T*) st_smem;
    __shared__ __align__(8) unsigned long long s_bar[kStWarps];
    __shared__ T s_warp[kStWarps];
    __shared__ T s_prefix;
    __shared__ T s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int op = p.op;
    const T ident = scan_identity<T>(op);
    constexpr bool chained = CHAINED;
    // Tiles are taken in launch order: CTAs of a 1-D grid are dispatched by ascending blockIdx.x (the
    // assumption CUB's decoupled look-back makes too), so every predecessor of a tile is running or done.
    const uint32_t tile = blockIdx.x;
    if (lane == 0 && warp < kStWarps) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[warp])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const uint32_t row = tile / p.tiles_per_row;
    const uint32_t trow = tile - row * p.tiles_per_row;
    if (chained && warp == kStWarps) {
        // ---- the look-back warp: gathers the predecessors while the tile's data is still in flight ----
        T part;
        const T e = tile_lookback<T>(p, op, row, trow, lane, &part);
        asm volatile("bar.sync 2, 64;" ::: "memory");          // scan warp 0 has stored the tile aggregate
        if (lane == 0) {
            const uint32_t kb = trow / kScanWindow;
            if (trow - kb * kScanWindow == kScanWindow - 1 && trow + 1 < p.tiles_per_row) {
                const uint32_t blocks_per_row = (p.tiles_per_row + kScanWindow - 1) / kScanWindow;
                slot_publish<T>(p.prefix, row * blocks_per_row + kb, scan_op<T>(op, part, *(volatile T*) &s_total));
            }
            s_prefix = e;
        }
        __syncthreads();
        return;
    }
    const int tile_elems = p.st_elems;
    const int chunk = tile_elems / kStWarps;                 // elements per warp, multiple of SUB
    const int64_t tbase = (int64_t) trow * tile_elems;
    const int isz = dtype_size(p.in_dtype);
    const char* in_row = p.in + scan_offset(row, p.n_outer, p.outer_shape, p.outer_stride, p.outer_div) * isz;
    T* out_row = (T*) p.out + (int64_t) row * p.n;
    const int64_t left = p.n - tbase - (int64_t) warp * chunk;          // valid elements from my chunk on
    const int cvalid = left <= 0 ? 0 : (left < chunk ? (int) left : chunk);
    T* my = sm + warp * chunk;
    const bool bulk = p.vec_io && ((size_t) cvalid * sizeof(T)) % 16 == 0;   // warp-uniform
    if (bulk && cvalid > 0) {
        if (lane == 0) {
            const uint32_t bytes = (uint32_t) cvalid * (uint32_t) sizeof(T);
            const uint32_t bar = smem_u32(&s_bar[warp]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(my)), "l"(in_row + (tbase + (int64_t) warp * chunk) * (int64_t) sizeof(T)), "r"(bytes), "r"(bar)
                         : "memory");
        }
        // wait for my slice (phase 0 of my barrier)
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(smem_u32(&s_bar[warp])) : "memory");
        }
    }
    // ---- phase A: scan my slice in place (sub-chunks of 128 vectors, striped over the lanes) ----
    T carry = ident;
    const int nsub = chunk / SUB;
    for (int sc = 0; sc < nsub; ++sc) {
        const int sbase = sc * SUB;                           // within my chunk
        if (sbase >= cvalid) break;                           // warp-uniform
        T x[NV][VEC];
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int e0 = sbase + (q * 32 + lane) * VEC;
            if (bulk) {
                if (e0 + VEC <= cvalid) {
                    const uint4 r = *(const uint4*) (my + e0);
                    memcpy(&x[q][0], &r, 16);
                } else {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) x[q][i] = ident;   // cvalid is a multiple of VEC here
                }
            } else {
                const int64_t g0 = tbase + (int64_t) warp * chunk + e0;
#pragma unroll
                for (int i = 0; i < VEC; ++i)
                    x[q][i] = e0 + i < cvalid ? load_cast<T>(in_row + (g0 + i) * p.in_axis_stride * isz, p.in_dtype) : ident;
            }
        }
        T inc[NV];
#pragma unroll
        for (int q = 0; q < NV; ++q) {
#pragma unroll
            for (int i = 1; i < VEC; ++i) x[q][i] = scan_op<T>(op, x[q][i - 1], x[q][i]);
            inc[q] = x[q][VEC - 1];
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const T y = shfl_up_t<T>(inc[q], d);
                if (lane >= d) inc[q] = scan_op<T>(op, y, inc[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const T ex = shfl_up_t<T>(inc[q], 1);
            const T rowtot = shfl_idx_t<T>(inc[q], 31);
            const T o = lane == 0 ? carry : scan_op<T>(op, carry, ex);
#pragma unroll
            for (int i = 0; i < VEC; ++i) x[q][i] = scan_op<T>(op, o, x[q][i]);
            carry = scan_op<T>(op, carry, rowtot);
            uint4 r;
            memcpy(&r, &x[q][0], 16);
            *(uint4*) (my + sbase + (q * 32 + lane) * VEC) = r;
        }
    }
    // ---- offsets of the warps, tile aggregate, look-back ----
    if (lane == 0) s_warp[warp] = carry;
    asm volatile("bar.sync 1, %0;" ::"n"(kStThreads) : "memory");      // the scan warps
    T wv = lane < kStWarps ? s_warp[lane] : ident;
#pragma unroll
    for (int d = 1; d < kStWarps; d <<= 1) {
        const T y = shfl_up_t<T>(wv, d);
        if (lane >= d) wv = scan_op<T>(op, y, wv);
    }
    const T tile_total = shfl_idx_t<T>(wv, kStWarps - 1);
    const T warp_off = shfl_idx_t<T>(wv, warp > 0 ? warp - 1 : 0);
    T base_off = warp > 0 ? warp_off : ident;
    bool have_off = warp > 0;
    if (chained) {
        if (warp == 0) {
            // publish the aggregate at once; successors never wait for this tile's own look-back
            if (lane == 0) {
                if (trow + 1 < p.tiles_per_row) slot_publish<T>(p.aggregate, tile, tile_total);
                s_total = tile_total;
            }
            __syncwarp();
            asm volatile("bar.arrive 2, 64;" ::: "memory");
        }
        __syncthreads();                                       // the look-back warp has stored s_prefix
        if (trow > 0) {
            base_off = warp > 0 ? scan_op<T>(op, s_prefix, warp_off) : s_prefix;
            have_off = true;
        }
    }
    // ---- phase B: add the offset, stream out ----
    T* dst = out_row + tbase + (int64_t) warp * chunk;
    for (int sc = 0; sc < nsub; ++sc) {
        const int sbase = sc * SUB;
        if (sbase >= cvalid) break;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int e0 = sbase + (q * 32 + lane) * VEC;
            if (e0 >= cvalid) continue;
            T v[VEC];
            const uint4 r = *(const uint4*) (my + e0);
            memcpy(&v[0], &r, 16);
            if (have_off) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) v[i] = scan_op<T>(op, base_off, v[i]);
            }
            if (p.vec_io && e0 + VEC <= cvalid) {
                uint4 w;
                memcpy(&w, &v[0], 16);
                stg_stream_16(dst + e0, w);
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i)
                    if (e0 + i < cvalid) dst[e0 + i] = v[i];
            }
        }
    }
}

// ---- strided axis, long axis: single-pass column tiles -------------------------------------------
// A CTA owns R rows x W <= 256 columns (<= 64 KB) of one (outer, column-strip): warp 0 fetches the rows
// with bulk asynchronous copies, every scan thread walks one column of the tile downwards in shared
// memory (row groups when the strip is narrow), and a dedicated look-back warp gathers -- while the tile
// is still in flight -- the column totals of the tiles above it through a fixed fan-16 tree
// (tile -> group of 16 -> group of 256 ...), so one read + one write of the data and a result that does
// not depend on timing.  One tile along the axis and >= 256 columns: exactly the reference's order.
constexpr int kCtThreads = 256;
constexpr int kCtFan = 16;
constexpr int kCtMaxLevels = 4;
constexpr int kCtTileBytes = 64 * 1024;

struct ColTileParams {
    int32_t W, R, G;                   // columns / rows per tile, row groups per tile (= 256 / W)
    int32_t strips, chunks, levels;    // tiles across the columns / along the axis, tree levels
    int32_t bulk;                      // tile rows can be fetched with cp.async.bulk
    int32_t tma;                       // the whole tile is one TMA tensor copy (tensor map passed alongside)
    int64_t units[kCtMaxLevels];       // units along the axis at each tree level: ceil(chunks / 16^l)
    int64_t level_off[kCtMaxLevels];   // first slot of each level
    char* vals;                        // [slot][256] column totals, each a {flag, value} word (zeroed before the launch)
};

// gather the totals of the `cnt` (< 16) siblings that precede this tile at one tree level, for one column
template <class T> struct ColGather {
    T v[kCtFan - 1];
    uint32_t pending;          // bit j: sibling j not yet seen
    XTB_DEV void issue(const char* vals, int64_t first, uint32_t cnt, int col) {
        pending = 0;
#pragma unroll
        for (int j = 0; j < kCtFan - 1; ++j) {
            v[j] = T(0);
            if ((uint32_t) j < cnt && !slot_try<T>(vals, (uint32_t) ((first + j) * 256 + col), v[j])) pending |= 1u << j;
        }
    }
    XTB_DEV T finish(int op, const char* vals, int64_t first, uint32_t cnt, int col) {
        while (pending) {
            __nanosleep(64);
#pragma unroll
            for (int j = 0; j < kCtFan - 1; ++j)
                if ((pending >> j) & 1u)
                    if (slot_try<T>(vals, (uint32_t) ((first + j) * 256 + col), v[j])) pending &= ~(1u << j);
        }
        T acc = scan_identity<T>(op);
#pragma unroll
        for (int j = 0; j < kCtFan - 1; ++j)
            if ((uint32_t) j < cnt) acc = j == 0 ? v[0] : scan_op<T>(op, acc, v[j]);
        return acc;
    }
};

template <class T, bool BULK>
__global__ void __launch_bounds__(kCtThreads, 3) k_scan_coltile(const __grid_constant__ ScanParams p, const __grid_constant__ ColTileParams c,
                                                                 const __grid_constant__ CUtensorMap tmap) {
    constexpr int CV = BULK ? 16 / (int) sizeof(T) : 1;   // adjacent columns per scan thread (one 128-bit access)
    extern __shared__ __align__(128) unsigned char ct_smem[];
    T* sm = (T*) ct_smem;                                  // [R][W]
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ __align__(16) T s_gt[kCtThreads * CV];      // totals of the row groups [G][W]
    __shared__ T s_tot[kCtThreads];                        // column totals of the tile
    __shared__ T s_off[kCtThreads];                        // exclusive prefix of the tile per column
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int op = p.op;
    const T ident = scan_identity<T>(op);
    const bool chained = c.chunks > 1;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // Tiles are taken in launch order (CTAs of a 1-D grid are dispatched by ascending blockIdx.x, the
    // same assumption CUB's decoupled look-back makes): strip-fastest, then along the axis, then over
    // the outer index, so whatever a tile waits for was dispatched at least `strips` CTAs earlier.
    const uint32_t tile = blockIdx.x;
    const uint32_t per_o = (uint32_t) c.chunks * (uint32_t) c.strips;
    const uint32_t o = tile / per_o;
    const uint32_t rem = tile - o * per_o;
    const uint32_t chunk = rem / (uint32_t) c.strips;
    const uint32_t strip = rem - chunk * (uint32_t) c.strips;
    const int W = c.W, R = c.R;
    const int64_t col0 = (int64_t) strip * 256;
    const int wv = (int) (p.inner - col0 < W ? p.inner - col0 : W);        // valid columns
    const int64_t r0 = (int64_t) chunk * R;
    const int rvalid = (int) (p.n - r0 < R ? p.n - r0 : R);
    const int isz = dtype_size(p.in_dtype);
    const char* in_o = p.in + scan_offset(o, p.n_outer, p.outer_shape, p.outer_stride, p.outer_div) * isz;

    // ---- fetch the tile ----
    if constexpr (BULK) {
        if (c.tma) {
            // one TMA tensor copy for the whole R x W box (rows / columns outside the array are zero-filled
            // and count as transferred)
            if (tid == 0) {
                const uint32_t bar = smem_u32(&s_bar);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t) (R * W * (int) sizeof(T))) : "memory");
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             ::"r"(smem_u32(sm)), "l"(&tmap), "r"((int) col0), "r"((int) r0), "r"((int) o), "r"(bar) : "memory");
            }
        } else if (warp == 0) {
            const uint32_t row_bytes = (uint32_t) wv * (uint32_t) sizeof(T);
            const uint32_t bar = smem_u32(&s_bar);
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes * (uint32_t) rvalid) : "memory");
            __syncwarp();
            for (int r = lane; r < rvalid; r += 32) {
                const char* src = in_o + ((r0 + r) * p.in_axis_stride + col0) * (int64_t) sizeof(T);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(sm + (size_t) r * W)), "l"(src), "r"(row_bytes), "r"(bar) : "memory");
            }
        }
    }
    // ---- look-back, overlapped with the fetch: thread `tid` owns column `tid` of the strip ----
    // Every published column total is a {flag, value} slot of its own (one 64-/128-bit access): no fences.
    // Levels >= 1 were published long ago; the level-0 siblings (the tiles right above) may still be in
    // flight, so their first request goes out now and what is missing is picked up after phase A.
    const bool lb = chained && tid < wv;
    const int64_t lane_base = (int64_t) o * c.strips + strip;
    // the first requests for the two lowest levels go out now; nothing is waited for before this tile
    // has published its own totals (no tile's publication may depend on another tile's look-back)
    ColGather<T> g0, g1;
    g0.pending = g1.pending = 0;
    const uint32_t cnt0 = chunk % kCtFan, u1 = chunk / kCtFan, cnt1 = u1 % kCtFan;
    const int64_t first0 = c.level_off[0] + lane_base * c.units[0] + (chunk - cnt0);
    const int64_t first1 = c.levels > 1 ? c.level_off[1] + lane_base * c.units[1] + (u1 - cnt1) : 0;
    if (lb) {
        if (c.levels > 1) g1.issue(c.vals, first1, cnt1, tid);
        g0.issue(c.vals, first0, cnt0, tid);
    }
    if constexpr (BULK) {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(smem_u32(&s_bar)) : "memory");
        }
    }
    // ---- phase A: thread (g, cv) walks rows [ra, rb) of its CV adjacent columns downwards, in place ----
    const int wt = W / CV;                                 // threads across a tile row
    const int g = tid / wt, col = (tid - g * wt) * CV;
    const bool active = g < c.G && col < wv;               // wv is a multiple of CV
    const int rg = (R + c.G - 1) / c.G;
    const int ra = g * rg;
    const int rb = ra + rg < rvalid ? ra + rg : rvalid;
    T acc[CV];
#pragma unroll
    for (int k = 0; k < CV; ++k) acc[k] = ident;
    if (active && ra < rb) {
        T* q = sm + col;
        if constexpr (BULK) {
            uint4 v0 = *(const uint4*) (q + (size_t) ra * W);
            memcpy(&acc[0], &v0, 16);
#pragma unroll 4
            for (int r = ra + 1; r < rb; ++r) {
                T x[CV];
                uint4 v = *(const uint4*) (q + (size_t) r * W);
                memcpy(&x[0], &v, 16);
#pragma unroll
                for (int k = 0; k < CV; ++k) acc[k] = scan_op<T>(op, acc[k], x[k]);
                memcpy(&v, &acc[0], 16);
                *(uint4*) (q + (size_t) r * W) = v;
            }
        } else {
            const char* src = in_o + scan_offset((uint32_t) (col0 + col), p.n_inner, p.inner_shape, p.inner_stride, p.inner_div) * isz;
            const int64_t step = p.in_axis_stride * isz;
            bool have = false;
            for (int r = ra; r < rb; r += 8) {
                T x[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) x[k] = r + k < rb ? load_cast<T>(src + (r0 + r + k) * step, p.in_dtype) : ident;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (r + k < rb) {
                        acc[0] = have ? scan_op<T>(op, acc[0], x[k]) : x[k];
                        have = true;
                        q[(size_t) (r + k) * W] = acc[0];
                    }
                }
            }
        }
    }
    if (tid < c.G * wt) {
#pragma unroll
        for (int k = 0; k < CV; ++k) s_gt[g * W + col + k] = acc[k];
    }
    __syncthreads();
    // offset of my row group inside the tile; column totals
    T go[CV];
#pragma unroll
    for (int k = 0; k < CV; ++k) go[k] = ident;
    if (active) {
        for (int gg = 0; gg < g; ++gg) {
#pragma unroll
            for (int k = 0; k < CV; ++k) go[k] = gg == 0 ? s_gt[col + k] : scan_op<T>(op, go[k], s_gt[gg * W + col + k]);
        }
    }
    if (chained) {
        const bool more = chunk + 1 < (uint32_t) c.chunks;          // something follows this tile
        if (active && g == 0) {
            const int64_t slot = c.level_off[0] + lane_base * c.units[0] + chunk;
#pragma unroll
            for (int k = 0; k < CV; ++k) {
                T tot = s_gt[col + k];
                for (int gg = 1; gg < c.G; ++gg) tot = scan_op<T>(op, tot, s_gt[gg * W + col + k]);
                s_tot[col + k] = tot;
                if (more) slot_publish<T>(c.vals, (uint32_t) (slot * 256 + col + k), tot);
            }
        }
        __syncthreads();                                            // s_tot complete
        if (lb) {
            // walk up the tree: part[l] = the siblings before me at level l; right after part[l] is known
            // the total of my level-(l+1) unit can be published if I am its last chunk
            T below = ident;                                        // part[l-1] + ... + part[0] (position order)
            T low = ident;                                          // everything above the tile
            uint32_t u = chunk, span = 1;
#pragma unroll
            for (int l = 0; l < kCtMaxLevels; ++l) {
                if (l < c.levels) {
                    T part;
                    if (l == 0) {
                        part = g0.finish(op, c.vals, first0, cnt0, tid);
                    } else if (l == 1) {
                        part = g1.finish(op, c.vals, first1, cnt1, tid);
                    } else {
                        const uint32_t cnt = u % kCtFan;
                        const int64_t first = c.level_off[l] + lane_base * c.units[l] + (u - cnt);
                        ColGather<T> gl;
                        gl.issue(c.vals, first, cnt, tid);
                        part = gl.finish(op, c.vals, first, cnt, tid);
                    }
                    below = l == 0 ? part : scan_op<T>(op, part, below);
                    low = below;
                    span *= kCtFan;
                    u /= kCtFan;
                    if (l + 1 < c.levels && more && (chunk + 1) % span == 0) {
                        const int64_t slot = c.level_off[l + 1] + lane_base * c.units[l + 1] + chunk / span;
                        slot_publish<T>(c.vals, (uint32_t) (slot * 256 + tid), scan_op<T>(op, below, s_tot[tid]));
                    }
                }
            }
            s_off[tid] = low;
        }
        __syncthreads();
    }
    // ---- phase B: add what lies above (other tiles, earlier row groups), stream out ----
    if (active && ra < rb) {
        const bool from_tiles = chained && chunk > 0;
        const bool have_off = from_tiles || g > 0;
        T off[CV];
#pragma unroll
        for (int k = 0; k < CV; ++k) {
            off[k] = go[k];
            if (from_tiles) off[k] = g > 0 ? scan_op<T>(op, s_off[col + k], go[k]) : s_off[col + k];
        }
        T* dst = (T*) p.out + ((int64_t) o * p.n + r0) * p.inner + col0 + col;
        const T* q = sm + col;
#pragma unroll 4
        for (int r = ra; r < rb; ++r) {
            if constexpr (BULK) {
                T x[CV];
                uint4 v = *(const uint4*) (q + (size_t) r * W);
                memcpy(&x[0], &v, 16);
                if (have_off) {
#pragma unroll
                    for (int k = 0; k < CV; ++k) x[k] = scan_op<T>(op, off[k], x[k]);
                }
                memcpy(&v, &x[0], 16);
                stg_stream_16(dst + (int64_t) r * p.inner, v);
            } else {
                T v = q[(size_t) r * W];
                if (have_off) v = scan_op<T>(op, off[0], v);
                dst[(int64_t) r * p.inner] = v;
            }
        }
    }
}

// ---- strided axis, many columns: column walkers over a TMA ring ---------------------------------------
// With about one strip of columns per SM there is enough parallelism ACROSS the columns, so nothing has to be
// chained along the axis: a walker (one CTA: a producer warp and a consumer warp) owns a strip of 32 x CV
// columns and walks the whole axis top to bottom.  The producer keeps `stages` TMA boxes of 64 rows in flight
// in a shared-memory ring (cp.async.bulk.tensor -> mbarrier).  The consumer scans a box IN PLACE in shared
// memory -- per row one LDS, CV adds into the running column totals it holds in registers, one STS, all with
// immediate offsets -- and hands the box back to the TMA unit, which writes it to the output
// (cp.async.bulk.tensor global <- shared): no global address arithmetic and no per-row global stores in the
// instruction stream, rows and columns outside the array are clipped by the tensor maps.  No look-back, no
// second pass, and every column is accumulated in exactly the reference's order (accumulator_impl,
// xaccumulator.hpp:282-294): bit-exact for floating point too.  One read + one write of the data.
constexpr int kCwRowsMax = 64;       // rows per TMA box: 32 (64 is kept for experiments)

struct ColWalkParams {
    int32_t W;           // columns per walker (32 * CV)
    int32_t strips;      // walkers across the columns
    int32_t stages;      // ring depth
    int32_t chunks;      // boxes along the axis = ceil(n / box rows)
    int32_t box_rows;
};

template <int N> struct VecOf;
template <> struct VecOf<4> { using type = uint32_t; };
template <> struct VecOf<8> { using type = uint2; };
template <> struct VecOf<16> { using type = uint4; };

template <class T, int CV, int OP, int kCwRows>
__global__ void __launch_bounds__(64) k_scan_colwalk(const __grid_constant__ ScanParams p, const __grid_constant__ ColWalkParams c,
                                                      const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_out) {
    using Vec = typename VecOf<(int) sizeof(T) * CV>::type;
    extern __shared__ __align__(128) unsigned char cw_smem[];
    __shared__ __align__(8) unsigned long long s_full[16], s_empty[16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int W = 32 * CV;
    constexpr uint32_t stage_bytes = (uint32_t) (kCwRows * W * (int) sizeof(T));
    const uint32_t walker = blockIdx.x;
    const uint32_t o = walker / (uint32_t) c.strips;
    const uint32_t strip = walker - o * (uint32_t) c.strips;
    const int col0 = (int) (strip * W);
    if (threadIdx.x == 0) {
        for (int s = 0; s < c.stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_empty[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        // ---- producer: one elected lane keeps the ring full ----
        if (lane == 0) {
            int s = 0;
            uint32_t round = 0;
            for (int it = 0; it < c.chunks; ++it) {
                if (round > 0) mbar_wait(smem_u32(&s_empty[s]), (round + 1) & 1u);
                const uint32_t bar = smem_u32(&s_full[s]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(stage_bytes) : "memory");
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             ::"r"(smem_u32(cw_smem + (size_t) s * stage_bytes)), "l"(&tmap_in), "r"(col0), "r"(it * kCwRows), "r"((int) o), "r"(bar)
                             : "memory");
                if (++s == c.stages) {
                    s = 0;
                    ++round;
                }
            }
        }
        return;
    }
    // ---- consumer: lane owns CV adjacent columns of the box ----
    T acc[CV];
#pragma unroll
    for (int k = 0; k < CV; ++k) acc[k] = sident<OP, T>();
    constexpr int kLag = 2;          // boxes a stage stays with its TMA store before it is handed back
    int s = 0;
    uint32_t round = 0;
    for (int it = 0; it < c.chunks; ++it) {
        mbar_wait(smem_u32(&s_full[s]), round & 1u);
        Vec* q = (Vec*) (cw_smem + (size_t) s * stage_bytes) + lane;      // row r: q[r * 32]
        if (it > 0) {
            // rows past the end of the axis were zero-filled by the load and are clipped by the store
#pragma unroll
            for (int r = 0; r < kCwRows; ++r) {
                T x[CV];
                const Vec xv = q[r * 32];
                memcpy(&x[0], &xv, sizeof(Vec));
#pragma unroll
                for (int k = 0; k < CV; ++k) acc[k] = sop<OP, T>(acc[k], x[k]);
                Vec ov;
                memcpy(&ov, &acc[0], sizeof(Vec));
                q[r * 32] = ov;
            }
        } else {
            {   // out[0] = in[0] (a leading -0.0 survives), and the row stays as it is
                const Vec xv = q[0];
                memcpy(&acc[0], &xv, sizeof(Vec));
            }
#pragma unroll
            for (int r = 1; r < kCwRows; ++r) {
                T x[CV];
                const Vec xv = q[r * 32];
                memcpy(&x[0], &xv, sizeof(Vec));
#pragma unroll
                for (int k = 0; k < CV; ++k) acc[k] = sop<OP, T>(acc[k], x[k]);
                Vec ov;
                memcpy(&ov, &acc[0], sizeof(Vec));
                q[r * 32] = ov;
            }
        }
        // the box goes back to the TMA unit: writes of all lanes -> async proxy, then one lane issues the store
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                         ::"l"(&tmap_out), "r"(col0), "r"(it * kCwRows), "r"((int) o), "r"(smem_u32(cw_smem + (size_t) s * stage_bytes))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // the store issued kLag boxes ago has read its shared memory by now (no stall in the common case):
            // that stage may be refilled
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kLag) : "memory");
            if (it >= kLag) {
                int sr = s - kLag;
                if (sr < 0) sr += c.stages;
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_empty[sr])) : "memory");
            }
        }
        if (++s == c.stages) {
            s = 0;
            ++round;
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <class T> static int scan_coltile(ScanParams p, DeviceCtx* ctx) {
    ColTileParams c;
    memset(&c, 0, sizeof(c));
    c.W = (int32_t) std::min<int64_t>(p.inner, 256);
    c.strips = (int32_t) ((p.inner + 255) / 256);
    int64_t R = kCtTileBytes / ((int64_t) c.W * (int64_t) sizeof(T));
    if (R > p.n) R = p.n;
    c.R = (int32_t) R;
    c.chunks = (int32_t) ((p.n + R - 1) / R);
    const int64_t tiles = (int64_t) c.chunks * c.strips * p.rows;
    if (tiles >= 0x7fffffffLL) XTB_FAIL(XTB_ERR_UNSUPPORTED, "scan: too many tiles");
    // bulk copies need the tile rows to be dense, 16-byte aligned runs of the accumulator dtype
    const int64_t asz = sizeof(T);
    bool bulk = std::is_same<T, float>::value ? p.in_dtype == XTB_F32
              : std::is_same<T, double>::value ? p.in_dtype == XTB_F64
              : sizeof(T) == 4 ? (p.in_dtype == XTB_I32 || p.in_dtype == XTB_U32)      // same bits modulo 2^32
                               : (p.in_dtype == XTB_I64 || p.in_dtype == XTB_U64);
    bulk = bulk && (uintptr_t) p.in % 16 == 0 && (p.inner * asz) % 16 == 0 && (p.in_axis_stride * asz) % 16 == 0;
    int64_t expect = 1;
    for (int d = p.n_inner - 1; d >= 0 && bulk; --d) {
        if (p.inner_shape[d] != 1 && p.inner_stride[d] != expect) bulk = false;
        expect *= p.inner_shape[d];
    }
    for (int d = 0; d < p.n_outer && bulk; ++d) bulk = (p.outer_stride[d] * asz) % 16 == 0;
    c.bulk = bulk ? 1 : 0;
    // the output is dense, but its rows are 16-byte aligned only if the inner extent is (bulk implies it)
    bulk = bulk && (uintptr_t) p.out % 16 == 0;
    c.bulk = bulk ? 1 : 0;
    c.G = std::max(1, kCtThreads / (c.W / (bulk ? 16 / (int) sizeof(T) : 1)));
    if (c.G > c.R) c.G = c.R;
    if (c.chunks > 1) {
        int64_t slots = 0, units = c.chunks, span = 1;
        while (span < c.chunks) {
            if (c.levels >= kCtMaxLevels) return 1;      // axis too long for the tree: the caller walks the columns instead
            c.units[c.levels] = units;
            c.level_off[c.levels] = slots;
            slots += units * c.strips * p.rows;
            units = (units + kCtFan - 1) / kCtFan;
            span *= kCtFan;
            ++c.levels;
        }
        if (slots * 256 >= 0x7fffffffLL) XTB_FAIL(XTB_ERR_UNSUPPORTED, "scan: too many tiles");
        const size_t val_bytes = (size_t) slots * 256 * 2 * sizeof(T);
        void* scratch = nullptr;
        XTB_TRY(ensure_scratch(ctx, val_bytes, &scratch));
        c.vals = (char*) scratch;
        XTB_CUDA(cudaMemsetAsync(scratch, 0, val_bytes, ctx->stream));
    }
    const size_t smem = (size_t) c.R * c.W * sizeof(T);
    const unsigned threads = kCtThreads;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (bulk && p.n_outer <= 1 && c.R <= 256 && !options().no_tma) {
        // 3-D view (inner, axis, outer) of the input; box = (W, R, 1)
        static PFN_cuTensorMapEncodeTiled encode = nullptr;
        if (!encode) {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
                encode = (PFN_cuTensorMapEncodeTiled) fn;
        }
        const int64_t outer_stride = p.n_outer == 1 && p.outer_shape[0] > 1 ? p.outer_stride[0] : p.n * p.in_axis_stride;
        if (encode && (outer_stride * asz) % 16 == 0 && p.in_axis_stride > 0 && outer_stride > 0) {
            const CUtensorMapDataType dt = std::is_same<T, float>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                         : std::is_same<T, double>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64
                                         : sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64;
            const cuuint64_t gdim[3] = {(cuuint64_t) p.inner, (cuuint64_t) p.n, (cuuint64_t) p.rows};
            const cuuint64_t gstr[2] = {(cuuint64_t) (p.in_axis_stride * asz), (cuuint64_t) (outer_stride * asz)};
            const cuuint32_t box[3] = {(cuuint32_t) c.W, (cuuint32_t) c.R, 1u};
            const cuuint32_t estr[3] = {1u, 1u, 1u};
            const CUresult r = encode(&tmap, dt, 3, (void*) p.in, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            c.tma = r == CUDA_SUCCESS ? 1 : 0;
        }
    }
    if (bulk) {
        XTB_CUDA(cudaFuncSetAttribute(k_scan_coltile<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCtTileBytes));
        k_scan_coltile<T, true><<<(unsigned) tiles, threads, smem, ctx->stream>>>(p, c, tmap);
    } else {
        XTB_CUDA(cudaFuncSetAttribute(k_scan_coltile<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCtTileBytes));
        k_scan_coltile<T, false><<<(unsigned) tiles, threads, smem, ctx->stream>>>(p, c, tmap);
    }
    note_launch(c.chunks > 1 ? (c.tma ? "k_scan_coltile[TMA, look-back]" : "k_scan_coltile[look-back]") : (c.tma ? "k_scan_coltile[TMA]" : "k_scan_coltile"));
    return check_launch("k_scan_coltile");
}

// Column walkers apply when the tile rows can be fetched by TMA (dense inner dims of the accumulator dtype,
// 16-byte aligned) and there are enough strips to occupy the GPU.  Returns 1 when they do not apply.
template <class T> static int scan_colwalk(const ScanParams& p, DeviceCtx* ctx) {
    const int64_t asz = sizeof(T);
    bool ok = std::is_same<T, float>::value ? p.in_dtype == XTB_F32
            : std::is_same<T, double>::value ? p.in_dtype == XTB_F64
            : sizeof(T) == 4 ? (p.in_dtype == XTB_I32 || p.in_dtype == XTB_U32)
                             : (p.in_dtype == XTB_I64 || p.in_dtype == XTB_U64);
    ok = ok && !options().no_tma && (uintptr_t) p.in % 16 == 0 && (uintptr_t) p.out % 16 == 0 && (p.inner * asz) % 16 == 0 &&
         (p.in_axis_stride * asz) % 16 == 0 && p.in_axis_stride > 0 && p.n_outer <= 1 && p.n >= 4 * kCwRowsMax;
    int64_t expect = 1;
    for (int d = p.n_inner - 1; d >= 0 && ok; --d) {
        if (p.inner_shape[d] != 1 && p.inner_stride[d] != expect) ok = false;
        expect *= p.inner_shape[d];
    }
    const int64_t outer_stride = p.n_outer == 1 && p.outer_shape[0] > 1 ? p.outer_stride[0] : p.n * p.in_axis_stride;
    ok = ok && (outer_stride * asz) % 16 == 0 && outer_stride > 0;
    if (!ok) return 1;
    // Strip width: the widest box (32 lanes x 16 bytes = 512-byte rows: best use of the DRAM pages) that still gives
    // ~85 % of the SMs a walker -- a walker's throughput is bounded by its SM's TMA / L2 ports -- else narrower.
    const int cv_max = 16 / (int) asz;
    int cv = cv_max;
    while (cv > 1 && ((p.inner + 32 * cv - 1) / (32 * cv)) * p.rows < (int64_t) ctx->sm_count * 85 / 100) cv >>= 1;
    if (options().tile_variant > 0 && options().tile_variant <= cv_max) cv = options().tile_variant;     // development: forced strip width
    ColWalkParams c;
    c.W = 32 * cv;
    c.strips = (int32_t) ((p.inner + c.W - 1) / c.W);
    const int64_t walkers = (int64_t) c.strips * p.rows;
    if (walkers < ctx->sm_count / 3 || walkers >= 0x7fffffffLL || p.inner >= 0x7fffffffLL || p.n >= 0x7fffffffLL) return 1;
    int box_rows = 32;
    if (options().scan_nv == 32 || options().scan_nv == 64) box_rows = options().scan_nv;      // development
    c.box_rows = box_rows;
    c.chunks = (int32_t) ((p.n + box_rows - 1) / box_rows);
    // ring: ~96 KB per walker in boxes of 32 rows (measured on (8192, 8192): fp32 0.96 / fp64 0.94 of the copy peak;
    // deeper rings were SLOWER -- more boxes in flight spread the DRAM accesses over more rows at once)
    const int64_t stage_bytes = (int64_t) box_rows * c.W * asz;
    int64_t stages = (96 * 1024) / stage_bytes;
    stages = std::max<int64_t>(4, std::min<int64_t>(stages, 12));      // a stage is handed back two boxes late (its TMA store must have read it)
    if (options().scan_variant < 0) stages = std::max<int64_t>(4, std::min<int64_t>(-options().scan_variant, 16));
    stages = std::min<int64_t>(stages, std::max<int64_t>(4, c.chunks));
    c.stages = (int32_t) stages;
    static PFN_cuTensorMapEncodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            encode = (PFN_cuTensorMapEncodeTiled) fn;
    }
    if (!encode) return 1;
    CUtensorMap tmap, tmap_out;
    memset(&tmap, 0, sizeof(tmap));
    memset(&tmap_out, 0, sizeof(tmap_out));
    // integer data goes through the maps as raw 32- / 64-bit words; floating-point maps must not flush or canonicalise
    const CUtensorMapDataType dt = sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64;
    const cuuint64_t gdim[3] = {(cuuint64_t) p.inner, (cuuint64_t) p.n, (cuuint64_t) p.rows};
    const cuuint64_t gstr[2] = {(cuuint64_t) (p.in_axis_stride * asz), (cuuint64_t) (outer_stride * asz)};
    const cuuint64_t ostr[2] = {(cuuint64_t) (p.inner * asz), (cuuint64_t) (p.n * p.inner * asz)};     // the result is dense
    const cuuint32_t box[3] = {(cuuint32_t) c.W, (cuuint32_t) box_rows, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    if (encode(&tmap, dt, 3, (void*) p.in, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 1;
    if (encode(&tmap_out, dt, 3, (void*) p.out, gdim, ostr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 1;
    const size_t smem = (size_t) stages * (size_t) stage_bytes;
#define XTB_CW_LAUNCH2(CVV, OPP, RR)                                                                                             \
    do {                                                                                                                         \
        XTB_CUDA(cudaFuncSetAttribute(k_scan_colwalk<T, CVV, OPP, RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024)); \
        k_scan_colwalk<T, CVV, OPP, RR><<<(unsigned) walkers, 64, smem, ctx->stream>>>(p, c, tmap, tmap_out);                    \
    } while (0)
#define XTB_CW_LAUNCH(CVV)                                                                                                       \
    do {                                                                                                                         \
        if (p.op == XTB_RED_PROD) {                                                                                              \
            if (box_rows == 32) XTB_CW_LAUNCH2(CVV, XTB_RED_PROD, 32);                                                           \
            else XTB_CW_LAUNCH2(CVV, XTB_RED_PROD, 64);                                                                          \
        } else {                                                                                                                 \
            if (box_rows == 32) XTB_CW_LAUNCH2(CVV, XTB_RED_SUM, 32);                                                            \
            else XTB_CW_LAUNCH2(CVV, XTB_RED_SUM, 64);                                                                           \
        }                                                                                                                        \
    } while (0)
    if constexpr (sizeof(T) == 4) {
        if (cv == 4) XTB_CW_LAUNCH(4);
        else if (cv == 2) XTB_CW_LAUNCH(2);
        else XTB_CW_LAUNCH(1);
    } else {
        if (cv == 2) XTB_CW_LAUNCH(2);
        else XTB_CW_LAUNCH(1);
    }
#undef XTB_CW_LAUNCH
#undef XTB_CW_LAUNCH2
    note_launch("k_scan_colwalk[TMA ring, reference order]");
    return check_launch("k_scan_colwalk");
}

// ---- strided axis, many columns, pitch or base not 16-byte aligned: the same walkers without the TMA unit --------
// Odd extents ((8191, 8190) along axis 0: the row pitch is 32760 bytes) cannot be described by a tensor map.  The
// walker keeps its shape -- producer warp, consumer warp, shared-memory ring of 32-row boxes, every column accumulated
// in the reference's order -- but the producer fills a box with element-sized cp.async (lane = column: one coalesced
// 128-byte row per instruction) and signals the box through cp.async.mbarrier.arrive; the consumer adds down its column
// and stores each running value straight to the dense output (coalesced rows).  A strip is 32 columns wide.
constexpr int kCpRows = 32;

template <class T, int OP>
__global__ void __launch_bounds__(64) k_scan_colwalk_plain(const __grid_constant__ ScanParams p, const __grid_constant__ ColWalkParams c,
                                                            int64_t outer_stride) {
    extern __shared__ __align__(128) unsigned char cwp_smem[];
    __shared__ __align__(8) unsigned long long s_full[16], s_empty[16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr uint32_t stage_bytes = (uint32_t) (kCpRows * 32 * (int) sizeof(T));
    const uint32_t walker = blockIdx.x;
    const uint32_t o = walker / (uint32_t) c.strips;
    const uint32_t strip = walker - o * (uint32_t) c.strips;
    const int64_t col = (int64_t) strip * 32 + lane;
    const bool live = col < p.inner;
    if (threadIdx.x == 0) {
        for (int s = 0; s < c.stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(smem_u32(&s_full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_empty[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        // ---- producer: every lane copies its column's element of each row of the box ----
        const char* src = p.in + ((int64_t) o * outer_stride + col) * (int64_t) sizeof(T);
        const int64_t sstep = p.in_axis_stride * (int64_t) sizeof(T);
        int s = 0;
        uint32_t round = 0;
        for (int it = 0; it < c.chunks; ++it) {
            if (round > 0) mbar_wait(smem_u32(&s_empty[s]), (round + 1) & 1u);
            const int64_t r0 = (int64_t) it * kCpRows;
            const int nr = (int) (p.n - r0 < kCpRows ? p.n - r0 : kCpRows);
            const uint32_t dst = smem_u32(cwp_smem + (size_t) s * stage_bytes) + (uint32_t) lane * (uint32_t) sizeof(T);
            if (live) {
                if (nr == kCpRows) {
#pragma unroll
                    for (int r = 0; r < kCpRows; ++r)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst + (uint32_t) (r * 32 * (int) sizeof(T))),
                                     "l"(src + (r0 + r) * sstep), "n"((int) sizeof(T)) : "memory");
                } else {
                    for (int r = 0; r < nr; ++r)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst + (uint32_t) (r * 32 * (int) sizeof(T))),
                                     "l"(src + (r0 + r) * sstep), "n"((int) sizeof(T)) : "memory");
                }
            }
            // arrives on the box's barrier once this lane's copies have landed (the 32 lanes are the barrier's count)
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&s_full[s])) : "memory");
            if (++s == c.stages) {
                s = 0;
                ++round;
            }
        }
        return;
    }
    // ---- consumer: lane owns one column ----
    T acc = sident<OP, T>();
    T* out = (T*) p.out + (int64_t) o * p.n * p.inner + col;
    int s = 0;
    uint32_t round = 0;
    for (int it = 0; it < c.chunks; ++it) {
        mbar_wait(smem_u32(&s_full[s]), round & 1u);
        const T* q = (const T*) (cwp_smem + (size_t) s * stage_bytes) + lane;
        const int64_t r0 = (int64_t) it * kCpRows;
        const int nr = (int) (p.n - r0 < kCpRows ? p.n - r0 : kCpRows);
        if (live) {
            T* dst = out + r0 * p.inner;
            if (nr == kCpRows && it > 0) {
#pragma unroll
                for (int r = 0; r < kCpRows; ++r) {
                    acc = sop<OP, T>(acc, q[r * 32]);
                    dst[(int64_t) r * p.inner] = acc;
                }
            } else {
                for (int r = 0; r < nr; ++r) {
                    // out[0] = in[0] (a leading -0.0 survives)
                    acc = (it == 0 && r == 0) ? q[0] : sop<OP, T>(acc, q[r * 32]);
                    dst[(int64_t) r * p.inner] = acc;
                }
            }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_empty[s])) : "memory");
        if (++s == c.stages) {
            s = 0;
            ++round;
        }
    }
}

template <class T> static int scan_colwalk_plain(const ScanParams& p, DeviceCtx* ctx) {
    const int64_t asz = sizeof(T);
    bool ok = std::is_same<T, float>::value ? p.in_dtype == XTB_F32
            : std::is_same<T, double>::value ? p.in_dtype == XTB_F64
            : sizeof(T) == 4 ? (p.in_dtype == XTB_I32 || p.in_dtype == XTB_U32)
                             : (p.in_dtype == XTB_I64 || p.in_dtype == XTB_U64);
    ok = ok && (uintptr_t) p.in % asz == 0 && (uintptr_t) p.out % asz == 0 && p.in_axis_stride > 0 && p.n_outer <= 1 && p.n >= 4 * kCpRows;
    int64_t expect = 1;
    for (int d = p.n_inner - 1; d >= 0 && ok; --d) {
        if (p.inner_shape[d] != 1 && p.inner_stride[d] != expect) ok = false;
        expect *= p.inner_shape[d];
    }
    const int64_t outer_stride = p.n_outer == 1 && p.outer_shape[0] > 1 ? p.outer_stride[0] : p.n * p.in_axis_stride;
    ok = ok && outer_stride > 0;
    if (!ok) return 1;
    ColWalkParams c;
    c.W = 32;
    c.strips = (int32_t) ((p.inner + 31) / 32);
    const int64_t walkers = (int64_t) c.strips * p.rows;
    // one walker streams ~40 GB/s: worth it only with about a walker per SM or more
    if (walkers < ctx->sm_count * 3 / 4 || walkers >= 0x7fffffffLL || p.inner >= 0x7fffffffLL || p.n >= 0x7fffffffLL) return 1;
    c.box_rows = kCpRows;
    c.chunks = (int32_t) ((p.n + kCpRows - 1) / kCpRows);
    const int64_t stage_bytes = (int64_t) kCpRows * 32 * asz;
    // a walker reads 128-byte row pieces a pitch apart: latency-bound, so a deep ring (measured on (8191, 8190) fp32:
    // 8 boxes 3.0 TB/s, 12 boxes 2.9, 16 boxes 4.3; fp64 (4095, 4097): 4 boxes 3.4, 12 boxes 4.4, 16 boxes 4.5)
    int64_t stages = asz == 4 ? 16 : 12;
    if (options().scan_variant < 0) stages = std::max<int64_t>(2, std::min<int64_t>(-options().scan_variant, 16));
    stages = std::min<int64_t>(stages, std::max<int64_t>(2, c.chunks));
    c.stages = (int32_t) stages;
    const size_t smem = (size_t) stages * (size_t) stage_bytes;
    if (p.op == XTB_RED_PROD) {
        XTB_CUDA(cudaFuncSetAttribute(k_scan_colwalk_plain<T, XTB_RED_PROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
        k_scan_colwalk_plain<T, XTB_RED_PROD><<<(unsigned) walkers, 64, smem, ctx->stream>>>(p, c, outer_stride);
    } else {
        XTB_CUDA(cudaFuncSetAttribute(k_scan_colwalk_plain<T, XTB_RED_SUM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
        k_scan_colwalk_plain<T, XTB_RED_SUM><<<(unsigned) walkers, 64, smem, ctx->stream>>>(p, c, outer_stride);
    }
    note_launch("k_scan_colwalk_plain[cp.async ring, reference order]");
    return check_launch("k_scan_colwalk_plain");
}

// ---- strided axis: one thread per column ------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256) k_scan_columns(const __grid_constant__ ScanParams p) {
    const int64_t cols = p.rows * p.inner;
    const int isz = dtype_size(p.in_dtype);
    for (int64_t c = (int64_t) blockIdx.x * 256 + threadIdx.x; c < cols; c += (int64_t) gridDim.x * 256) {
        const uint32_t o = (uint32_t) (c / p.inner);
        const uint32_t in_i = (uint32_t) (c - (int64_t) o * p.inner);
        const char* src = p.in + (scan_offset(o, p.n_outer, p.outer_shape, p.outer_stride, p.outer_div) +
                                  scan_offset(in_i, p.n_inner, p.inner_shape, p.inner_stride, p.inner_div)) * isz;
        T* dst = (T*) p.out + (int64_t) o * p.n * p.inner + in_i;
        const int64_t sstep = p.in_axis_stride * isz;
        T acc = load_cast<T>(src, p.in_dtype);
        dst[0] = acc;
        int64_t i = 1;
        for (; i + 4 <= p.n; i += 4) {
            T v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = load_cast<T>(src + (i + u) * sstep, p.in_dtype);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc = scan_op<T>(p.op, acc, v[u]);
                dst[(i + u) * p.inner] = acc;
            }
        }
        for (; i < p.n; ++i) {
            acc = scan_op<T>(p.op, acc, load_cast<T>(src + i * sstep, p.in_dtype));
            dst[i * p.inner] = acc;
        }
    }
}

// rows longer than one staged tile: reduce ahead, scan from L2 (k_scan_ahead)
template <class T, int NV, int TH> static int launch_scan_ahead(ScanParams q, DeviceCtx* ctx) {
    using C = ScanTile<T>;
    const int64_t te = (int64_t) TH * NV * (16 / (int64_t) sizeof(T));       // 16 KB tiles (256 threads x 4 vectors)
    const int64_t tpr = (q.n + te - 1) / te;
    const int64_t tiles = tpr * q.rows;
    LbTree tr;
    memset(&tr, 0, sizeof(tr));
    // look-ahead distance: 24-32 MB = 1536-2048 tiles -- more than twice the ~740-890 CTAs in flight, so that a tile's scan visit
    // never meets aggregates that are still being computed; it has to stay in the 126 MB L2 next to as many
    // bytes of stores.  At most half the work.
    int64_t ahead_mb = options().scan_variant > 0 ? options().scan_variant : (TH == 128 && NV == 8 ? 32 : 24);
    int64_t ahead = std::min<int64_t>((ahead_mb << 20) / (te * (int64_t) sizeof(T)), std::max<int64_t>(64, tiles / 2)) / 64 * 64;
    if (ahead < 64) ahead = 64;
    tr.ahead = (uint32_t) ahead;
    if (tiles + ahead >= 0x7fffffffLL) XTB_FAIL(XTB_ERR_UNSUPPORTED, "scan: too many tiles");
    q.tiles_per_row = (uint32_t) tpr;
    q.total_tiles = (uint32_t) tiles;
    int64_t slots = 0, units = tpr;
    while (true) {
        if (tr.levels >= kLbMaxLevels) XTB_FAIL(XTB_ERR_UNSUPPORTED, "scan: row too long");
        tr.units[tr.levels] = (uint32_t) units;
        tr.level_off[tr.levels] = (uint32_t) slots;
        slots += units * q.rows;
        ++tr.levels;
        if (units <= kLbFan) break;
        units = (units + kLbFan - 1) / kLbFan;
    }
    if (slots >= 0x7fffffffLL) XTB_FAIL(XTB_ERR_UNSUPPORTED, "scan: too many tiles");
    const size_t bytes = ((size_t) slots * C::SLOT + 255) / 256 * 256;
    void* scratch = nullptr;
    XTB_TRY(ensure_scratch(ctx, bytes, &scratch));
    tr.slots = (char*) scratch;
    XTB_CUDA(cudaMemsetAsync(scratch, 0, bytes, ctx->stream));
    if (q.op == XTB_RED_PROD) k_scan_ahead<T, NV, XTB_RED_PROD, TH><<<(unsigned) (tiles + ahead), TH, 0, ctx->stream>>>(q, tr);
    else k_scan_ahead<T, NV, XTB_RED_SUM, TH><<<(unsigned) (tiles + ahead), TH, 0, ctx->stream>>>(q, tr);
    note_launch("k_scan_ahead[reduce ahead, scan from L2]");
    return check_launch("k_scan_ahead");
}

template <class T> static int launch_scan(const ScanParams& p, DeviceCtx* ctx, bool columns) {
    if (columns) {
        const int64_t cols = p.rows * p.inner;
        const unsigned gx = (unsigned) std::min<int64_t>((cols + 255) / 256, (int64_t) ctx->sm_count * 32);
        k_scan_columns<T><<<gx, 256, 0, ctx->stream>>>(p);
        note_launch("k_scan_columns");
        return check_launch("k_scan_columns");
    }
    // contiguous: state = one slot per tile | one slot per block of kScanWindow tiles
    ScanParams q = p;
    using C = ScanTile<T>;
    if (q.rows > 1 && q.n <= C::WARP_ELEMS) {
        // short rows: a warp per row, or -- dense input, power-of-two row bytes -- several rows per warp
        const int64_t row_bytes = q.n * (int64_t) sizeof(T);
        bool dense = q.vec_io != 0 && row_bytes >= 16 && (row_bytes & (row_bytes - 1)) == 0;
        int64_t expect = q.n;
        for (int d = q.n_outer - 1; d >= 0 && dense; --d) {
            if (q.outer_shape[d] != 1 && q.outer_stride[d] != expect) dense = false;
            expect *= q.outer_shape[d];
        }
        int64_t units = q.rows;
        if (dense) {
            q.packed = 1;
            q.seg_vecs = (int32_t) (row_bytes / 16);
            units = (q.rows * q.n + C::WARP_ELEMS - 1) / C::WARP_ELEMS;
        }
        const unsigned grid = (unsigned) ((units + kScanWarpsPerTile - 1) / kScanWarpsPerTile);
        k_scan_tiles<T, true><<<grid, kScanThreads, 0, ctx->stream>>>(q);
        note_launch(dense ? "k_scan_tiles[rows packed per warp]" : "k_scan_tiles[warp per row]");
        return check_launch("k_scan_tiles");
    }
    const int64_t row_bytes = q.n * (int64_t) sizeof(T);
    // rows longer than one staged tile: reduce ahead, scan from L2 (k_scan_ahead)
    // rows longer than one staged tile: reduce ahead, scan from L2 (k_scan_ahead)
    if (q.n * (int64_t) sizeof(T) > 64 * 1024) {
        // 4-warp CTAs, 8 vectors per thread (16 KB tiles), 32 MB of look-ahead: the CTA barrier of the scan visit waits for
        // the slowest of 4 warps instead of 8 (measured on 2^26 fp32: 5.55 TB/s vs 5.33 with 8-warp CTAs / 4 vectors / 24 MB;
        // sweep in profiles/r02_scan_sweep.log).  scan_nv = 4 / tile_variant = 256 select the other shapes.
        if (options().tile_variant == 256) return options().scan_nv == 8 ? launch_scan_ahead<T, 8, 256>(q, ctx) : launch_scan_ahead<T, 4, 256>(q, ctx);
        return options().scan_nv == 4 ? launch_scan_ahead<T, 4, 128>(q, ctx) : launch_scan_ahead<T, 8, 128>(q, ctx);
    }
    // staged super-tile configuration: threads per CTA / CTAs per SM / super-tile bytes
    // staged super-tiles: long rows (several tiles, look-back) use 8 scan warps + the look-back warp on 64 KB,
    // 3 CTAs per SM; rows of one tile use 4 warps on up to 32 KB, 6 CTAs per SM
    const bool long_rows = row_bytes > 32 * 1024;
    const int st_threads = long_rows ? 256 : 128;
    const int64_t st_max = long_rows ? 64 * 1024 : 32 * 1024;
    const int64_t granule = (st_threads / 32) * 128 * 16;    // every warp owns whole 128-vector sub-chunks
    const bool staged = row_bytes >= granule;
    int64_t tile_elems = C::TILE;
    if (staged) {
        int64_t tb = row_bytes <= st_max ? (row_bytes + granule - 1) / granule * granule : st_max;
        tile_elems = tb / (int64_t) sizeof(T);
        q.st_elems = (int32_t) tile_elems;
    }
    const int64_t tpr = (q.n + tile_elems - 1) / tile_elems;
    const int64_t tiles = tpr * q.rows;
    if (tiles >= 0x7fffffffLL) XTB_FAIL(XTB_ERR_UNSUPPORTED, "scan: too many tiles");
    q.tiles_per_row = (uint32_t) tpr;
    q.total_tiles = (uint32_t) tiles;
    if (tpr > 1) {
        const int64_t bpr = (tpr + kScanWindow - 1) / kScanWindow;
        const size_t agg_bytes = ((size_t) tiles * C::SLOT + 255) / 256 * 256;
        const size_t blk_bytes = ((size_t) (bpr * q.rows) * C::SLOT + 255) / 256 * 256;
        void* scratch = nullptr;
        XTB_TRY(ensure_scratch(ctx, agg_bytes + blk_bytes, &scratch));
        char* s = (char*) scratch;
        q.aggregate = s;
        q.prefix = s + agg_bytes;
        XTB_CUDA(cudaMemsetAsync(s, 0, agg_bytes + blk_bytes, ctx->stream));
    }
    if (staged) {
        const size_t smem = (size_t) tile_elems * sizeof(T);
        if (tpr > 1) {
            XTB_CUDA(cudaFuncSetAttribute(k_scan_stile<T, 256, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStMaxBytesLimit));
            k_scan_stile<T, 256, 3, true><<<q.total_tiles, 256 + 32, smem, ctx->stream>>>(q);
        } else if (long_rows) {
            XTB_CUDA(cudaFuncSetAttribute(k_scan_stile<T, 256, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStMaxBytesLimit));
            k_scan_stile<T, 256, 3, false><<<q.total_tiles, 256, smem, ctx->stream>>>(q);
        } else {
            XTB_CUDA(cudaFuncSetAttribute(k_scan_stile<T, 128, 6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStMaxBytesLimit));
            k_scan_stile<T, 128, 6, false><<<q.total_tiles, 128, smem, ctx->stream>>>(q);
        }
        note_launch(tpr > 1 ? "k_scan_stile[look-back]" : "k_scan_stile");
        return check_launch("k_scan_stile");
    }
    k_scan_tiles<T, false><<<q.total_tiles, kScanThreads, 0, ctx->stream>>>(q);
    note_launch(tpr > 1 ? "k_scan_tiles[look-back]" : "k_scan_tiles");
    return check_launch("k_scan_tiles");
}

}  // namespace xtb

using namespace xtb;

extern "C" int xtb_scan(int op, int acc_type, const xtb_operand* in, int axis, const xtb_operand* out) {
    if (!in || !out) XTB_FAIL(XTB_ERR_INVALID, "null argument");
    if (op != XTB_RED_SUM && op != XTB_RED_PROD) XTB_FAIL(XTB_ERR_INVALID, "scan supports sum and prod");
    if (acc_type < XTB_I32 || acc_type > XTB_F64) XTB_FAIL(XTB_ERR_INVALID, "accumulator must be a register type");
    if (in->ndim < 0 || in->ndim > XTB_MAX_DIM) XTB_FAIL(XTB_ERR_INVALID, "rank out of range");
    if (in->dtype < 0 || in->dtype >= XTB_DTYPE_COUNT) XTB_FAIL(XTB_ERR_INVALID, "bad dtype");
    if (axis >= in->ndim) XTB_FAIL(XTB_ERR_AXIS, "Axis larger than expression dimension in accumulator.");
    if (out->dtype != acc_type) XTB_FAIL(XTB_ERR_INVALID, "scan output must have the accumulator dtype");
    const int nd = in->ndim;
    int64_t total = 1;
    for (int d = 0; d < nd; ++d) total *= in->shape[d];
    // the result is dense row-major (the reference scans a fresh copy of the input)
    {
        int64_t ototal = 1, expect = 1;
        for (int d = 0; d < out->ndim; ++d) ototal *= out->shape[d];
        if (ototal != total) XTB_FAIL(XTB_ERR_SHAPE, "scan output has %lld elements, input %lld", (long long) ototal, (long long) total);
        for (int d = out->ndim - 1; d >= 0; --d) {
            if (out->shape[d] != 1 && out->stride[d] != expect) XTB_FAIL(XTB_ERR_UNSUPPORTED, "scan output must be dense row-major");
            expect *= out->shape[d];
        }
        if (axis >= 0 && out->ndim != nd) XTB_FAIL(XTB_ERR_SHAPE, "scan output rank mismatch");
    }
    if (total == 0) return XTB_OK;
    DeviceCtx* ctx;
    XTB_TRY(get_ctx(&ctx));
    XTB_LAUNCH_LOCK(ctx);

    ScanParams p;
    memset(&p, 0, sizeof(p));
    p.in = operand_ptr(in, dtype_size(in->dtype));
    p.out = operand_ptr(out, dtype_size(out->dtype));
    p.in_dtype = in->dtype;
    p.op = op;
    int64_t st[XTB_MAX_DIM];
    for (int d = 0; d < nd; ++d) st[d] = in->shape[d] == 1 ? 0 : in->stride[d];
    bool columns = false;
    if (axis < 0) {
        // flat scan in row-major traversal order: needs a dense row-major view of the input to be one run
        bool dense = true;
        int64_t expect = 1;
        for (int d = nd - 1; d >= 0; --d) {
            if (in->shape[d] != 1 && st[d] != expect) dense = false;
            expect *= in->shape[d];
        }
        if (dense) {
            p.n = total;
            p.rows = 1;
            p.in_axis_stride = 1;
            p.n_outer = 0;
        } else {
            XTB_FAIL(XTB_ERR_UNSUPPORTED, "flat scan of a non-contiguous view: evaluate it into a container first");
        }
    } else {
        p.n = in->shape[axis];
        p.in_axis_stride = st[axis];
        int64_t outer = 1, inner = 1;
        for (int d = 0; d < axis; ++d) {
            p.outer_shape[p.n_outer] = in->shape[d];
            p.outer_stride[p.n_outer] = st[d];
            ++p.n_outer;
            outer *= in->shape[d];
        }
        for (int d = axis + 1; d < nd; ++d) {
            p.inner_shape[p.n_inner] = in->shape[d];
            p.inner_stride[p.n_inner] = st[d];
            ++p.n_inner;
            inner *= in->shape[d];
        }
        p.rows = outer;
        p.inner = inner;
        columns = inner > 1;
        if (outer >= 0x7fffffffLL || inner >= 0x7fffffffLL) XTB_FAIL(XTB_ERR_UNSUPPORTED, "scan: too many independent rows");
    }
    for (int d = 0; d < p.n_outer; ++d) p.outer_div[d] = make_fastdiv((uint32_t) std::min<int64_t>(p.outer_shape[d], 0x7fffffff));
    for (int d = 0; d < p.n_inner; ++d) p.inner_div[d] = make_fastdiv((uint32_t) std::min<int64_t>(p.inner_shape[d], 0x7fffffff));
    if (!columns) {
        const size_t asz = dtype_size(acc_type);
        p.vec_io = in->dtype == acc_type && p.in_axis_stride == 1 && ((uintptr_t) p.in % 16 == 0) && ((uintptr_t) p.out % 16 == 0) &&
                   (p.rows == 1 || (p.n * asz) % 16 == 0);
        if (p.vec_io && p.rows > 1)
            for (int d = 0; d < p.n_outer; ++d) p.vec_io = p.vec_io && (p.outer_stride[d] * asz) % 16 == 0;
    } else {
        // a long axis: column tiles (single pass).  Short axes with very many columns keep the plain
        // thread-per-column walk (exactly the reference's order, enough parallelism from the columns).
        if (p.n >= 128 && p.inner < 0x7fffffffLL - 256) {
            int r;
            // enough column strips to fill the GPU: walkers (no chaining along the axis, reference order)
            switch (acc_type) {
                case XTB_I32: case XTB_U32: r = scan_colwalk<uint32_t>(p, ctx); break;
                case XTB_I64: case XTB_U64: r = scan_colwalk<unsigned long long>(p, ctx); break;
                case XTB_F32: r = scan_colwalk<float>(p, ctx); break;
                default: r = scan_colwalk<double>(p, ctx); break;
            }
            if (r <= 0) return r;
            // the same walkers without TMA (unaligned pitch / base)
            switch (acc_type) {
                case XTB_I32: case XTB_U32: r = scan_colwalk_plain<uint32_t>(p, ctx); break;
                case XTB_I64: case XTB_U64: r = scan_colwalk_plain<unsigned long long>(p, ctx); break;
                case XTB_F32: r = scan_colwalk_plain<float>(p, ctx); break;
                default: r = scan_colwalk_plain<double>(p, ctx); break;
            }
            if (r <= 0) return r;
            switch (acc_type) {
                case XTB_I32: case XTB_U32: r = scan_coltile<uint32_t>(p, ctx); break;
                case XTB_I64: case XTB_U64: r = scan_coltile<unsigned long long>(p, ctx); break;
                case XTB_F32: r = scan_coltile<float>(p, ctx); break;
                default: r = scan_coltile<double>(p, ctx); break;
            }
            if (r <= 0) return r;
        }
    }
    switch (acc_type) {
        case XTB_I32: case XTB_U32: return launch_scan<uint32_t>(p, ctx, columns);
        case XTB_I64: case XTB_U64: return launch_scan<unsigned long long>(p, ctx, columns);
        case XTB_F32: return launch_scan<float>(p, ctx, columns);
        default: return launch_scan<double>(p, ctx, columns);
    }
}
