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
#include "xtb_common.hpp"
#include "xtb_ops.cuh"

namespace xtb {

constexpr int kScanThreads = 512;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;
constexpr int kLookbackBuf = 1024;

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
    uint32_t* ticket;
    uint32_t* status;             // 0 = not ready, 1 = aggregate ready, 2 = inclusive prefix ready
    char* aggregate;
    char* prefix;
    unsigned long long* packed;   // 4-byte accumulators: {status, value} in one 64-bit word per tile
    int32_t vec_io;               // input is the accumulator dtype, unit stride, 16-byte aligned rows
    const char* tile_incl;        // != nullptr: inclusive scan of the tile totals [rows][tiles_per_row] is precomputed
                                  // (long rows: reduce-then-scan, no look-back chain)
    // chunked column scan
    int64_t chunk;                // rows per chunk (0: whole axis)
    const char* carry;            // [nchunks][rows * inner] inclusive chunk totals scan (acc dtype)
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

// ---- contiguous scan, decoupled look-back ---------------------------------------------------
template <class T>
__global__ void __launch_bounds__(kScanThreads, 2) k_scan_lookback(const __grid_constant__ ScanParams p) {
    __shared__ uint32_t s_tile;
    __shared__ T s_warp[kScanThreads / 32];
    __shared__ T s_prefix;
    __shared__ T s_buf[kLookbackBuf];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int op = p.op;
    // tiles are taken in launch order, so every predecessor of a tile is already running
    if (tid == 0) s_tile = atomicAdd(p.ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= p.total_tiles) return;
    const uint32_t row = tile / p.tiles_per_row;
    const uint32_t trow = tile - row * p.tiles_per_row;   // tile index within its row
    const int64_t base = (int64_t) trow * kScanTile + (int64_t) tid * kScanItems;
    const int isz = dtype_size(p.in_dtype);
    const char* in_row = p.in + scan_offset(row, p.n_outer, p.outer_shape, p.outer_stride, p.outer_div) * isz;
    // load + thread-local inclusive scan
    T x[kScanItems];
    const bool full_tile = base + kScanItems <= p.n;
    if (p.vec_io && full_tile) {
        constexpr int PER = 16 / sizeof(T);
        const uint4* src = (const uint4*) (in_row + base * sizeof(T));
#pragma unroll
        for (int q = 0; q < kScanItems / PER; ++q) {
            const uint4 r = ldg_stream_16(src + q);
            memcpy(&x[q * PER], &r, 16);
        }
    } else {
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            const int64_t j = base + i;
            x[i] = j < p.n ? load_cast<T>(in_row + j * p.in_axis_stride * isz, p.in_dtype) : scan_identity<T>(op);
        }
    }
#pragma unroll
    for (int i = 1; i < kScanItems; ++i) x[i] = scan_op<T>(op, x[i - 1], x[i]);
    // block-wide exclusive scan of the thread totals
    T incl = x[kScanItems - 1];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const T y = shfl_up_t<T>(incl, d);
        if (lane >= d) incl = scan_op<T>(op, y, incl);
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    T warp_off = scan_identity<T>(op);
    for (int w = 0; w < warp; ++w) warp_off = scan_op<T>(op, warp_off, s_warp[w]);
    const T excl_lane = shfl_up_t<T>(incl, 1);
    T thread_off = lane == 0 ? warp_off : scan_op<T>(op, warp_off, excl_lane);
    // tile aggregate = total of the last warp prefix
    T tile_prefix = scan_identity<T>(op);
    if (p.tile_incl != nullptr) {
        if (trow > 0) tile_prefix = ((const T*) p.tile_incl)[tile - 1];
    } else if (p.tiles_per_row > 1) {
        T* agg = (T*) p.aggregate;
        T* pre = (T*) p.prefix;
        if (tid == kScanThreads - 1) {
            const T total = scan_op<T>(op, thread_off, x[kScanItems - 1]);
            if constexpr (sizeof(T) == 4) {
                uint32_t bits;
                memcpy(&bits, &total, 4);
                const unsigned long long w = ((unsigned long long) (trow == 0 ? 2u : 1u) << 32) | bits;
                *((volatile unsigned long long*) &p.packed[tile]) = w;     // one 64-bit store: status + value
            } else if (trow == 0) {
                pre[tile] = total;
                __threadfence();
                atomicExch(&p.status[tile], 2u);
            } else {
                agg[tile] = total;
                __threadfence();
                atomicExch(&p.status[tile], 1u);
            }
            s_prefix = total;  // reused below as the tile's own aggregate
        }
        __syncthreads();
        if (trow > 0) {
            if (warp == 0) {
                // look back over the predecessors in this row, 32 at a time; collect aggregates
                // (nearest first) until a tile with an inclusive prefix is found
                int collected = 0;
                T found_prefix = scan_identity<T>(op);
                int64_t look = (int64_t) trow - 1;    // nearest predecessor still to inspect
                bool done = false;
                while (!done) {
                    const int64_t mine = look - lane;
                    uint32_t st = 2u;  // lanes before the start of the row behave like "prefix = identity"
                    T v = scan_identity<T>(op);
                    if (mine >= 0) {
                        const uint32_t idx = row * p.tiles_per_row + (uint32_t) mine;
                        if constexpr (sizeof(T) == 4) {
                            unsigned long long w;
                            do {
                                w = *((volatile unsigned long long*) &p.packed[idx]);
                            } while ((w >> 32) == 0ull);
                            st = (uint32_t) (w >> 32);
                            const uint32_t bits = (uint32_t) w;
                            memcpy(&v, &bits, 4);
                        } else {
                            do {
                                st = *((volatile uint32_t*) &p.status[idx]);
                            } while (st == 0u);
                            __threadfence();
                            v = st == 2u ? ((volatile T*) pre)[idx] : ((volatile T*) agg)[idx];
                        }
                    }
                    const uint32_t has_prefix = __ballot_sync(0xffffffffu, st == 2u);
                    const int first = has_prefix ? __ffs(has_prefix) - 1 : 32;   // nearest lane with a prefix
                    if (lane < first && collected + lane < kLookbackBuf) s_buf[collected + lane] = v;
                    if (has_prefix) {
                        found_prefix = __shfl_sync(0xffffffffu, v, first);
                        collected += first;
                        done = true;
                    } else {
                        collected += 32;
                        look -= 32;
                        if (collected + 32 > kLookbackBuf) {
                            // buffer full: wait for the inclusive prefix of the next tile instead
                            const uint32_t idx = row * p.tiles_per_row + (uint32_t) look;
                            if (lane == 0) {
                                if constexpr (sizeof(T) == 4) {
                                    unsigned long long w;
                                    do {
                                        w = *((volatile unsigned long long*) &p.packed[idx]);
                                    } while ((w >> 32) != 2ull);
                                    const uint32_t bits = (uint32_t) w;
                                    memcpy(&found_prefix, &bits, 4);
                                } else {
                                    uint32_t s2;
                                    do {
                                        s2 = *((volatile uint32_t*) &p.status[idx]);
                                    } while (s2 != 2u);
                                    __threadfence();
                                    found_prefix = ((volatile T*) pre)[idx];
                                }
                            }
                            found_prefix = __shfl_sync(0xffffffffu, found_prefix, 0);
                            done = true;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    // ascending tile order: prefix, then the collected aggregates from far to near
                    T acc = found_prefix;
                    for (int i = collected - 1; i >= 0; --i) acc = scan_op<T>(op, acc, s_buf[i]);
                    const T own = s_prefix;
                    const T incl_prefix = scan_op<T>(op, acc, own);
                    if constexpr (sizeof(T) == 4) {
                        uint32_t bits;
                        memcpy(&bits, &incl_prefix, 4);
                        *((volatile unsigned long long*) &p.packed[tile]) = (2ull << 32) | bits;
                    } else {
                        pre[tile] = incl_prefix;
                        __threadfence();
                        atomicExch(&p.status[tile], 2u);
                    }
                    s_prefix = acc;
                }
            }
            __syncthreads();
            tile_prefix = s_prefix;
        }
    }
    const T off = (p.tiles_per_row > 1 && trow > 0) ? scan_op<T>(op, tile_prefix, thread_off) : thread_off;
    T* out_row = (T*) p.out + (int64_t) row * p.n;
    const bool has_off = !(trow == 0 && tid == 0);
    if (has_off) {
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) x[i] = scan_op<T>(op, off, x[i]);
    }
    if (p.vec_io && full_tile && ((p.n * sizeof(T)) % 16 == 0 || p.rows == 1)) {
        constexpr int PER = 16 / sizeof(T);
        uint4* dst = (uint4*) (out_row + base);
#pragma unroll
        for (int q = 0; q < kScanItems / PER; ++q) {
            uint4 r;
            memcpy(&r, &x[q * PER], 16);
            stg_stream_16(dst + q, r);
        }
    } else {
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            const int64_t j = base + i;
            if (j < p.n) out_row[j] = x[i];
        }
    }
}

// per-tile totals for long rows (phase 1 of reduce-then-scan)
template <class T>
__global__ void __launch_bounds__(kScanThreads, 2) k_scan_tile_sums(const __grid_constant__ ScanParams p, T* sums) {
    __shared__ T s_warp[kScanThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    const uint32_t row = tile / p.tiles_per_row;
    const uint32_t trow = tile - row * p.tiles_per_row;
    const int64_t base = (int64_t) trow * kScanTile + (int64_t) tid * kScanItems;
    const int isz = dtype_size(p.in_dtype);
    const char* in_row = p.in + scan_offset(row, p.n_outer, p.outer_shape, p.outer_stride, p.outer_div) * isz;
    T x[kScanItems];
    if (p.vec_io && base + kScanItems <= p.n) {
        constexpr int PER = 16 / sizeof(T);
        const uint4* src = (const uint4*) (in_row + base * sizeof(T));
#pragma unroll
        for (int q = 0; q < kScanItems / PER; ++q) {
            const uint4 r = ldg_stream_16(src + q);
            memcpy(&x[q * PER], &r, 16);
        }
    } else {
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            const int64_t j = base + i;
            x[i] = j < p.n ? load_cast<T>(in_row + j * p.in_axis_stride * isz, p.in_dtype) : scan_identity<T>(p.op);
        }
    }
    T acc = x[0];
#pragma unroll
    for (int i = 1; i < kScanItems; ++i) acc = scan_op<T>(p.op, acc, x[i]);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const T y = shfl_up_t<T>(acc, d);
        if (lane >= d) acc = scan_op<T>(p.op, y, acc);
    }
    if (lane == 31) s_warp[warp] = acc;
    __syncthreads();
    if (tid == 0) {
        T t = s_warp[0];
        for (int w = 1; w < kScanThreads / 32; ++w) t = scan_op<T>(p.op, t, s_warp[w]);
        sums[tile] = t;
    }
}

// ---- strided axis: one thread per column ------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256) k_scan_columns(const __grid_constant__ ScanParams p) {
    const int64_t cols = p.rows * p.inner;
    const int isz = dtype_size(p.in_dtype);
    if (p.chunk > 0) {
        // chunked: blockIdx.y owns rows [i0, i1) of every column and starts from the carry of the
        // previous chunks (inclusive scan of the per-chunk totals, computed by two earlier launches)
        const int64_t i0 = (int64_t) blockIdx.y * p.chunk;
        const int64_t i1 = i0 + p.chunk < p.n ? i0 + p.chunk : p.n;
        for (int64_t c = (int64_t) blockIdx.x * 256 + threadIdx.x; c < cols; c += (int64_t) gridDim.x * 256) {
            const uint32_t o = (uint32_t) (c / p.inner);
            const uint32_t in_i = (uint32_t) (c - (int64_t) o * p.inner);
            const char* src = p.in + (scan_offset(o, p.n_outer, p.outer_shape, p.outer_stride, p.outer_div) +
                                      scan_offset(in_i, p.n_inner, p.inner_shape, p.inner_stride, p.inner_div)) * isz;
            T* dst = (T*) p.out + (int64_t) o * p.n * p.inner + in_i;
            const int64_t sstep = p.in_axis_stride * isz;
            T acc = scan_identity<T>(p.op);
            bool have = false;
            if (blockIdx.y > 0) {
                acc = ((const T*) p.carry)[((int64_t) o * gridDim.y + (blockIdx.y - 1)) * p.inner + in_i];
                have = true;
            }
            int64_t i = i0;
            for (; i + 4 <= i1; i += 4) {
                T v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = load_cast<T>(src + (i + u) * sstep, p.in_dtype);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    acc = have ? scan_op<T>(p.op, acc, v[u]) : v[u];
                    have = true;
                    dst[(i + u) * p.inner] = acc;
                }
            }
            for (; i < i1; ++i) {
                const T v = load_cast<T>(src + i * sstep, p.in_dtype);
                acc = have ? scan_op<T>(p.op, acc, v) : v;
                have = true;
                dst[i * p.inner] = acc;
            }
        }
        return;
    }
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

// per-chunk column totals: totals[(o * nch + c) * inner + i] = op over rows [c*chunk, (c+1)*chunk)
template <class T>
__global__ void __launch_bounds__(256) k_scan_chunk_totals(const __grid_constant__ ScanParams p, T* totals) {
    const int64_t cols = p.rows * p.inner;
    const int isz = dtype_size(p.in_dtype);
    const int64_t i0 = (int64_t) blockIdx.y * p.chunk;
    const int64_t i1 = i0 + p.chunk < p.n ? i0 + p.chunk : p.n;
    for (int64_t c = (int64_t) blockIdx.x * 256 + threadIdx.x; c < cols; c += (int64_t) gridDim.x * 256) {
        const uint32_t o = (uint32_t) (c / p.inner);
        const uint32_t in_i = (uint32_t) (c - (int64_t) o * p.inner);
        const char* src = p.in + (scan_offset(o, p.n_outer, p.outer_shape, p.outer_stride, p.outer_div) +
                                  scan_offset(in_i, p.n_inner, p.inner_shape, p.inner_stride, p.inner_div)) * isz;
        const int64_t sstep = p.in_axis_stride * isz;
        T acc = load_cast<T>(src + i0 * sstep, p.in_dtype);
        int64_t i = i0 + 1;
        for (; i + 8 <= i1; i += 8) {
            T v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = load_cast<T>(src + (i + u) * sstep, p.in_dtype);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = scan_op<T>(p.op, acc, v[u]);
        }
        for (; i < i1; ++i) acc = scan_op<T>(p.op, acc, load_cast<T>(src + i * sstep, p.in_dtype));
        totals[((int64_t) o * gridDim.y + blockIdx.y) * p.inner + in_i] = acc;
    }
}

template <class T> static int scan_chunked_t(ScanParams p, int64_t chunk, int64_t nch, DeviceCtx* ctx) {
    const int64_t cols = p.rows * p.inner;
    void* scratch = nullptr;
    const size_t each = ((size_t) cols * nch * sizeof(T) + 255) / 256 * 256;
    XTB_TRY(ensure_scratch(ctx, 2 * each, &scratch));
    T* totals = (T*) scratch;
    T* carry = (T*) ((char*) scratch + each);
    p.chunk = chunk;
    const unsigned gx = (unsigned) std::min<int64_t>((cols + 255) / 256, (int64_t) ctx->sm_count * 32);
    k_scan_chunk_totals<T><<<dim3(gx, (unsigned) nch), 256, 0, ctx->stream>>>(p, totals);
    note_launch("k_scan_chunk_totals");
    XTB_TRY(check_launch("k_scan_chunk_totals"));
    // inclusive scan of the chunk totals along the chunk index (small): totals[o][c][i] -> carry[o][c][i]
    ScanParams q;
    memset(&q, 0, sizeof(q));
    q.in = (const char*) totals;
    q.out = (char*) carry;
    q.in_dtype = sizeof(T) == 4 ? (std::is_same<T, float>::value ? XTB_F32 : XTB_U32) : (std::is_same<T, double>::value ? XTB_F64 : XTB_U64);
    q.op = p.op;
    q.n = nch;
    q.rows = p.rows;
    q.inner = p.inner;
    q.in_axis_stride = p.inner;
    q.n_outer = 1;
    q.outer_shape[0] = p.rows;
    q.outer_stride[0] = nch * p.inner;
    q.outer_div[0] = make_fastdiv((uint32_t) std::min<int64_t>(p.rows, 0x7fffffff));
    q.n_inner = 1;
    q.inner_shape[0] = p.inner;
    q.inner_stride[0] = 1;
    q.inner_div[0] = make_fastdiv((uint32_t) std::min<int64_t>(p.inner, 0x7fffffff));
    k_scan_columns<T><<<gx, 256, 0, ctx->stream>>>(q);
    note_launch("k_scan_columns[carry]");
    XTB_TRY(check_launch("k_scan_columns"));
    p.carry = (const char*) carry;
    k_scan_columns<T><<<dim3(gx, (unsigned) nch), 256, 0, ctx->stream>>>(p);
    note_launch("k_scan_columns[chunked]");
    return check_launch("k_scan_columns");
}

static int scan_chunked(int acc_type, const ScanParams& p, int64_t chunk, int64_t nch, DeviceCtx* ctx) {
    switch (acc_type) {
        case XTB_I32: case XTB_U32: return scan_chunked_t<uint32_t>(p, chunk, nch, ctx);
        case XTB_I64: case XTB_U64: return scan_chunked_t<unsigned long long>(p, chunk, nch, ctx);
        case XTB_F32: return scan_chunked_t<float>(p, chunk, nch, ctx);
        default: return scan_chunked_t<double>(p, chunk, nch, ctx);
    }
}

template <class T> static int launch_scan(const ScanParams& p, DeviceCtx* ctx, bool columns) {
    if (columns) {
        const int64_t cols = p.rows * p.inner;
        const unsigned gx = (unsigned) std::min<int64_t>((cols + 255) / 256, (int64_t) ctx->sm_count * 32);
        const unsigned gy = p.chunk > 0 ? (unsigned) ((p.n + p.chunk - 1) / p.chunk) : 1u;
        k_scan_columns<T><<<dim3(gx, gy), 256, 0, ctx->stream>>>(p);
        note_launch(p.chunk > 0 ? "k_scan_columns[chunked]" : "k_scan_columns");
        return check_launch("k_scan_columns");
    }
    if (p.tiles_per_row > 32 && p.tile_incl == nullptr) {
        // long rows: the look-back chain (32 tiles per poll round) would bound throughput; do
        // reduce-then-scan instead: tile totals -> their inclusive scan (recursively, a short row) ->
        // final pass with the tile prefixes read from memory.  Deterministic, 3 passes over memory.
        T* sums = nullptr;
        T* incl = nullptr;
        const size_t bytes = (size_t) p.total_tiles * sizeof(T);
        XTB_CUDA(cudaMallocAsync((void**) &sums, bytes, ctx->stream));
        XTB_CUDA(cudaMallocAsync((void**) &incl, bytes, ctx->stream));
        k_scan_tile_sums<T><<<p.total_tiles, kScanThreads, 0, ctx->stream>>>(p, sums);
        note_launch("k_scan_tile_sums");
        XTB_TRY(check_launch("k_scan_tile_sums"));
        ScanParams q;
        memset(&q, 0, sizeof(q));
        q.in = (const char*) sums;
        q.out = (char*) incl;
        q.in_dtype = sizeof(T) == 4 ? (std::is_same<T, float>::value ? XTB_F32 : XTB_U32) : (std::is_same<T, double>::value ? XTB_F64 : XTB_U64);
        q.op = p.op;
        q.n = p.tiles_per_row;
        q.rows = p.rows;
        q.in_axis_stride = 1;
        q.n_outer = 1;
        q.outer_shape[0] = p.rows;
        q.outer_stride[0] = p.tiles_per_row;
        q.outer_div[0] = make_fastdiv((uint32_t) std::min<int64_t>(p.rows, 0x7fffffff));
        const int64_t tpr2 = (q.n + kScanTile - 1) / kScanTile;
        q.tiles_per_row = (uint32_t) tpr2;
        q.total_tiles = (uint32_t) (tpr2 * p.rows);
        // look-back state of the inner scan lives in the context scratch after the outer scan's state
        void* scratch = nullptr;
        const size_t state = 256 + ((size_t) q.total_tiles * 4 + 255) / 256 * 256 + 2 * (((size_t) q.total_tiles * sizeof(T) + 255) / 256 * 256);
        XTB_CUDA(cudaMallocAsync(&scratch, state, ctx->stream));
        char* s = (char*) scratch;
        q.ticket = (uint32_t*) s;
        q.status = (uint32_t*) (s + 256);
        size_t off = 256 + ((size_t) q.total_tiles * 4 + 255) / 256 * 256;
        q.aggregate = s + off;
        q.prefix = s + off + ((size_t) q.total_tiles * sizeof(T) + 255) / 256 * 256;
        q.packed = (unsigned long long*) q.aggregate;
        XTB_CUDA(cudaMemsetAsync(s, 0, state, ctx->stream));
        XTB_TRY((launch_scan<T>(q, ctx, false)));
        ScanParams r = p;
        r.tile_incl = (const char*) incl;
        k_scan_lookback<T><<<p.total_tiles, kScanThreads, 0, ctx->stream>>>(r);
        note_launch("k_scan_lookback[tile prefixes precomputed]");
        XTB_TRY(check_launch("k_scan_lookback"));
        XTB_CUDA(cudaFreeAsync(scratch, ctx->stream));
        XTB_CUDA(cudaFreeAsync(sums, ctx->stream));
        XTB_CUDA(cudaFreeAsync(incl, ctx->stream));
        return XTB_OK;
    }
    k_scan_lookback<T><<<p.total_tiles, kScanThreads, 0, ctx->stream>>>(p);
    note_launch("k_scan_lookback");
    return check_launch("k_scan_lookback");
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
        const int64_t tpr = (p.n + kScanTile - 1) / kScanTile;
        const int64_t tiles = tpr * p.rows;
        if (tiles >= 0x7fffffffLL) XTB_FAIL(XTB_ERR_UNSUPPORTED, "scan: too many tiles");
        p.tiles_per_row = (uint32_t) tpr;
        p.total_tiles = (uint32_t) tiles;
        // ticket + status + aggregate + prefix
        const size_t asz = dtype_size(acc_type);
        const size_t bytes = 256 + (size_t) tiles * 4 + 256 + 2 * ((size_t) tiles * asz + 256);
        void* scratch = nullptr;
        XTB_TRY(ensure_scratch(ctx, bytes, &scratch));
        char* s = (char*) scratch;
        p.ticket = (uint32_t*) s;
        p.status = (uint32_t*) (s + 256);
        size_t off = 256 + (((size_t) tiles * 4 + 255) / 256) * 256;
        p.aggregate = s + off;
        off += (((size_t) tiles * asz + 255) / 256) * 256;
        p.prefix = s + off;
        p.packed = (unsigned long long*) p.aggregate;   // 4-byte accumulators: 8 bytes per tile fit in aggregate+prefix
        XTB_CUDA(cudaMemsetAsync(s, 0, asz == 4 ? off + (((size_t) tiles * asz + 255) / 256) * 256 : 256 + (size_t) tiles * 4, ctx->stream));
        p.vec_io = in->dtype == acc_type && p.in_axis_stride == 1 && ((uintptr_t) p.in % 16 == 0) && ((uintptr_t) p.out % 16 == 0) &&
                   (p.rows == 1 || (p.n * asz) % 16 == 0);
        if (p.vec_io && p.rows > 1)
            for (int d = 0; d < p.n_outer; ++d) p.vec_io = p.vec_io && (p.outer_stride[d] * asz) % 16 == 0;
    } else {
        // few columns and a long axis: chunk the axis.  The per-chunk totals come from the reduction
        // kernel (one extra read of the input), their scan from this kernel on a small array.
        const int64_t cols = p.rows * p.inner;
        const int64_t target = (int64_t) ctx->sm_count * 2048;
        if (cols * 4 < target && p.n >= 256) {
            int64_t nch = std::min<int64_t>(std::max<int64_t>(target / std::max<int64_t>(cols, 1), 1), p.n / 64);
            nch = std::min<int64_t>(nch, 1024);
            if (nch > 1) {
                const int64_t chunk = (p.n + nch - 1) / nch;
                nch = (p.n + chunk - 1) / chunk;
                return scan_chunked(acc_type, p, chunk, nch, ctx);
            }
        }
    }
    switch (acc_type) {
        case XTB_I32: case XTB_U32: return launch_scan<uint32_t>(p, ctx, columns);
        case XTB_I64: case XTB_U64: return launch_scan<unsigned long long>(p, ctx, columns);
        case XTB_F32: return launch_scan<float>(p, ctx, columns);
        default: return launch_scan<double>(p, ctx, columns);
    }
}
