// xtb_ew_tma.cuh -- elementwise assignment with transposed leaves, every operand moved by TMA.
//
// Replaces the generic stepper loop for expressions like BASELINE cfg4
//     xt::noalias(out) = xt::transpose(a) + xt::view(b, xt::range(0, _, 2), xt::all())
// (stepper_assigner::run, include/xtensor/core/xassign.hpp:644-695, with a stride-N leaf walked by
// increment_stepper, core/xiterator.hpp:589-631 -- cache-hostile on the CPU).
//
// One CTA owns a TI x TJ tile of the (collapsed, rank-2) output.  One elected thread issues
//   * for every TRANSPOSED leaf (contiguous along the output's row dim i): boxes [TJ rows (j)] x [128 bytes of i] in the
//     leaf's OWN layout, 128-byte swizzled, so that the later read "column i of the box" -- 32 lanes = 32 consecutive
//     j, each a 16-byte chunk -- is free of shared-memory bank conflicts;
//   * for every DIRECT leaf (contiguous along j, like the output): one box [TI rows] x [TJ];
// (cp.async.bulk.tensor.2d -> mbarrier), all threads wait on the mbarrier, evaluate the program on 8 elements each
// (a 16-byte chunk of the transposed leaf = E consecutive i for one j; lanes run along j, so direct leaves and the
// result are accessed row-wise, conflict-free), write the result tile to shared memory, and the elected thread
// stores it with one TMA store.  No global load / store instruction is issued by the LSU (the round-1 kernel was
// bound by them: lg_throttle 2.3, short_scoreboard 2.6 per issue, 0.86 of the copy peak); out-of-range parts of
// edge tiles are zero-filled on load and clipped on store by the TMA unit.
// Several CTAs per SM (48 KB of shared memory each for one transposed + one direct leaf) overlap load, math, store.
#pragma once
#include <cuda.h>   // CUtensorMap
#include "xtb_ew.cuh"

namespace xtb {

constexpr int kTmaTI = 64;          // tile extent along i (the output's row dim = the transposed leaves' fast dim)
constexpr int kTmaTJ = 32;          // tile extent along j (the output's fast dim): one warp wide
constexpr int kTmaMaxLeaves = 3;

struct alignas(64) TmaTileParams {
    CUtensorMap leaf_map[kTmaMaxLeaves];
    CUtensorMap out_map;
    DevProgram prog;                     // the static evaluators read its immediates
    int32_t transposed[kTmaMaxLeaves];   // 1: boxes in the leaf's own (i-fast) layout
    uint32_t ntile_j;
    FastDiv div_ntj;
};

XTB_DEV uint32_t tma_smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

template <class Eval, class S>
__global__ void __launch_bounds__(256) k_ew_tile_tma(const __grid_constant__ TmaTileParams p) {
    constexpr int NL = Eval::kLeaves;
    constexpr int E = 16 / (int) sizeof(S);                 // elements per 16-byte chunk
    constexpr int IB = 128 / (int) sizeof(S);               // i-extent of one swizzled box (128-byte rows)
    constexpr int NB = kTmaTI / IB;                         // boxes per transposed leaf
    constexpr int CH = kTmaTI / E;                          // 16-byte chunks along i
    constexpr int CPW = CH / 8;                             // chunks per warp
    constexpr int U = CPW * E;                              // elements per thread (= TI / 8)
    constexpr uint32_t kTileBytes = (uint32_t) (kTmaTI * kTmaTJ * (int) sizeof(S));
    extern __shared__ unsigned char tma_smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar;
    // the 128-byte swizzle is a function of the shared-memory ADDRESS (bits 4-6 ^= bits 7-9): tiles start 1024-aligned
    unsigned char* const tma_smem = tma_smem_raw + ((1024u - (tma_smem_u32(tma_smem_raw) & 1023u)) & 1023u);
    unsigned char* const s_out = tma_smem + (size_t) NL * kTileBytes;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ti = fd_div(blockIdx.x, p.div_ntj);
    const uint32_t tj = blockIdx.x - ti * p.ntile_j;
    const int i0 = (int) (ti * kTmaTI), j0 = (int) (tj * kTmaTJ);
    const uint32_t bar = tma_smem_u32(&s_bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t) NL * kTileBytes) : "memory");
#pragma unroll
        for (int k = 0; k < NL; ++k) {
            const uint32_t dst = tma_smem_u32(tma_smem + (size_t) k * kTileBytes);
            if (p.transposed[k]) {
#pragma unroll
                for (int b = 0; b < NB; ++b)
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(dst + (uint32_t) b * (uint32_t) (kTmaTJ * 128)), "l"(&p.leaf_map[k]), "r"(i0 + b * IB), "r"(j0), "r"(bar)
                                 : "memory");
            } else {
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(dst), "l"(&p.leaf_map[k]), "r"(j0), "r"(i0), "r"(bar)
                             : "memory");
            }
        }
    }
    __syncthreads();                       // the barrier is initialised before anyone polls it
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n .reg .pred q;\n mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n selp.u32 %0, 1, 0, q;\n}"
                         : "=r"(done) : "r"(bar) : "memory");
    }
    // ---- gather this thread's U elements of every leaf: j = lane, i = (warp + 8 r) * E + e ----
    PreFetch<(NL > 0 ? NL : 1), U, S, 1> pf;
#pragma unroll
    for (int k = 0; k < NL; ++k) {
        const unsigned char* base = tma_smem + (size_t) k * kTileBytes;
        if (p.transposed[k]) {
#pragma unroll
            for (int r = 0; r < CPW; ++r) {
                const int c = warp + 8 * r;                 // chunk along i
                const int b = c >> 3, cc = c & 7;           // 8 chunks per 128-byte row
                const uint4 v = *(const uint4*) (base + (size_t) b * (kTmaTJ * 128) + (size_t) lane * 128 + (size_t) ((cc ^ (lane & 7)) << 4));
                if constexpr (sizeof(S) == 8) {
                    pf.pre[k][r * E + 0][0] = (S) (((uint64_t) v.y << 32) | v.x);
                    pf.pre[k][r * E + 1][0] = (S) (((uint64_t) v.w << 32) | v.z);
                } else {
                    pf.pre[k][r * E + 0][0] = (S) v.x;
                    pf.pre[k][r * E + 1][0] = (S) v.y;
                    pf.pre[k][r * E + 2][0] = (S) v.z;
                    pf.pre[k][r * E + 3][0] = (S) v.w;
                }
            }
        } else {
            const S* rows = (const S*) base;                // [TI][TJ]
#pragma unroll
            for (int r = 0; r < CPW; ++r)
#pragma unroll
                for (int e = 0; e < E; ++e) pf.pre[k][r * E + e][0] = rows[((warp + 8 * r) * E + e) * kTmaTJ + lane];
        }
    }
    // ---- evaluate, result tile [TI][TJ] in shared memory ----
    S* const orow = (S*) s_out;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        pf.u = u;
        S x[1];
        Eval::template run<S, 1>(p.prog, pf, x);
        orow[((warp + 8 * (u / E)) * E + (u % E)) * kTmaTJ + lane] = x[0];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the TMA unit
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                     ::"l"(&p.out_map), "r"(j0), "r"(i0), "r"(tma_smem_u32(s_out))
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory stays valid until the store has read it
    }
}

#ifndef XTB_RTC
typedef CUresult (*PFN_xtb_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_xtb_encodeTiled tma_encoder() {
    static PFN_xtb_encodeTiled encode = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            encode = (PFN_xtb_encodeTiled) fn;
    }
    return encode;
}

// Returns XTB_OK after a launch, 1 when the expression does not fit this kernel (the caller falls back to the
// register-staged tile kernel), or an error status.
template <class Eval, class S>
static int launch_ew_tile_tma(const EwParams& p, DeviceCtx* ctx, const char* evname) {
    constexpr int NL = Eval::kLeaves;
    if constexpr (NL < 1 || NL > kTmaMaxLeaves || !Eval::kPrefetch) {
        return 1;
    } else {
        if (options().no_tma || p.ndim != 2 || p.tile_i != 0 || p.n_leaves != NL) return 1;
        if (dtype_size(Eval::kResultType) != (int) sizeof(S) || p.out.dtype != Eval::kResultType) return 1;
        PFN_xtb_encodeTiled encode = tma_encoder();
        if (!encode) return 1;
        const int64_t Ni = p.shape[0], Nj = p.shape[1];
        if (Ni < kTmaTI || Nj < kTmaTJ || Ni >= (1ll << 31) || Nj >= (1ll << 31)) return 1;
        constexpr int sz = (int) sizeof(S);
        constexpr CUtensorMapDataType dt = sz == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64;   // raw words
        TmaTileParams q;
        memset(&q, 0, sizeof(q));
        q.prog = p.prog;
        const cuuint32_t estr[2] = {1u, 1u};
        auto ok_operand = [&](const char* ptr, int64_t s_fast, int64_t s_slow) {
            return s_fast == 1 && s_slow > 0 && ((uintptr_t) ptr % 16) == 0 && (s_slow * sz) % 16 == 0 && s_slow * sz < (1ll << 40);
        };
        bool any_transposed = false;
        for (int k = 0; k < NL; ++k) {
            const EwLeaf& L = p.leaf[k];
            if (dtype_size(L.dtype) != sz) return 1;
            if (L.mode == MODE_TILE) {
                // memory: i fast.  dims {Ni, Nj}, box {IB, TJ}, 128-byte swizzle
                if (!ok_operand(L.ptr, L.stride[0], L.stride[1])) return 1;
                const cuuint64_t gdim[2] = {(cuuint64_t) Ni, (cuuint64_t) Nj};
                const cuuint64_t gstr[1] = {(cuuint64_t) (L.stride[1] * sz)};
                const cuuint32_t box[2] = {(cuuint32_t) (128 / sz), (cuuint32_t) kTmaTJ};
                if (encode(&q.leaf_map[k], dt, 2, (void*) L.ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                    return 1;
                q.transposed[k] = 1;
                any_transposed = true;
            } else {
                if ((L.mode != MODE_VEC && L.mode != MODE_LINEAR) || !ok_operand(L.ptr, L.stride[1], L.stride[0])) return 1;
                const cuuint64_t gdim[2] = {(cuuint64_t) Nj, (cuuint64_t) Ni};
                const cuuint64_t gstr[1] = {(cuuint64_t) (L.stride[0] * sz)};
                const cuuint32_t box[2] = {(cuuint32_t) kTmaTJ, (cuuint32_t) kTmaTI};
                if (encode(&q.leaf_map[k], dt, 2, (void*) L.ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                    return 1;
            }
        }
        if (!any_transposed) return 1;
        {
            const EwLeaf& O = p.out;
            if (!ok_operand(O.ptr, O.stride[1], O.stride[0])) return 1;
            const cuuint64_t gdim[2] = {(cuuint64_t) Nj, (cuuint64_t) Ni};
            const cuuint64_t gstr[1] = {(cuuint64_t) (O.stride[0] * sz)};
            const cuuint32_t box[2] = {(cuuint32_t) kTmaTJ, (cuuint32_t) kTmaTI};
            if (encode(&q.out_map, dt, 2, (void*) O.ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return 1;
        }
        const int64_t nti = (Ni + kTmaTI - 1) / kTmaTI, ntj = (Nj + kTmaTJ - 1) / kTmaTJ;
        if (nti * ntj >= 0x7fffffffLL) return 1;
        q.ntile_j = (uint32_t) ntj;
        q.div_ntj = make_fastdiv(q.ntile_j);
        const size_t smem = (size_t) (NL + 1) * kTmaTI * kTmaTJ * sz + 1024;      // + alignment slack
        XTB_CUDA(cudaFuncSetAttribute(k_ew_tile_tma<Eval, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        char name[96];
        snprintf(name, sizeof(name), "k_ew_tile_tma<%s,S%d>", evname, sz * 8);
        k_ew_tile_tma<Eval, S><<<(unsigned) (nti * ntj), 256, smem, ctx->stream>>>(q);
        note_launch(name);
        return check_launch(name);
    }
}
#endif  // XTB_RTC

}  // namespace xtb
