// xtb_p2p.cuh -- device side of the NVLink peer-memory exchange (windows are set up in xtb_comm.cu).
//
// Push protocol with the flag inside the data (the idea of NCCL's LL protocol): the payload travels as
// 32-bit words, each packed with the call's epoch into ONE 64-bit store (single-copy atomic), so a
// reader that sees the epoch also sees the word -- no fence, no separate flag, no block- or grid-wide
// hand-shake.  The thread that owns word i stores it into slot[epoch & 1][my rank][i] of every peer's
// window (posted NVLink writes), then polls slot[epoch & 1][r][i] of its OWN window (local L2) for
// every other rank r; the caller combines the R values in rank order -- identical bits on every rank.
// Slot reuse: call k+1 of a rank starts after its call k-1 kernel completed (stream order), and a
// peer can only finish call k -- hence start k+1 and overwrite parity (k-1)&1 -- once this rank's
// call-k words have arrived, i.e. after this rank is done with call k-1.  The epoch lives in the
// window (not a kernel argument) so that a captured CUDA graph can be replayed.
#pragma once
#include "xtb_ops.cuh"

namespace xtb {

constexpr size_t kP2pMaxBytes = 256 * 1024;
constexpr uint32_t kP2pMaxWords = kP2pMaxBytes / 4;
constexpr int kP2pMaxWorld = 8;
constexpr size_t kP2pHeader = 4096;                           // u32 epoch counter at 0, CTA ticket at 4
constexpr size_t kP2pSlotBytes = (size_t) kP2pMaxWorld * kP2pMaxWords * 8;
constexpr size_t kP2pWindow = kP2pHeader + 2 * kP2pSlotBytes;  // two slots (epoch parity)
constexpr int kP2pWindows = 2;                                // [1] serves calls made inside xtb_fork_begin/end

struct P2pParams {
    char* win[kP2pMaxWorld];   // this call's window of every rank, as mapped into this process
    int32_t rank, world;
};

XTB_DEV void p2p_st_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
XTB_DEV unsigned long long p2p_ld_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// the epoch of this launch; every thread that exchanges reads it before p2p_finish of its block
XTB_DEV uint32_t p2p_epoch(const P2pParams& w) { return *(volatile uint32_t*) w.win[w.rank] + 1; }

// Exchange W consecutive words starting at word index i0: got[r][0..W) = rank r's words (r != rank untouched
// for r == rank: the caller uses its own registers).  Every rank's words are requested together; the
// peers' stores land at about the same time.
template <int W>
XTB_DEV void p2p_exchange(const P2pParams& w, uint32_t epoch, size_t i0, const uint32_t (&mine)[W], uint32_t (&got)[kP2pMaxWorld][W]) {
    const size_t slot = kP2pHeader + (size_t) (epoch & 1u) * kP2pSlotBytes;
#pragma unroll
    for (int r = 0; r < kP2pMaxWorld; ++r) {
        if (r < w.world && r != w.rank) {
            unsigned long long* dst = (unsigned long long*) (w.win[r] + slot) + (size_t) w.rank * kP2pMaxWords + i0;
#pragma unroll
            for (int k = 0; k < W; ++k) p2p_st_u64(dst + k, (unsigned long long) mine[k] | ((unsigned long long) epoch << 32));
        }
    }
    const unsigned long long* src = (const unsigned long long*) (w.win[w.rank] + slot) + i0;
    bool all;
    do {
        all = true;
#pragma unroll
        for (int r = 0; r < kP2pMaxWorld; ++r) {
            if (r < w.world && r != w.rank) {
#pragma unroll
                for (int k = 0; k < W; ++k) {
                    const unsigned long long v = p2p_ld_u64(src + (size_t) r * kP2pMaxWords + k);
                    got[r][k] = (uint32_t) v;
                    all = all && (uint32_t) (v >> 32) == epoch;
                }
            }
        }
    } while (!all);
}

// Called by ONE thread per block after the block's exchanges (behind a __syncthreads): the last block
// of the launch advances the epoch -- by then every block has read it.
XTB_DEV void p2p_finish(const P2pParams& w, uint32_t epoch) {
    uint32_t* hdr = (uint32_t*) w.win[w.rank];
    const uint32_t t = atomicAdd(hdr + 1, 1u);
    if (t == gridDim.x * gridDim.y * gridDim.z - 1) {
        *(volatile uint32_t*) (hdr + 1) = 0;
        *(volatile uint32_t*) hdr = epoch;
    }
}

}  // namespace xtb
