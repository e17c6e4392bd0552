// xtb_comm.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 /
// NVSwitch.  The only exchange step on the path is the merge of per-GPU reduction
// partials when the reduced axis is the sharded (leading) axis -- the analogue of
// the merge step of xblockwise_reducer functors
// (include/xtensor/reducers/xblockwise_reducer_functors.hpp:45-260).
// NCCL is resolved with dlopen at xtb_comm_init so that single-GPU users do not
// need it at load time.
//
// Small payloads (the (8192,) partials of cfg5 are 32 KB) are latency-bound in NCCL; for them
// k_allreduce_p2p does the exchange itself over NVLink peer memory: every rank owns a window
// (cudaMalloc + CUDA IPC, mapped into all ranks); a rank pushes its partial, word by word with the
// call's epoch packed next to each word, into every peer's window and then combines what the peers
// pushed into its own window in rank order -- one kernel, one one-way NVLink trip, no fences, and
// bit-identical results on all ranks.
#include <dlfcn.h>
#include <mutex>
#include "xtb_common.hpp"
#include "xtb_ops.cuh"
#include "xtb_p2p.cuh"

namespace xtb {

// minimal NCCL ABI (stable across 2.x): opaque comm, 128-byte unique id
struct NcclUniqueId { char internal[128]; };
typedef void* ncclComm_t;
typedef int (*fn_get_unique_id)(NcclUniqueId*);
typedef int (*fn_comm_init_rank)(ncclComm_t*, int, NcclUniqueId, int);
typedef int (*fn_comm_destroy)(ncclComm_t);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef const char* (*fn_get_error_string)(int);

struct Nccl {
    void* handle = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_get_error_string get_error_string = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};
// One communicator per process (the model is one process per GPU).  Bring-up, attach and destroy change this state
// and are serialised by g_comm_mutex; collectives read it and are stream ordered on the calling device.
static Nccl g_nccl;
static std::mutex g_comm_mutex;

static int load_nccl() {
    if (g_nccl.handle) return XTB_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) XTB_FAIL(XTB_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
    g_nccl.get_unique_id = (fn_get_unique_id) dlsym(h, "ncclGetUniqueId");
    g_nccl.comm_init_rank = (fn_comm_init_rank) dlsym(h, "ncclCommInitRank");
    g_nccl.comm_destroy = (fn_comm_destroy) dlsym(h, "ncclCommDestroy");
    g_nccl.all_reduce = (fn_all_reduce) dlsym(h, "ncclAllReduce");
    g_nccl.get_error_string = (fn_get_error_string) dlsym(h, "ncclGetErrorString");
    if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.comm_destroy || !g_nccl.all_reduce)
        XTB_FAIL(XTB_ERR_NCCL, "libnccl is missing required symbols");
    g_nccl.handle = h;
    return XTB_OK;
}
#define XTB_NCCL(call)                                                                                    \
    do {                                                                                                  \
        int r__ = (call);                                                                                 \
        if (r__ != 0)                                                                                     \
            XTB_FAIL(XTB_ERR_NCCL, "%s failed: %s", #call,                                                \
                     g_nccl.get_error_string ? g_nccl.get_error_string(r__) : "unknown NCCL error");      \
    } while (0)

// ncclDataType_t / ncclRedOp_t values (nccl.h): int8 0, uint8 1, int32 2, uint32 3,
// int64 4, uint64 5, float16 6, float32 7, float64 8; sum 0, prod 1, max 2, min 3
static int nccl_dtype(int dt) {
    switch (dt) {
        case XTB_I8: return 0;
        case XTB_U8: case XTB_BOOL: return 1;
        case XTB_I32: return 2;
        case XTB_U32: return 3;
        case XTB_I64: return 4;
        case XTB_U64: return 5;
        case XTB_F32: return 7;
        case XTB_F64: return 8;
        default: return -1;
    }
}
static int nccl_op(int op) {
    switch (op) {
        case XTB_RED_SUM: return 0;
        case XTB_RED_PROD: return 1;
        case XTB_RED_MAX: return 2;
        case XTB_RED_MIN: return 3;
        default: return -1;
    }
}


// ---- peer-memory allreduce (protocol: xtb_p2p.cuh) ---------------------------------------------------
struct P2p {
    bool ready = false;
    char* local = nullptr;                    // kP2pWindows windows
    char* peer[kP2pMaxWorld] = {nullptr};     // base of every rank's allocation as mapped here
};
static P2p g_p2p;

template <class T> XTB_DEV T p2p_op(int op, T a, T b) {
    switch (op) {
        case XTB_RED_SUM: return (T) (a + b);
        case XTB_RED_PROD: return (T) (a * b);
        case XTB_RED_MAX: return a > b ? a : b;          // math::maximum: (a > b) ? a : b
        case XTB_RED_NANMIN: return (a != a) ? b : ((b != b) ? a : (a < b ? a : b));   // detail::nan_min
        case XTB_RED_NANMAX: return (a != a) ? b : ((b != b) ? a : (a > b ? a : b));   // detail::nan_max
        default: return a < b ? a : b;
    }
}

// thread i owns element i
template <class T>
__global__ void __launch_bounds__(256) k_allreduce_p2p(const __grid_constant__ P2pParams w, T* buf, uint32_t count, int op) {
    constexpr int W = sizeof(T) / 4;
    const uint32_t epoch = p2p_epoch(w);
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i < count) {
        const T mine = buf[i];
        uint32_t mw[W], got[kP2pMaxWorld][W];
        memcpy(mw, &mine, sizeof(T));
        p2p_exchange<W>(w, epoch, (size_t) i * W, mw, got);
        T acc = mine;
#pragma unroll
        for (int r = 0; r < kP2pMaxWorld; ++r) {
            if (r < w.world) {
                T v = mine;
                if (r != w.rank) memcpy(&v, got[r], sizeof(T));
                acc = r == 0 ? v : p2p_op<T>(op, acc, v);
            }
        }
        buf[i] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) p2p_finish(w, epoch);
}

// the windows an exchange issued now (on ctx's current stream) uses; false when peer memory is not attached
bool comm_p2p_params(DeviceCtx* ctx, P2pParams* w) {
    if (!g_p2p.ready) return false;
    memset(w, 0, sizeof(*w));
    w->rank = g_nccl.rank;
    w->world = g_nccl.world;
    const size_t woff = ctx->forked ? kP2pWindow : 0;
    for (int r = 0; r < w->world; ++r) w->win[r] = g_p2p.peer[r] + woff;
    return true;
}

template <class T> static int launch_p2p(DeviceCtx* ctx, void* buf, size_t count, int op) {
    P2pParams w;
    comm_p2p_params(ctx, &w);
    k_allreduce_p2p<T><<<(unsigned) ((count + 255) / 256), 256, 0, ctx->stream>>>(w, (T*) buf, (uint32_t) count, op);
    note_launch("k_allreduce_p2p");
    return check_launch("k_allreduce_p2p");
}

int comm_world() { return g_nccl.world; }

int comm_allreduce(DeviceCtx* ctx, void* buf, size_t count, int dtype, int op) {
    if (g_nccl.world <= 1 && !g_nccl.comm) return XTB_OK;  // single rank: nothing to merge
    if (!g_nccl.comm) XTB_FAIL(XTB_ERR_NCCL, "xtb_comm_init has not been called");
    const int dt = nccl_dtype(dtype), ro = nccl_op(op);
    if (dt < 0 || op < XTB_RED_SUM || op > XTB_RED_NANMAX) XTB_FAIL(XTB_ERR_UNSUPPORTED, "allreduce of dtype %d / op %d", dtype, op);
    if (count == 0) return XTB_OK;
    if (g_p2p.ready && count * (size_t) dtype_size(dtype) <= kP2pMaxBytes && dtype >= XTB_I32) {
        switch (dtype) {
            case XTB_I32: return launch_p2p<int32_t>(ctx, buf, count, op);
            case XTB_U32: return launch_p2p<uint32_t>(ctx, buf, count, op);
            case XTB_I64: return launch_p2p<long long>(ctx, buf, count, op);
            case XTB_U64: return launch_p2p<unsigned long long>(ctx, buf, count, op);
            case XTB_F32: return launch_p2p<float>(ctx, buf, count, op);
            default: return launch_p2p<double>(ctx, buf, count, op);
        }
    }
    // NCCL has no NaN-skipping extremes: nan_min / nan_max merge through the peer-memory kernel only
    if (ro < 0) XTB_FAIL(XTB_ERR_UNSUPPORTED, "allreduce op %d needs the peer-memory route (payload <= %d KB, xtb_comm_p2p_attach)", op, (int) (kP2pMaxBytes >> 10));
    XTB_NCCL(g_nccl.all_reduce(buf, buf, count, dt, ro, g_nccl.comm, ctx->stream));
    note_launch("ncclAllReduce");
    return XTB_OK;
}

}  // namespace xtb

using namespace xtb;

extern "C" {

int xtb_comm_unique_id(void* id128) {
    std::lock_guard<std::mutex> comm_lock(g_comm_mutex);
    if (!id128) XTB_FAIL(XTB_ERR_INVALID, "null id");
    XTB_TRY(load_nccl());
    NcclUniqueId id;
    XTB_NCCL(g_nccl.get_unique_id(&id));
    memcpy(id128, &id, sizeof(id));
    return XTB_OK;
}

int xtb_comm_init(int rank, int world, const void* id128) {
    std::lock_guard<std::mutex> comm_lock(g_comm_mutex);
    if (world < 1 || rank < 0 || rank >= world) XTB_FAIL(XTB_ERR_INVALID, "bad rank %d / world %d", rank, world);
    if (g_nccl.comm) XTB_FAIL(XTB_ERR_INVALID, "communicator already initialised");
    DeviceCtx* ctx;
    XTB_TRY(get_ctx(&ctx));
    if (world == 1) {
        g_nccl.rank = 0;
        g_nccl.world = 1;
        return XTB_OK;
    }
    if (!id128) XTB_FAIL(XTB_ERR_INVALID, "null id");
    XTB_TRY(load_nccl());
    NcclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    XTB_NCCL(g_nccl.comm_init_rank(&g_nccl.comm, world, id, rank));
    g_nccl.rank = rank;
    g_nccl.world = world;
    return XTB_OK;
}

int xtb_comm_p2p_handle(void* handle64) {
    std::lock_guard<std::mutex> comm_lock(g_comm_mutex);
    if (!handle64) XTB_FAIL(XTB_ERR_INVALID, "null handle");
    DeviceCtx* ctx;
    XTB_TRY(get_ctx(&ctx));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    if (!g_p2p.local) {
        XTB_CUDA(cudaMalloc((void**) &g_p2p.local, kP2pWindows * kP2pWindow));
        XTB_CUDA(cudaMemset(g_p2p.local, 0, kP2pWindows * kP2pWindow));
    }
    cudaIpcMemHandle_t h;
    XTB_CUDA(cudaIpcGetMemHandle(&h, g_p2p.local));
    memcpy(handle64, &h, sizeof(h));
    return XTB_OK;
}

int xtb_comm_p2p_attach(const void* handles, int world) {
    std::lock_guard<std::mutex> comm_lock(g_comm_mutex);
    if (!handles && world == 0) {   // detach: back to NCCL for every payload (all ranks must do the same)
        g_p2p.ready = false;
        return XTB_OK;
    }
    if (!handles) XTB_FAIL(XTB_ERR_INVALID, "null handles");
    if (world != g_nccl.world || world > kP2pMaxWorld || !g_p2p.local) XTB_FAIL(XTB_ERR_INVALID, "xtb_comm_p2p_attach: call xtb_comm_init and xtb_comm_p2p_handle first (world <= %d)", kP2pMaxWorld);
    DeviceCtx* ctx;
    XTB_TRY(get_ctx(&ctx));
    for (int r = 0; r < world; ++r) {
        if (r == g_nccl.rank) {
            g_p2p.peer[r] = g_p2p.local;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*) handles + (size_t) r * sizeof(h), sizeof(h));
        void* ptr = nullptr;
        XTB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        g_p2p.peer[r] = (char*) ptr;
    }
    g_p2p.ready = true;
    return XTB_OK;
}

int xtb_comm_destroy(void) {
    std::lock_guard<std::mutex> comm_lock(g_comm_mutex);
    // Every rank must have finished its last exchange before any rank frees its window (peers push into it):
    // callers destroy collectively, after a barrier of their own or -- as here, when NCCL is up -- after one
    // last collective on the communicator that is about to go away.
    if (g_p2p.local) {
        DeviceCtx* ctx = nullptr;
        if (g_nccl.comm && g_nccl.world > 1 && get_ctx(&ctx) == XTB_OK) {
            g_nccl.all_reduce(g_p2p.local, g_p2p.local, 1, nccl_dtype(XTB_I32), 0, g_nccl.comm, ctx->stream);   // barrier
        }
        cudaDeviceSynchronize();
        // mappings opened by an attach stay open across a detach (ready == false): close whatever is mapped
        for (int r = 0; r < kP2pMaxWorld; ++r)
            if (r != g_nccl.rank && g_p2p.peer[r] && g_p2p.peer[r] != g_p2p.local) cudaIpcCloseMemHandle(g_p2p.peer[r]);
        cudaFree(g_p2p.local);
        g_p2p = P2p();
    }
    if (g_nccl.comm) {
        XTB_NCCL(g_nccl.comm_destroy(g_nccl.comm));
        g_nccl.comm = nullptr;
    }
    g_nccl.rank = 0;
    g_nccl.world = 1;
    return XTB_OK;
}

int xtb_comm_info(int* rank, int* world) {
    if (rank) *rank = g_nccl.rank;
    if (world) *world = g_nccl.world;
    return XTB_OK;
}

int xtb_allreduce(void* buf, size_t count, int dtype, int op) {
    DeviceCtx* ctx;
    XTB_TRY(get_ctx(&ctx));
    return comm_allreduce(ctx, buf, count, dtype, op);
}

}  // extern "C"
