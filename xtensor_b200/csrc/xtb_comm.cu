// xtb_comm.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 /
// NVSwitch.  The only exchange step on the path is the merge of per-GPU reduction
// partials when the reduced axis is the sharded (leading) axis -- the analogue of
// the merge step of xblockwise_reducer functors
// (include/xtensor/reducers/xblockwise_reducer_functors.hpp:45-260).
// NCCL is resolved with dlopen at xtb_comm_init so that single-GPU users do not
// need it at load time.
#include <dlfcn.h>
#include "xtb_common.hpp"
#include "xtb_ops.cuh"

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
static Nccl g_nccl;

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

int comm_allreduce(DeviceCtx* ctx, void* buf, size_t count, int dtype, int op) {
    if (g_nccl.world <= 1 && !g_nccl.comm) return XTB_OK;  // single rank: nothing to merge
    if (!g_nccl.comm) XTB_FAIL(XTB_ERR_NCCL, "xtb_comm_init has not been called");
    const int dt = nccl_dtype(dtype), ro = nccl_op(op);
    if (dt < 0 || ro < 0) XTB_FAIL(XTB_ERR_UNSUPPORTED, "allreduce of dtype %d / op %d", dtype, op);
    if (count == 0) return XTB_OK;
    XTB_NCCL(g_nccl.all_reduce(buf, buf, count, dt, ro, g_nccl.comm, ctx->stream));
    note_launch("ncclAllReduce");
    return XTB_OK;
}

}  // namespace xtb

using namespace xtb;

extern "C" {

int xtb_comm_unique_id(void* id128) {
    if (!id128) XTB_FAIL(XTB_ERR_INVALID, "null id");
    XTB_TRY(load_nccl());
    NcclUniqueId id;
    XTB_NCCL(g_nccl.get_unique_id(&id));
    memcpy(id128, &id, sizeof(id));
    return XTB_OK;
}

int xtb_comm_init(int rank, int world, const void* id128) {
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

int xtb_comm_destroy(void) {
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
