// xtb_runtime.cu -- runtime of libxtb200: device contexts and streams, the
// device-resident storage behind xtb::device_uvector (uvector contract of
// include/xtensor/containers/xstorage.hpp:33-345: uninitialised contents,
// resize discards), stream-ordered copies, program validation and the
// operand canonicaliser used by every compute entry point.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <vector>
#include "xtb_common.hpp"
#include "xtb_ops.cuh"

namespace xtb {

// ---- error state -------------------------------------------------------------
static thread_local char g_err[512] = "";
static thread_local char g_last_kernel[128] = "";
static std::atomic<int64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void note_launch(const char* kernel_name, int n) {
    g_launches.fetch_add(n, std::memory_order_relaxed);
    if (kernel_name) {
        strncpy(g_last_kernel, kernel_name, sizeof(g_last_kernel) - 1);
        g_last_kernel[sizeof(g_last_kernel) - 1] = 0;
    }
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(XTB_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return XTB_OK;
}

// ---- process options -----------------------------------------------------------
static Options g_options;
static std::once_flag g_options_once;
Options& options() {
    std::call_once(g_options_once, [] {
        auto flag = [](const char* n) { const char* e = getenv(n); return e != nullptr && e[0] != 0 && strcmp(e, "0") != 0; };
        g_options.no_static = flag("XTB_NO_STATIC");
        g_options.no_jit = flag("XTB_NO_JIT");
        g_options.no_staged = flag("XTB_NO_STAGED");
        g_options.no_tma = flag("XTB_NO_TMA");
        g_options.jit_verbose = flag("XTB_JIT_VERBOSE");
        if (const char* e = getenv("XTB_JIT_MIN_ELEMS")) g_options.jit_min_elems = atoll(e);
        if (const char* e = getenv("XTB_SCAN_VARIANT")) g_options.scan_variant = atoi(e);
        if (const char* e = getenv("XTB_TILE_VARIANT")) g_options.tile_variant = atoi(e);
        if (const char* e = getenv("XTB_SCAN_NV")) g_options.scan_nv = atoi(e);
        if (const char* e = getenv("XTB_ARG_TWO_PASS")) g_options.arg_two_pass = atoi(e);
        g_options.no_pdl = flag("XTB_NO_PDL");
        g_options.no_decompose = flag("XTB_NO_DECOMPOSE");
        if (const char* e = getenv("XTB_REDUCE_G")) g_options.reduce_g = atoi(e);
        if (const char* e = getenv("XTB_REDUCE_SPLIT")) g_options.reduce_split = atoi(e);
    });
    return g_options;
}

// ---- device contexts ---------------------------------------------------------
static constexpr int kMaxDevices = 16;
static DeviceCtx g_ctx[kMaxDevices];
static std::mutex g_ctx_mutex;
static thread_local int t_device = -1;
static int g_device_count = -2;  // -2 = not probed

static int probe_devices() {
    std::lock_guard<std::mutex> lock(g_ctx_mutex);
    if (g_device_count == -2) {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess) {
            (void) cudaGetLastError();
            n = 0;
        }
        g_device_count = n;
    }
    return g_device_count;
}

static int init_device(int device) {
    int n = probe_devices();
    if (n <= 0)
        XTB_FAIL(XTB_ERR_NO_DEVICE,
                 "no CUDA device available: libxtb200 has no CPU fallback (hot path is sm_100a only)");
    if (device < 0) {
        if (t_device >= 0) device = t_device;
        else {
            int cur = 0;
            XTB_CUDA(cudaGetDevice(&cur));
            device = cur;
        }
    }
    if (device >= n || device >= kMaxDevices) XTB_FAIL(XTB_ERR_INVALID, "device %d out of range (%d devices)", device, n);
    XTB_CUDA(cudaSetDevice(device));
    std::lock_guard<std::mutex> lock(g_ctx_mutex);
    DeviceCtx& c = g_ctx[device];
    if (!c.ready) {
        c.device = device;
        cudaDeviceProp prop;
        XTB_CUDA(cudaGetDeviceProperties(&prop, device));
        c.sm_count = prop.multiProcessorCount;
        c.l2_bytes = (size_t) prop.l2CacheSize;
        XTB_CUDA(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
        c.stream = c.own_stream;
        // caching, stream-ordered allocator: keep freed blocks in the pool
        cudaMemPool_t pool;
        XTB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t threshold = UINT64_MAX;
        XTB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
        c.ready = true;
    }
    t_device = device;
    return XTB_OK;
}

int get_ctx(DeviceCtx** ctx) {
    if (t_device < 0 || !g_ctx[t_device].ready) XTB_TRY(init_device(-1));
    else {
        int cur = -1;
        cudaGetDevice(&cur);
        if (cur != t_device) XTB_CUDA(cudaSetDevice(t_device));
    }
    *ctx = &g_ctx[t_device];
    return XTB_OK;
}

int ensure_scratch(DeviceCtx* ctx, size_t bytes, void** ptr) {
    void*& buf = ctx->forked ? ctx->fork_scratch : ctx->scratch;
    size_t& have = ctx->forked ? ctx->fork_scratch_bytes : ctx->scratch_bytes;
    if (!ctx->forked) {
        // the buffer is reused call after call: stream order protects it on one stream; after xtb_set_stream
        // moved the library to another stream, the new stream first waits for the old one's last user
        if (ctx->scratch_stream && ctx->scratch_stream != ctx->stream) {
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            cudaStreamIsCapturing(ctx->stream, &cs);
            if (cs == cudaStreamCaptureStatusNone) {
                if (!ctx->scratch_ev) XTB_CUDA(cudaEventCreateWithFlags(&ctx->scratch_ev, cudaEventDisableTiming));
                XTB_CUDA(cudaEventRecord(ctx->scratch_ev, ctx->scratch_stream));
                XTB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->scratch_ev, 0));
            }
        }
        ctx->scratch_stream = ctx->stream;
    }
    if (bytes > have) {
        if (buf) XTB_CUDA(cudaFreeAsync(buf, ctx->stream));
        size_t want = std::max(bytes, (size_t) 1 << 20);
        buf = nullptr;
        have = 0;
        XTB_CUDA(cudaMallocAsync(&buf, want, ctx->stream));
        have = want;
    }
    *ptr = buf;
    return XTB_OK;
}

// ---- canonicaliser -----------------------------------------------------------
int align_operand(const xtb_operand* op, int ndim, const int64_t* shape, int64_t* stride_out, const char* what) {
    if (op->ndim < 0 || op->ndim > XTB_MAX_DIM) XTB_FAIL(XTB_ERR_INVALID, "%s: rank %d out of range", what, op->ndim);
    if (op->ndim > ndim) XTB_FAIL(XTB_ERR_SHAPE, "%s: rank %d exceeds iteration rank %d", what, op->ndim, ndim);
    if (op->dtype < 0 || op->dtype >= XTB_DTYPE_COUNT) XTB_FAIL(XTB_ERR_INVALID, "%s: bad dtype %d", what, op->dtype);
    const int shift = ndim - op->ndim;
    for (int d = 0; d < ndim; ++d) {
        if (d < shift) {
            stride_out[d] = 0;
            continue;
        }
        const int64_t ext = op->shape[d - shift];
        if (ext == shape[d]) stride_out[d] = (ext == 1) ? 0 : op->stride[d - shift];
        else if (ext == 1) stride_out[d] = 0;
        else
            XTB_FAIL(XTB_ERR_SHAPE, "%s: extent %lld of dim %d is not broadcastable to %lld", what, (long long) ext,
                     d - shift, (long long) shape[d]);
    }
    return XTB_OK;
}

void collapse_space(Space* s) {
    // drop extent-1 dims
    int w = 0;
    for (int d = 0; d < s->ndim; ++d) {
        if (s->shape[d] == 1) continue;
        s->shape[w] = s->shape[d];
        s->reduced[w] = s->reduced[d];
        for (int k = 0; k < s->n_ops; ++k) s->stride[k][w] = s->stride[k][d];
        ++w;
    }
    s->ndim = w;
    // merge (d, d+1) when stride[d] == shape[d+1] * stride[d+1] for every operand
    int o = 0;
    for (int d = 1; d < s->ndim; ++d) {
        bool ok = s->reduced[o] == s->reduced[d];
        for (int k = 0; k < s->n_ops && ok; ++k) ok = s->stride[k][o] == s->shape[d] * s->stride[k][d];
        if (ok) {
            s->shape[o] *= s->shape[d];
            for (int k = 0; k < s->n_ops; ++k) s->stride[k][o] = s->stride[k][d];
        } else {
            ++o;
            s->shape[o] = s->shape[d];
            s->reduced[o] = s->reduced[d];
            for (int k = 0; k < s->n_ops; ++k) s->stride[k][o] = s->stride[k][d];
        }
    }
    if (s->ndim > 0) s->ndim = o + 1;
    s->total = 1;
    for (int d = 0; d < s->ndim; ++d) s->total *= s->shape[d];
}

void sort_space_by(Space* s, int key) {
    // stable insertion sort on |stride[key]| descending
    for (int i = 1; i < s->ndim; ++i) {
        for (int j = i; j > 0; --j) {
            int64_t a = s->stride[key][j - 1], b = s->stride[key][j];
            if (a < 0) a = -a;
            if (b < 0) b = -b;
            if (a >= b) break;
            std::swap(s->shape[j - 1], s->shape[j]);
            std::swap(s->reduced[j - 1], s->reduced[j]);
            for (int k = 0; k < s->n_ops; ++k) std::swap(s->stride[k][j - 1], s->stride[k][j]);
        }
    }
}

// ---- program validation --------------------------------------------------------
static bool is_reg_type(int t) { return t >= XTB_I32 && t <= XTB_F64; }
static bool is64(int t) { return t == XTB_I64 || t == XTB_U64 || t == XTB_F64; }
static bool float_only_unary(int op) {
    return (op >= XTB_OP_EXP && op <= XTB_OP_RINT) || op == XTB_OP_DEG2RAD || op == XTB_OP_RAD2DEG;
}
static bool float_only_binary(int op) { return op >= XTB_OP_FMOD && op <= XTB_OP_ATAN2; }
static bool int_only_binary(int op) {
    return op == XTB_OP_MOD || (op >= XTB_OP_BOR && op <= XTB_OP_SHR);
}

int validate_program(const xtb_program* p, const int32_t* leaf_dtypes, int* result_type, bool* needs64) {
    if (!p) XTB_FAIL(XTB_ERR_INVALID, "null program");
    if (p->n_insns <= 0 || p->n_insns > XTB_MAX_INSNS) XTB_FAIL(XTB_ERR_INVALID, "program has %d instructions", p->n_insns);
    if (p->n_leaves < 0 || p->n_leaves > XTB_MAX_LEAVES) XTB_FAIL(XTB_ERR_INVALID, "program has %d leaves", p->n_leaves);
    if (p->n_imms < 0 || p->n_imms > XTB_MAX_IMMS) XTB_FAIL(XTB_ERR_INVALID, "program has %d immediates", p->n_imms);
    int st[XTB_MAX_STACK];
    int n = 0;
    bool w64 = false;
    for (int pc = 0; pc < p->n_insns; ++pc) {
        const xtb_insn in = p->insns[pc];
        const int op = in.op;
        if (op == XTB_OP_PUSH) {
            if (n >= XTB_MAX_STACK) XTB_FAIL(XTB_ERR_INVALID, "insn %d: stack overflow", pc);
            if (in.src == XTB_SRC_LEAF) {
                if (in.arg >= p->n_leaves) XTB_FAIL(XTB_ERR_INVALID, "insn %d: leaf %d out of range", pc, in.arg);
                if (in.type >= XTB_DTYPE_COUNT) XTB_FAIL(XTB_ERR_INVALID, "insn %d: bad leaf dtype", pc);
                if (leaf_dtypes && leaf_dtypes[in.arg] != in.type)
                    XTB_FAIL(XTB_ERR_INVALID, "insn %d: leaf %d has dtype %d, program says %d", pc, in.arg,
                             leaf_dtypes[in.arg], in.type);
                st[n++] = regtype_of(in.type);
            } else if (in.src == XTB_SRC_IMM) {
                if (in.arg >= p->n_imms) XTB_FAIL(XTB_ERR_INVALID, "insn %d: imm %d out of range", pc, in.arg);
                if (!is_reg_type(in.type)) XTB_FAIL(XTB_ERR_INVALID, "insn %d: imm must have a register type", pc);
                st[n++] = in.type;
            } else
                XTB_FAIL(XTB_ERR_INVALID, "insn %d: bad PUSH source", pc);
            w64 |= is64(st[n - 1]);
            continue;
        }
        if (!is_reg_type(in.type)) XTB_FAIL(XTB_ERR_INVALID, "insn %d: type %d is not a register type", pc, in.type);
        w64 |= is64(in.type);
        if (op < XTB_OP_ADD) {  // unary
            if (op < XTB_OP_CAST || op > XTB_OP_ORDKEY) XTB_FAIL(XTB_ERR_INVALID, "insn %d: unknown opcode %d", pc, op);
            if (n < 1) XTB_FAIL(XTB_ERR_INVALID, "insn %d: stack underflow", pc);
            if (st[n - 1] != in.type)
                XTB_FAIL(XTB_ERR_INVALID, "insn %d: operand has type %d, insn says %d", pc, st[n - 1], in.type);
            if (op == XTB_OP_CAST) {
                if (in.arg >= XTB_DTYPE_COUNT) XTB_FAIL(XTB_ERR_INVALID, "insn %d: bad cast target", pc);
                st[n - 1] = regtype_of(in.arg);
                w64 |= is64(st[n - 1]);
            } else if (op == XTB_OP_ORDKEY) {
                if (in.type != XTB_F32 && in.type != XTB_I32 && in.type != XTB_U32) XTB_FAIL(XTB_ERR_INVALID, "insn %d: ORDKEY takes a 32-bit value", pc);
                if (in.arg > 1) XTB_FAIL(XTB_ERR_INVALID, "insn %d: ORDKEY arg is 0 (min) or 1 (max)", pc);
                st[n - 1] = XTB_U64;
                w64 = true;
            } else if (is_pred_op(op)) {
                st[n - 1] = XTB_I32;
            } else {
                if (float_only_unary(op) && !is_float_type(in.type))
                    XTB_FAIL(XTB_ERR_INVALID, "insn %d: opcode %d needs a floating type (lowering must cast)", pc, op);
                if (op == XTB_OP_BITNOT && is_float_type(in.type)) XTB_FAIL(XTB_ERR_INVALID, "insn %d: ~ on float", pc);
            }
        } else if (op < XTB_OP_WHERE) {  // binary
            if (op > XTB_OP_NANMAX) XTB_FAIL(XTB_ERR_INVALID, "insn %d: unknown opcode %d", pc, op);
            const int kind = in.src & 3;
            int ty;
            if (kind == XTB_SRC_STACK) {
                if (n < 2) XTB_FAIL(XTB_ERR_INVALID, "insn %d: stack underflow", pc);
                ty = st[n - 1];
                --n;
            } else if (kind == XTB_SRC_LEAF) {
                if (in.arg >= p->n_leaves) XTB_FAIL(XTB_ERR_INVALID, "insn %d: leaf %d out of range", pc, in.arg);
                if (!leaf_dtypes) XTB_FAIL(XTB_ERR_INVALID, "insn %d: fused leaf operand needs leaf dtypes", pc);
                if (n < 1) XTB_FAIL(XTB_ERR_INVALID, "insn %d: stack underflow", pc);
                if (leaf_dtypes[in.arg] != in.type)
                    XTB_FAIL(XTB_ERR_INVALID, "insn %d: fused leaf %d must be stored as the insn type", pc, in.arg);
                ty = in.type;
            } else if (kind == XTB_SRC_IMM) {
                if (in.arg >= p->n_imms) XTB_FAIL(XTB_ERR_INVALID, "insn %d: imm %d out of range", pc, in.arg);
                if (n < 1) XTB_FAIL(XTB_ERR_INVALID, "insn %d: stack underflow", pc);
                ty = in.type;
            } else
                XTB_FAIL(XTB_ERR_INVALID, "insn %d: bad operand source", pc);
            if (st[n - 1] != in.type || ty != in.type)
                XTB_FAIL(XTB_ERR_INVALID, "insn %d: operand types (%d,%d) differ from insn type %d", pc, st[n - 1], ty, in.type);
            if (float_only_binary(op) && !is_float_type(in.type)) XTB_FAIL(XTB_ERR_INVALID, "insn %d: needs floating type", pc);
            if (int_only_binary(op) && is_float_type(in.type)) XTB_FAIL(XTB_ERR_INVALID, "insn %d: needs integer type", pc);
            if (is_cmp_op(op)) st[n - 1] = XTB_I32;
        } else {  // ternary
            if (op > XTB_OP_CLAMP) XTB_FAIL(XTB_ERR_INVALID, "insn %d: unknown opcode %d", pc, op);
            if (n < 3) XTB_FAIL(XTB_ERR_INVALID, "insn %d: stack underflow", pc);
            if (st[n - 1] != in.type || st[n - 2] != in.type) XTB_FAIL(XTB_ERR_INVALID, "insn %d: operand types differ", pc);
            if (op == XTB_OP_WHERE) {
                if (st[n - 3] != XTB_I32) XTB_FAIL(XTB_ERR_INVALID, "insn %d: condition must be bool/int", pc);
            } else if (st[n - 3] != in.type)
                XTB_FAIL(XTB_ERR_INVALID, "insn %d: operand types differ", pc);
            n -= 2;
            st[n - 1] = in.type;
        }
    }
    if (n != 1) XTB_FAIL(XTB_ERR_INVALID, "program leaves %d values on the stack", n);
    if (result_type) *result_type = st[0];
    if (needs64) *needs64 = w64;
    return XTB_OK;
}

}  // namespace xtb

using namespace xtb;

extern "C" {

int xtb_abi_version(void) { return XTB_ABI_VERSION; }

int xtb_init(int device) { return init_device(device); }

int xtb_device_count(int* count) {
    if (!count) XTB_FAIL(XTB_ERR_INVALID, "null count");
    *count = std::max(0, probe_devices());
    return XTB_OK;
}

int xtb_sync(void) {
    DeviceCtx* c;
    XTB_TRY(get_ctx(&c));
    XTB_CUDA(cudaStreamSynchronize(c->stream));
    return XTB_OK;
}

const char* xtb_last_error(void) { return g_err; }

int xtb_set_stream(void* cuda_stream) {
    DeviceCtx* c;
    XTB_TRY(get_ctx(&c));
    c->stream = cuda_stream ? (cudaStream_t) cuda_stream : c->own_stream;
    return XTB_OK;
}

void* xtb_get_stream(void) {
    DeviceCtx* c;
    if (get_ctx(&c) != XTB_OK) return nullptr;
    return (void*) c->stream;
}

int xtb_malloc(size_t bytes, void** ptr) {
    if (!ptr) XTB_FAIL(XTB_ERR_INVALID, "null ptr");
    *ptr = nullptr;
    DeviceCtx* c;
    XTB_TRY(get_ctx(&c));
    if (bytes == 0) return XTB_OK;
    XTB_CUDA(cudaMallocAsync(ptr, bytes, c->stream));
    return XTB_OK;
}

int xtb_free(void* ptr) {
    if (!ptr) return XTB_OK;
    DeviceCtx* c;
    XTB_TRY(get_ctx(&c));
    XTB_CUDA(cudaFreeAsync(ptr, c->stream));
    return XTB_OK;
}

int xtb_memcpy(void* dst, const void* src, size_t bytes, int kind) {
    DeviceCtx* c;
    XTB_TRY(get_ctx(&c));
    if (bytes == 0) return XTB_OK;
    cudaMemcpyKind k;
    switch (kind) {
        case XTB_H2D: k = cudaMemcpyHostToDevice; break;
        case XTB_D2H: k = cudaMemcpyDeviceToHost; break;
        case XTB_D2D: k = cudaMemcpyDeviceToDevice; break;
        default: XTB_FAIL(XTB_ERR_INVALID, "bad copy kind %d", kind);
    }
    XTB_CUDA(cudaMemcpyAsync(dst, src, bytes, k, c->stream));
    if (kind == XTB_D2H) XTB_CUDA(cudaStreamSynchronize(c->stream));
    return XTB_OK;
}

int xtb_memset(void* dst, int byte, size_t bytes) {
    DeviceCtx* c;
    XTB_TRY(get_ctx(&c));
    if (bytes == 0) return XTB_OK;
    XTB_CUDA(cudaMemsetAsync(dst, byte, bytes, c->stream));
    return XTB_OK;
}

int xtb_host_alloc(size_t bytes, void** ptr) {
    if (!ptr) XTB_FAIL(XTB_ERR_INVALID, "null ptr");
    DeviceCtx* c;
    XTB_TRY(get_ctx(&c));
    XTB_CUDA(cudaMallocHost(ptr, bytes ? bytes : 1));
    return XTB_OK;
}

int xtb_host_free(void* ptr) {
    if (!ptr) return XTB_OK;
    XTB_CUDA(cudaFreeHost(ptr));
    return XTB_OK;
}

int xtb_event_create(void** event) {
    if (!event) XTB_FAIL(XTB_ERR_INVALID, "null event");
    DeviceCtx* c;
    XTB_TRY(get_ctx(&c));
    cudaEvent_t e;
    XTB_CUDA(cudaEventCreate(&e));
    *event = (void*) e;
    return XTB_OK;
}

int xtb_event_record(void* event) {
    DeviceCtx* c;
    XTB_TRY(get_ctx(&c));
    XTB_CUDA(cudaEventRecord((cudaEvent_t) event, c->stream));
    return XTB_OK;
}

int xtb_event_elapsed_ms(void* start, void* stop, float* ms) {
    if (!ms) XTB_FAIL(XTB_ERR_INVALID, "null ms");
    XTB_CUDA(cudaEventSynchronize((cudaEvent_t) stop));
    XTB_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t) start, (cudaEvent_t) stop));
    return XTB_OK;
}

int xtb_event_destroy(void* event) {
    if (event) XTB_CUDA(cudaEventDestroy((cudaEvent_t) event));
    return XTB_OK;
}

int xtb_fork_begin(void) {
    DeviceCtx* c;
    XTB_TRY(get_ctx(&c));
    if (c->forked) XTB_FAIL(XTB_ERR_INVALID, "xtb_fork_begin: already forked");
    if (!c->fork_stream) {
        // highest priority: the forked section is the short side chain (reduce, merge, exchange, finalize);
        // its blocks are placed ahead of the main sequence's bulk kernel so that the exchange overlaps it
        int prio_lo = 0, prio_hi = 0;
        XTB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        XTB_CUDA(cudaStreamCreateWithPriority(&c->fork_stream, cudaStreamNonBlocking, prio_hi));
        XTB_CUDA(cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming));
        XTB_CUDA(cudaEventCreateWithFlags(&c->join_ev, cudaEventDisableTiming));
    }
    XTB_CUDA(cudaEventRecord(c->fork_ev, c->stream));
    XTB_CUDA(cudaStreamWaitEvent(c->fork_stream, c->fork_ev, 0));
    c->fork_saved = c->stream;
    c->stream = c->fork_stream;
    c->forked = true;
    return XTB_OK;
}

int xtb_fork_end(void) {
    DeviceCtx* c;
    XTB_TRY(get_ctx(&c));
    if (!c->forked) XTB_FAIL(XTB_ERR_INVALID, "xtb_fork_end without xtb_fork_begin");
    XTB_CUDA(cudaEventRecord(c->join_ev, c->fork_stream));
    c->stream = c->fork_saved;
    c->forked = false;
    return XTB_OK;
}

int xtb_fork_join(void) {
    DeviceCtx* c;
    XTB_TRY(get_ctx(&c));
    if (c->forked || !c->join_ev) XTB_FAIL(XTB_ERR_INVALID, "xtb_fork_join: no finished fork section");
    XTB_CUDA(cudaStreamWaitEvent(c->stream, c->join_ev, 0));
    return XTB_OK;
}

int xtb_graph_begin(void) {
    DeviceCtx* c;
    XTB_TRY(get_ctx(&c));
    XTB_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    return XTB_OK;
}

static std::mutex g_graph_mutex;
static std::unordered_map<void*, int> g_graph_kernels;   // kernel nodes per instantiated graph (launch accounting)

int xtb_graph_end(void** graph_exec) {
    if (!graph_exec) XTB_FAIL(XTB_ERR_INVALID, "null graph");
    DeviceCtx* c;
    XTB_TRY(get_ctx(&c));
    cudaGraph_t graph = nullptr;
    XTB_CUDA(cudaStreamEndCapture(c->stream, &graph));
    int kernels = 0;
    {
        size_t n = 0;
        if (cudaGraphGetNodes(graph, nullptr, &n) == cudaSuccess && n > 0) {
            std::vector<cudaGraphNode_t> nodes(n);
            if (cudaGraphGetNodes(graph, nodes.data(), &n) == cudaSuccess) {
                for (size_t i = 0; i < n; ++i) {
                    cudaGraphNodeType t;
                    if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) ++kernels;
                }
            }
        }
    }
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) XTB_FAIL(XTB_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    {
        std::lock_guard<std::mutex> lock(g_graph_mutex);
        g_graph_kernels[(void*) exec] = kernels;
    }
    *graph_exec = (void*) exec;
    return XTB_OK;
}

int xtb_graph_kernel_count(void* graph_exec) {
    std::lock_guard<std::mutex> lock(g_graph_mutex);
    auto it = g_graph_kernels.find(graph_exec);
    return it == g_graph_kernels.end() ? -1 : it->second;
}

int xtb_graph_launch(void* graph_exec) {
    DeviceCtx* c;
    XTB_TRY(get_ctx(&c));
    XTB_CUDA(cudaGraphLaunch((cudaGraphExec_t) graph_exec, c->stream));
    int kernels = 1;
    {
        std::lock_guard<std::mutex> lock(g_graph_mutex);
        auto it = g_graph_kernels.find(graph_exec);
        if (it != g_graph_kernels.end()) kernels = it->second;
    }
    note_launch(nullptr, kernels);       // every kernel node of the graph is one launch of ours
    return XTB_OK;
}

int xtb_graph_destroy(void* graph_exec) {
    if (graph_exec) {
        {
            std::lock_guard<std::mutex> lock(g_graph_mutex);
            g_graph_kernels.erase(graph_exec);
        }
        XTB_CUDA(cudaGraphExecDestroy((cudaGraphExec_t) graph_exec));
    }
    return XTB_OK;
}

int xtb_set_option(const char* name, long long value) {
    if (!name) XTB_FAIL(XTB_ERR_INVALID, "null option name");
    Options& o = options();
    if (!strcmp(name, "no_static")) o.no_static = (int) value;
    else if (!strcmp(name, "no_jit")) o.no_jit = (int) value;
    else if (!strcmp(name, "no_staged")) o.no_staged = (int) value;
    else if (!strcmp(name, "no_tma")) o.no_tma = (int) value;
    else if (!strcmp(name, "jit_verbose")) o.jit_verbose = (int) value;
    else if (!strcmp(name, "jit_min_elems")) o.jit_min_elems = value;
    else if (!strcmp(name, "scan_variant")) o.scan_variant = (int) value;
    else if (!strcmp(name, "tile_variant")) o.tile_variant = (int) value;
    else if (!strcmp(name, "scan_nv")) o.scan_nv = (int) value;
    else if (!strcmp(name, "arg_two_pass")) o.arg_two_pass = (int) value;
    else if (!strcmp(name, "no_pdl")) o.no_pdl = (int) value;
    else if (!strcmp(name, "no_decompose")) o.no_decompose = (int) value;
    else if (!strcmp(name, "reduce_g")) o.reduce_g = (int) value;
    else if (!strcmp(name, "reduce_split")) o.reduce_split = (int) value;
    else XTB_FAIL(XTB_ERR_INVALID, "unknown option '%s'", name);
    return XTB_OK;
}

long long xtb_get_option(const char* name) {
    if (!name) return -1;
    const Options& o = options();
    if (!strcmp(name, "no_static")) return o.no_static;
    if (!strcmp(name, "no_jit")) return o.no_jit;
    if (!strcmp(name, "no_staged")) return o.no_staged;
    if (!strcmp(name, "no_tma")) return o.no_tma;
    if (!strcmp(name, "jit_verbose")) return o.jit_verbose;
    if (!strcmp(name, "jit_min_elems")) return o.jit_min_elems;
    if (!strcmp(name, "scan_variant")) return o.scan_variant;
    if (!strcmp(name, "tile_variant")) return o.tile_variant;
    if (!strcmp(name, "scan_nv")) return o.scan_nv;
    if (!strcmp(name, "arg_two_pass")) return o.arg_two_pass;
    if (!strcmp(name, "no_pdl")) return o.no_pdl;
    if (!strcmp(name, "no_decompose")) return o.no_decompose;
    if (!strcmp(name, "reduce_g")) return o.reduce_g;
    if (!strcmp(name, "reduce_split")) return o.reduce_split;
    return -1;
}

int64_t xtb_launch_count(int reset) {
    int64_t v = g_launches.load();
    if (reset) g_launches.store(0);
    return v;
}

const char* xtb_last_kernel(void) { return g_last_kernel; }

int xtb_program_result_type(const xtb_program* program, const int32_t* leaf_dtypes) {
    int rt = -1;
    bool w64 = false;
    int r = validate_program(program, leaf_dtypes, &rt, &w64);
    if (r != XTB_OK) return r;
    return rt;
}

}  // extern "C"
