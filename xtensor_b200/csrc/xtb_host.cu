// xtb_host.cu -- xtb_assign_host: evaluate an expression whose operands live in HOST memory.
//
// The reference evaluates host containers in place; a caller that switches to the backend with
// host-resident data pays PCIe in both directions.  This entry point hides as much of that as
// the link allows: the output's leading dimension is cut into chunks, and three streams
// pipeline  H2D(chunk i+1) | kernel(chunk i) | D2H(chunk i-1)  over a ring of device staging
// slots, so the end-to-end time approaches max(bytes in, bytes out) / PCIe bandwidth instead of
// their sum.  Operands that do not span the leading dimension (broadcast / lower rank) are
// uploaded once.  Host buffers should be pinned (xtb_host_alloc) for the copies to be async.
#include <algorithm>
#include <mutex>
#include <vector>
#include "xtb_common.hpp"
#include "xtb_ops.cuh"

namespace xtb {

struct HostPipe {
    cudaStream_t in = nullptr, out = nullptr;
    static constexpr int kSlots = 3;
    cudaEvent_t in_done[kSlots] = {}, k_done[kSlots] = {}, out_done[kSlots] = {};
    bool ready = false;
};
// one pipeline (two copy streams + events) per device; a call holds the device's mutex from its first chunk to its
// last, so two host threads assigning on the same device take turns instead of sharing streams and events
static HostPipe g_pipe[16];
static std::mutex g_pipe_mutex[16];

static int pipe_for(DeviceCtx* ctx, HostPipe** pp) {
    HostPipe& p = g_pipe[ctx->device];
    if (!p.ready) {
        XTB_CUDA(cudaStreamCreateWithFlags(&p.in, cudaStreamNonBlocking));
        XTB_CUDA(cudaStreamCreateWithFlags(&p.out, cudaStreamNonBlocking));
        for (int i = 0; i < HostPipe::kSlots; ++i) {
            XTB_CUDA(cudaEventCreateWithFlags(&p.in_done[i], cudaEventDisableTiming));
            XTB_CUDA(cudaEventCreateWithFlags(&p.k_done[i], cudaEventDisableTiming));
            XTB_CUDA(cudaEventCreateWithFlags(&p.out_done[i], cudaEventDisableTiming));
        }
        p.ready = true;
    }
    *pp = &p;
    return XTB_OK;
}

static bool dense_row_major(const xtb_operand* op) {
    int64_t expect = 1;
    for (int d = op->ndim - 1; d >= 0; --d) {
        if (op->shape[d] != 1 && op->stride[d] != expect) return false;
        expect *= op->shape[d];
    }
    return true;
}

}  // namespace xtb

using namespace xtb;

extern "C" int xtb_assign_host(const xtb_program* prog, const xtb_operand* out, const xtb_operand* leaves, int64_t chunk_bytes) {
    if (!prog || !out) XTB_FAIL(XTB_ERR_INVALID, "null argument");
    if (prog->n_leaves < 0 || prog->n_leaves > XTB_MAX_LEAVES) XTB_FAIL(XTB_ERR_INVALID, "bad leaf count");
    if (!dense_row_major(out)) XTB_FAIL(XTB_ERR_UNSUPPORTED, "xtb_assign_host: output must be dense row-major");
    for (int k = 0; k < prog->n_leaves; ++k)
        if (!dense_row_major(&leaves[k])) XTB_FAIL(XTB_ERR_UNSUPPORTED, "xtb_assign_host: leaf %d must be dense row-major", k);
    DeviceCtx* ctx;
    XTB_TRY(get_ctx(&ctx));
    if (ctx->device < 0 || ctx->device >= 16) XTB_FAIL(XTB_ERR_UNSUPPORTED, "xtb_assign_host: device index %d", ctx->device);
    std::lock_guard<std::mutex> pipe_lock(g_pipe_mutex[ctx->device]);
    HostPipe* pipe = nullptr;
    XTB_TRY(pipe_for(ctx, &pipe));
    const int nd = out->ndim;
    int64_t total = 1;
    for (int d = 0; d < nd; ++d) total *= out->shape[d];
    if (total == 0) return XTB_OK;
    const int64_t rows = nd > 0 ? out->shape[0] : 1;
    const int64_t out_row_elems = total / std::max<int64_t>(rows, 1);
    const int osz = dtype_size(out->dtype);
    // which leaves are sliced along the leading dimension?
    bool sliced[XTB_MAX_LEAVES];
    int64_t leaf_row_elems[XTB_MAX_LEAVES], leaf_total[XTB_MAX_LEAVES];
    int64_t bytes_per_row = out_row_elems * osz;
    for (int k = 0; k < prog->n_leaves; ++k) {
        const xtb_operand& L = leaves[k];
        int64_t t = 1;
        for (int d = 0; d < L.ndim; ++d) t *= L.shape[d];
        leaf_total[k] = t;
        sliced[k] = nd > 0 && L.ndim == nd && L.shape[0] == rows && rows > 1;
        leaf_row_elems[k] = sliced[k] ? t / rows : 0;
        if (sliced[k]) bytes_per_row += leaf_row_elems[k] * dtype_size(L.dtype);
    }
    if (chunk_bytes <= 0) chunk_bytes = (int64_t) 48 << 20;
    int64_t chunk_rows = std::max<int64_t>(1, chunk_bytes / std::max<int64_t>(bytes_per_row, 1));
    chunk_rows = std::min(chunk_rows, rows);
    const int64_t n_chunks = (rows + chunk_rows - 1) / chunk_rows;
    constexpr int NS = HostPipe::kSlots;

    // device staging: per slot, one region per sliced leaf + the output slice; unsliced leaves once
    size_t slot_bytes = 0, fixed_bytes = 0;
    size_t leaf_off[XTB_MAX_LEAVES], out_off = 0;
    auto align = [](size_t x) { return (x + 255) / 256 * 256; };
    for (int k = 0; k < prog->n_leaves; ++k) {
        const int sz = dtype_size(leaves[k].dtype);
        if (sliced[k]) {
            leaf_off[k] = slot_bytes;
            slot_bytes += align((size_t) chunk_rows * leaf_row_elems[k] * sz);
        } else {
            leaf_off[k] = fixed_bytes;
            fixed_bytes += align((size_t) leaf_total[k] * sz);
        }
    }
    out_off = slot_bytes;
    slot_bytes += align((size_t) chunk_rows * out_row_elems * osz);
    char* dev = nullptr;
    XTB_CUDA(cudaMallocAsync((void**) &dev, fixed_bytes + NS * slot_bytes + 256, ctx->stream));
    XTB_CUDA(cudaStreamSynchronize(ctx->stream));
    char* fixed = dev;
    char* slots = dev + align(fixed_bytes);

    cudaStream_t user_stream = ctx->stream;
    int status = XTB_OK;
    auto fail = [&](int code) { status = code; };
    // unsliced leaves
    for (int k = 0; k < prog->n_leaves && status == XTB_OK; ++k) {
        if (sliced[k]) continue;
        const int sz = dtype_size(leaves[k].dtype);
        const char* src = (const char*) leaves[k].base + leaves[k].offset * sz;
        if (cudaMemcpyAsync(fixed + leaf_off[k], src, (size_t) leaf_total[k] * sz, cudaMemcpyHostToDevice, pipe->in) != cudaSuccess)
            fail(set_error(XTB_ERR_CUDA, "H2D of leaf %d failed", k));
    }
    for (int64_t c = 0; c < n_chunks && status == XTB_OK; ++c) {
        const int s = (int) (c % NS);
        const int64_t r0 = c * chunk_rows;
        const int64_t nr = std::min(chunk_rows, rows - r0);
        char* slot = slots + (size_t) s * slot_bytes;
        // the slot's previous occupant must have been consumed (kernel read inputs, D2H drained output)
        if (c >= NS) {
            cudaStreamWaitEvent(pipe->in, pipe->k_done[s], 0);
        }
        for (int k = 0; k < prog->n_leaves; ++k) {
            if (!sliced[k]) continue;
            const int sz = dtype_size(leaves[k].dtype);
            const char* src = (const char*) leaves[k].base + (leaves[k].offset + r0 * leaf_row_elems[k]) * sz;
            if (cudaMemcpyAsync(slot + leaf_off[k], src, (size_t) nr * leaf_row_elems[k] * sz, cudaMemcpyHostToDevice, pipe->in) != cudaSuccess)
                fail(set_error(XTB_ERR_CUDA, "H2D of leaf %d failed", k));
        }
        cudaEventRecord(pipe->in_done[s], pipe->in);
        // compute on the library stream
        cudaStreamWaitEvent(user_stream, pipe->in_done[s], 0);
        if (c >= NS) cudaStreamWaitEvent(user_stream, pipe->out_done[s], 0);
        xtb_operand dl[XTB_MAX_LEAVES], dout = *out;
        for (int k = 0; k < prog->n_leaves; ++k) {
            dl[k] = leaves[k];
            dl[k].offset = 0;
            if (sliced[k]) {
                dl[k].base = slot + leaf_off[k];
                dl[k].shape[0] = nr;
            } else {
                dl[k].base = fixed + leaf_off[k];
            }
        }
        dout.base = slot + out_off;
        dout.offset = 0;
        if (nd > 0) dout.shape[0] = nr;
        const int r = xtb_assign(prog, &dout, dl);
        if (r != XTB_OK) { fail(r); break; }
        cudaEventRecord(pipe->k_done[s], user_stream);
        // drain the output slice
        cudaStreamWaitEvent(pipe->out, pipe->k_done[s], 0);
        char* dst = (char*) out->base + (out->offset + r0 * out_row_elems) * osz;
        if (cudaMemcpyAsync(dst, slot + out_off, (size_t) nr * out_row_elems * osz, cudaMemcpyDeviceToHost, pipe->out) != cudaSuccess)
            fail(set_error(XTB_ERR_CUDA, "D2H of the result failed"));
        cudaEventRecord(pipe->out_done[s], pipe->out);
    }
    // the call returns when the host result is complete (the reference's assignment is synchronous)
    cudaStreamSynchronize(pipe->in);
    cudaStreamSynchronize(user_stream);
    cudaStreamSynchronize(pipe->out);
    cudaFreeAsync(dev, user_stream);
    if (status == XTB_OK) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return set_error(XTB_ERR_CUDA, "xtb_assign_host: %s", cudaGetErrorString(e));
    }
    return status;
}
