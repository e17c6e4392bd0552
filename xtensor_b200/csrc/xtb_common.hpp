// xtb_common.hpp -- host-side plumbing shared by the translation units of
// libxtb200: per-thread error state, per-device context (stream, SM count,
// scratch), launch accounting and the operand canonicaliser.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <utility>
#include <cuda_runtime.h>
#include "../../include/xtb200.h"

namespace xtb {

// ---- error state -------------------------------------------------------------
int set_error(int code, const char* fmt, ...);
#define XTB_FAIL(code, ...) return ::xtb::set_error((code), __VA_ARGS__)
#define XTB_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return ::xtb::set_error(e__ == cudaErrorMemoryAllocation ? XTB_ERR_OOM : XTB_ERR_CUDA, \
                                    "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),   \
                                    __FILE__, __LINE__);                                       \
    } while (0)
#define XTB_TRY(call)             \
    do {                          \
        int r__ = (call);         \
        if (r__ != XTB_OK) return r__; \
    } while (0)

// ---- device context ----------------------------------------------------------
struct DeviceCtx {
    int device = -1;
    bool ready = false;
    cudaStream_t own_stream = nullptr;   // created by the library
    cudaStream_t stream = nullptr;       // stream in use (own or adopted)
    int sm_count = 0;
    size_t l2_bytes = 0;
    void* scratch = nullptr;             // reduction partials / scan state
    size_t scratch_bytes = 0;
    // xtb_fork_begin / end / join: a second stream for work that may overlap the main sequence
    cudaStream_t fork_stream = nullptr;
    cudaStream_t fork_saved = nullptr;   // the stream to return to at xtb_fork_end
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    bool forked = false;
    void* fork_scratch = nullptr;        // calls on the fork stream get their own scratch
    size_t fork_scratch_bytes = 0;
    // the stream whose work last touched `scratch`: a call on another stream first waits for it
    cudaStream_t scratch_stream = nullptr;
    cudaEvent_t scratch_ev = nullptr;
    // One compute call (xtb_assign / xtb_reduce / xtb_scan / xtb_argreduce) is a SEQUENCE of launches that share this
    // context's scratch buffer; host threads working on the same device take turns per call (the launches themselves
    // are asynchronous, so the lock is held for microseconds).  Recursive: xtb_argreduce calls xtb_reduce.
    // Stream selection (xtb_set_stream), fork sections and graph capture remain per-device state: drive them from one
    // thread at a time.
    std::recursive_mutex launch_mutex;
};
#define XTB_LAUNCH_LOCK(ctx) std::lock_guard<std::recursive_mutex> launch_lock__((ctx)->launch_mutex)
// ---- process options -----------------------------------------------------------
// Read from the environment ONCE (first use) and changeable through xtb_set_option; the dispatchers never
// call getenv.  Names are the environment variables without the XTB_ prefix, lower case.
struct Options {
    int no_static = 0;        // XTB_NO_STATIC: skip the ahead-of-time instantiations (tests: evaluators agree)
    int no_jit = 0;           // XTB_NO_JIT: skip run-time specialisation
    int no_staged = 0;        // XTB_NO_STAGED: interpreter without cp.async staging
    int no_tma = 0;           // XTB_NO_TMA: tile kernels fetch with plain / bulk copies
    int jit_verbose = 0;      // XTB_JIT_VERBOSE
    long long jit_min_elems = 1 << 20;   // XTB_JIT_MIN_ELEMS: problems below this size use the interpreter
    int scan_variant = 0;     // XTB_SCAN_VARIANT: development switch of xtb_scan
    int scan_nv = 0;          // XTB_SCAN_NV: 128-bit vectors per thread of k_scan_ahead (4 or 8)
    int tile_variant = 0;     // XTB_TILE_VARIANT: development switch of the transposed-leaf kernel
    int arg_two_pass = 0;     // XTB_ARG_TWO_PASS: argmin / argmax of 32-bit types through the two-pass formulation (tests)
    int reduce_split = 0;     // XTB_REDUCE_SPLIT: development override of the row-split count of k_reduce_outer
    int reduce_g = 0;         // XTB_REDUCE_G: development override of the lanes per output of the contiguous-axis kernels (32 / 256)
    int no_decompose = 0;     // XTB_NO_DECOMPOSE: single-pass reductions only (tests: both formulations agree)
    int no_pdl = 0;           // XTB_NO_PDL: launch every kernel fully serialised (no programmatic dependent launch)
};
Options& options();

// Context of the calling thread's device; fails with XTB_ERR_NO_DEVICE when
// there is no GPU.  (No CPU fallback by design.)
int get_ctx(DeviceCtx** ctx);
int ensure_scratch(DeviceCtx* ctx, size_t bytes, void** ptr);
void note_launch(const char* kernel_name, int n = 1);
int check_launch(const char* what);

// ---- programmatic dependent launch ----------------------------------------------
// The bandwidth-bound kernels of a pipeline (reduce -> merge -> reduce -> merge, map) are tens to hundreds of
// microseconds long, so the 2-4 us between two dependent launches is worth removing: a kernel that starts with
// pdl_enter() (xtb_ops.cuh: griddepcontrol.launch_dependents + griddepcontrol.wait) may be launched with the
// programmatic-stream-serialization attribute; its CTAs then become resident while the previous kernel's last CTAs
// drain and wait, on the device, for that kernel's completion and memory flush.  Works eagerly and under stream
// capture (the edge becomes a programmatic graph edge).  ONLY for kernels whose every thread executes pdl_enter()
// before touching global memory.
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = options().no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- iteration-space canonicaliser --------------------------------------------
// Operands (leaves + out) are re-expressed over one common iteration space:
// extent-1 dims dropped, dims optionally permuted, adjacent dims merged when
// every operand walks them with one stride.  This replaces the per-element
// stepper bookkeeping of xiterator.hpp:484-631 by a handful of integers.
struct Space {
    int ndim = 0;
    int64_t shape[XTB_MAX_DIM] = {0};
    int n_ops = 0;                                        // operands tracked
    int64_t stride[XTB_MAX_LEAVES + 2][XTB_MAX_DIM] = {{0}};  // elements
    bool reduced[XTB_MAX_DIM] = {false};                  // reduce planning only
    int64_t total = 1;                                    // number of points
};

// Broadcast-align `op` (rank <= ndim, right aligned) against shape[ndim] and
// write its per-dimension strides (0 where broadcast). Returns XTB_ERR_SHAPE if
// not broadcastable.
int align_operand(const xtb_operand* op, int ndim, const int64_t* shape, int64_t* stride_out,
                  const char* what);
// Drop extent-1 dims; merge adjacent dims (never across a reduced/kept boundary).
void collapse_space(Space* s);
// Sort dims so that operand `key` has non-increasing |stride| (elementwise only).
void sort_space_by(Space* s, int key);

inline char* operand_ptr(const xtb_operand* op, int elem_size) {
    return (char*) op->base + op->offset * (int64_t) elem_size;
}

int validate_program(const xtb_program* p, const int32_t* leaf_dtypes, int* result_type,
                     bool* needs64);

}  // namespace xtb
