/*
 * xtb200.h -- C ABI of libxtb200.so, the B200 (sm_100a) evaluation backend for
 * xtensor's assignment / reduction / accumulation hot path.
 *
 * The reference (xtensor 0.27.1) is header-only C++ and has no FFI of its own;
 * its evaluator is a set of template loops.  Each entry point below replaces
 * one of those loops and is what the header-only boundary
 * (include/xtb200/xtensor_b200.hpp: a new expression tag + a specialisation of
 * xt::xexpression_assigner_base) binds to:
 *
 *   xtb_assign   <- xexpression_assigner_base<xtensor_expression_tag>::assign_data
 *                   include/xtensor/core/xassign.hpp:439-478 and the three loops
 *                   it dispatches to: linear_assigner::run :701-849,
 *                   strided_loop_assigner::run :1100-1344,
 *                   stepper_assigner::run :644-695 (+ increment_stepper,
 *                   include/xtensor/core/xiterator.hpp:589-631)
 *   xtb_reduce   <- reduce_immediate  include/xtensor/reducers/xreducer.hpp:289-565
 *                   and xreducer_stepper::aggregate_impl :1778-1868 (lazy reducers)
 *   xtb_scan     <- detail::accumulator_impl include/xtensor/reducers/xaccumulator.hpp:215-341
 *   xtb_argreduce <- argmin / argmax, detail::arg_func_impl include/xtensor/misc/xsort.hpp:1150-1300
 *   xtb_malloc/xtb_free/xtb_memcpy
 *                <- uvector<T,A> storage   include/xtensor/containers/xstorage.hpp:33-345
 *   xtb_allreduce / reduce(..., allreduce=1)
 *                <- merge step of xblockwise_reducer functors
 *                   include/xtensor/reducers/xblockwise_reducer_functors.hpp:45-260
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on
 * success or a negative xtb_status, and xtb_last_error() gives the message for
 * the calling thread.  All work is enqueued on the calling device's stream and
 * is stream ordered; xtb_sync() or a device->host xtb_memcpy is the blocking
 * point (the reference's loops are synchronous; the boundary header syncs where
 * host code can observe results).  Descriptors are caller owned and only read
 * during the call.  There is no CPU fallback: without a CUDA device every
 * compute entry point fails with XTB_ERR_NO_DEVICE.
 */
#ifndef XTB200_H
#define XTB200_H

#ifndef __CUDACC_RTC__ /* NVRTC (run-time kernel specialisation) brings its own fixed-width types */
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define XTB_ABI_VERSION 1

#define XTB_MAX_DIM    8   /* rank limit of any operand / iteration space        */
#define XTB_MAX_LEAVES 8   /* tensor leaves of one lowered expression           */
#define XTB_MAX_INSNS  48  /* instructions of one lowered expression            */
#define XTB_MAX_IMMS   16  /* scalar immediates (xscalar leaves)                */
#define XTB_MAX_STACK  8   /* evaluation stack depth                            */

/* ---- status codes --------------------------------------------------------- */
typedef enum {
    XTB_OK = 0,
    XTB_ERR_INVALID = -1,      /* malformed descriptor / program                 */
    XTB_ERR_SHAPE = -2,        /* shapes not broadcastable (xt::broadcast_error)  */
    XTB_ERR_AXIS = -3,         /* reducer axes unsorted / duplicate / out of range */
    XTB_ERR_UNSUPPORTED = -4,  /* valid request the backend does not implement    */
    XTB_ERR_CUDA = -5,         /* CUDA runtime error                              */
    XTB_ERR_NO_DEVICE = -6,    /* no CUDA device: there is no CPU fallback        */
    XTB_ERR_NCCL = -7,
    XTB_ERR_OOM = -8
} xtb_status;

/* ---- element types -------------------------------------------------------- */
/* Storage dtypes of containers.  bool is stored as one byte (0/1), like
 * xt_simd's bool_load_type (include/xtensor/utils/xtensor_simd.hpp:298-299).   */
typedef enum {
    XTB_BOOL = 0, XTB_I8 = 1, XTB_U8 = 2, XTB_I16 = 3, XTB_U16 = 4,
    XTB_I32 = 5, XTB_U32 = 6, XTB_I64 = 7, XTB_U64 = 8, XTB_F32 = 9, XTB_F64 = 10,
    XTB_DTYPE_COUNT = 11
} xtb_dtype;

/* Register ("compute") types are the dtypes >= XTB_I32: C++ integral promotion
 * turns bool/int8/uint8/int16/uint16 into int before any arithmetic, so a leaf
 * of a narrow dtype is widened to XTB_I32 when it is pushed.                    */

/* ---- operand descriptor --------------------------------------------------- */
/* An affine view of device memory: element (i0..in-1) lives at
 *     base + (offset + sum_d i_d * stride[d]) * sizeof(dtype)
 * stride[d] is in elements, signed; 0 for broadcast dimensions and for extent-1
 * dimensions (include/xtensor/core/xstrides.hpp:503-530).  Leaves may have a
 * lower rank than the iteration space: they are aligned from the right
 * (xt::broadcast_shape, xstrides.hpp:737-780).                                  */
typedef struct {
    void*   base;                 /* storage().data() (device pointer)           */
    int64_t offset;               /* data_offset(), elements                     */
    int32_t dtype;                /* xtb_dtype                                   */
    int32_t ndim;                 /* 0..XTB_MAX_DIM                              */
    int64_t shape[XTB_MAX_DIM];
    int64_t stride[XTB_MAX_DIM];
} xtb_operand;

/* ---- lowered expression: postfix program over a typed value stack --------- */
typedef enum {
    /* stack */
    XTB_OP_PUSH = 0,   /* src=LEAF: push leaf[arg] (type = leaf storage dtype, widened);
                          src=IMM : push imm[arg] as register type `type`          */
    XTB_OP_CAST = 1,   /* top = static_cast<dtype arg>(top); `type` = source register type;
                          a narrow arg (e.g. XTB_I8) wraps then widens back to int */
    /* unary, xt::detail::{identity,negate,logical_not,bitwise_not}
       include/xtensor/core/xoperation.hpp:104-117 */
    XTB_OP_NEG = 2, XTB_OP_NOT = 3, XTB_OP_BITNOT = 4,
    /* unary math functors, include/xtensor/core/xmath.hpp:295-341 */
    XTB_OP_ABS = 5, XTB_OP_EXP = 6, XTB_OP_EXP2 = 7, XTB_OP_EXPM1 = 8, XTB_OP_LOG = 9,
    XTB_OP_LOG10 = 10, XTB_OP_LOG2 = 11, XTB_OP_LOG1P = 12, XTB_OP_SQRT = 13, XTB_OP_CBRT = 14,
    XTB_OP_SIN = 15, XTB_OP_COS = 16, XTB_OP_TAN = 17, XTB_OP_ASIN = 18, XTB_OP_ACOS = 19,
    XTB_OP_ATAN = 20, XTB_OP_SINH = 21, XTB_OP_COSH = 22, XTB_OP_TANH = 23, XTB_OP_ASINH = 24,
    XTB_OP_ACOSH = 25, XTB_OP_ATANH = 26, XTB_OP_ERF = 27, XTB_OP_ERFC = 28, XTB_OP_TGAMMA = 29,
    XTB_OP_LGAMMA = 30, XTB_OP_CEIL = 31, XTB_OP_FLOOR = 32, XTB_OP_TRUNC = 33, XTB_OP_ROUND = 34,
    XTB_OP_NEARBYINT = 35, XTB_OP_RINT = 36,
    XTB_OP_ISFINITE = 37, XTB_OP_ISINF = 38, XTB_OP_ISNAN = 39,   /* -> bool (I32 0/1) */
    XTB_OP_SIGN = 40,      /* math::sign_fun xmath.hpp:826-866 */
    XTB_OP_DEG2RAD = 41, XTB_OP_RAD2DEG = 42,                      /* xmath.hpp:619-672 */
    XTB_OP_SQUARE = 43, XTB_OP_CUBE = 44,                          /* xmath.hpp:1100-1127 */
    /* order key of a 32-bit value (type = F32 / I32 / U32) in the HIGH half of a u64: a monotone map of x (arg = 0) or
       of -x (arg = 1), -0.0 == +0.0, NaN -> 0xffffffff (never the minimum).  OR-ed with an index in the low half, a
       MIN reduction over such keys is argmin / argmax with the first index winning ties (xtb_argreduce) */
    XTB_OP_ORDKEY = 45,
    /* binary, xoperation.hpp:106-125 */
    XTB_OP_ADD = 64, XTB_OP_SUB = 65, XTB_OP_MUL = 66, XTB_OP_DIV = 67, XTB_OP_MOD = 68,
    XTB_OP_LOR = 69, XTB_OP_LAND = 70, XTB_OP_BOR = 71, XTB_OP_BAND = 72, XTB_OP_BXOR = 73,
    XTB_OP_SHL = 74, XTB_OP_SHR = 75,
    XTB_OP_LT = 76, XTB_OP_LE = 77, XTB_OP_GT = 78, XTB_OP_GE = 79, XTB_OP_EQ = 80, XTB_OP_NE = 81,
    /* binary math, xmath.hpp:297-321 */
    XTB_OP_FMOD = 82, XTB_OP_REMAINDER = 83, XTB_OP_FMAX = 84, XTB_OP_FMIN = 85, XTB_OP_FDIM = 86,
    XTB_OP_POW = 87, XTB_OP_HYPOT = 88, XTB_OP_ATAN2 = 89,
    /* select based, math::maximum / math::minimum xmath.hpp:570-602: (a>b)?a:b / (a<b)?a:b */
    XTB_OP_MAXIMUM = 90, XTB_OP_MINIMUM = 91,
    /* detail::nan_min / nan_max xmath.hpp:2333-2363: isnan(a) ? b : (isnan(b) ? a : minimum/maximum(a, b)) */
    XTB_OP_NANMIN = 92, XTB_OP_NANMAX = 93,
    /* ternary: stack [x, y, z] -> one value */
    XTB_OP_WHERE = 112,    /* x ? y : z      detail::conditional_ternary xoperation.hpp:126-143 */
    XTB_OP_FMA = 113,      /* std::fma(x,y,z) math::fma_fun */
    XTB_OP_CLAMP = 114     /* x<y ? y : (z<x ? z : x)   math::clamp_fun xmath.hpp:604-617 */
} xtb_opcode;

typedef enum {
    XTB_SRC_STACK = 0,     /* second operand of a binary op is popped from the stack */
    XTB_SRC_LEAF = 1,      /* ... is leaf[arg] (must already have register type `type`) */
    XTB_SRC_IMM = 2,       /* ... is imm[arg]                                        */
    XTB_SRC_REV = 4        /* flag: operands swapped, result = op(src, top)          */
} xtb_src;

typedef struct {
    uint8_t op;            /* xtb_opcode                                            */
    uint8_t type;          /* register type the op computes in (operand type for
                              comparisons / predicates / casts); for PUSH LEAF the
                              leaf's storage dtype                                  */
    uint8_t src;           /* xtb_src (PUSH: LEAF or IMM)                           */
    uint8_t arg;           /* leaf / imm index, or the CAST target dtype            */
} xtb_insn;

typedef struct {
    int32_t  n_insns;
    int32_t  n_leaves;
    int32_t  n_imms;
    int32_t  reserved;
    xtb_insn insns[XTB_MAX_INSNS];
    uint64_t imms[XTB_MAX_IMMS];  /* raw bits in the register type of the using insn
                                     (f32/i32/u32 in the low 32 bits)               */
} xtb_program;

/* ---- reducers / accumulators ---------------------------------------------- */
typedef enum {
    XTB_RED_SUM = 0,       /* detail::plus, init 0            xmath.hpp:1803        */
    XTB_RED_PROD = 1,      /* detail::multiplies, init 1      xmath.hpp:1823        */
    XTB_RED_MAX = 2,       /* math::maximum, init lowest()    xmath.hpp:777-782     */
    XTB_RED_MIN = 3,       /* math::minimum, init max()       xmath.hpp:795-800     */
    XTB_RED_NANMIN = 4,    /* detail::nan_min, init NaN       xmath.hpp:2333-2346, 2427 (integers: = MIN) */
    XTB_RED_NANMAX = 5     /* detail::nan_max, init NaN       xmath.hpp:2348-2363, 2442 (integers: = MAX) */
} xtb_reduce_op;

/* ---- runtime --------------------------------------------------------------- */
int  xtb_abi_version(void);
/* Select the CUDA device this thread's subsequent calls run on and create its
 * stream / memory pool on first use.  device < 0 keeps the current device.     */
int  xtb_init(int device);
int  xtb_device_count(int* count);
int  xtb_sync(void);
const char* xtb_last_error(void);
/* Adopt an existing cudaStream_t for the current device (NULL = library stream). */
int  xtb_set_stream(void* cuda_stream);
void* xtb_get_stream(void);

/* storage: device_uvector<T> (uvector contract: contents uninitialised,
 * resize discards; xstorage.hpp:217-228)                                        */
int  xtb_malloc(size_t bytes, void** ptr);
int  xtb_free(void* ptr);
typedef enum { XTB_H2D = 1, XTB_D2H = 2, XTB_D2D = 3 } xtb_copy_kind;
/* H2D/D2D are stream ordered (asynchronous w.r.t. the host when src is pinned);
 * D2H synchronises the stream before returning.                                */
int  xtb_memcpy(void* dst, const void* src, size_t bytes, int kind);
int  xtb_memset(void* dst, int byte, size_t bytes);
/* pinned host staging buffers for the end-to-end path */
int  xtb_host_alloc(size_t bytes, void** ptr);
int  xtb_host_free(void* ptr);

/* device-side timing on the library stream (CUDA events), for harnesses */
int  xtb_event_create(void** event);
int  xtb_event_record(void* event);
int  xtb_event_elapsed_ms(void* start, void* stop, float* ms);  /* syncs on `stop` */
int  xtb_event_destroy(void* event);

/* CUDA-graph capture of a sequence of calls (kernels, xtb_allreduce) on the library stream:
 * launch-bound pipelines (many short kernels, e.g. the sharded mean / variance / map step on
 * 8 GPUs) replay without host launch gaps.  Between begin and end the calls are recorded, not
 * executed; allocate outputs (and run the sequence once, to size internal scratch) beforehand. */
int  xtb_graph_begin(void);
int  xtb_graph_end(void** graph_exec);
int  xtb_graph_launch(void* graph_exec);
int  xtb_graph_destroy(void* graph_exec);

/* Overlap: calls between xtb_fork_begin and xtb_fork_end are enqueued on a second stream that starts
 * after everything enqueued so far; calls after xtb_fork_end continue on the main stream concurrently
 * with them; xtb_fork_join makes the main stream wait for the forked section.  Works inside graph
 * capture (join before xtb_graph_end).  Used to hide the allreduce of one reduction behind an
 * independent elementwise kernel (sharded variance || exp(a - mean)).  The forked section must not
 * write what the concurrent main-stream calls read or write. */
int  xtb_fork_begin(void);
int  xtb_fork_end(void);
int  xtb_fork_join(void);

/* ---- the hot path ---------------------------------------------------------- */
/* out(i...) = static_cast<out.dtype>( program(leaves...)(i...) ) over out's shape.
 * Every leaf must be broadcastable to out's shape (else XTB_ERR_SHAPE). */
int  xtb_assign(const xtb_program* program, const xtb_operand* out,
                const xtb_operand* leaves);

/* Same as xtb_assign, but `out` and every leaf live in HOST memory (dense row-major; pinned
 * memory from xtb_host_alloc makes the copies asynchronous).  The leading dimension is cut
 * into chunks of about `chunk_bytes` (<= 0: default) and H2D / kernel / D2H of consecutive
 * chunks are pipelined on three streams; returns when the host result is complete. */
int  xtb_assign_host(const xtb_program* program, const xtb_operand* out,
                     const xtb_operand* leaves, int64_t chunk_bytes);

/* out = reduce_{op}( program(leaves...), axes ).
 *   ndim / shape : the iteration space = broadcast shape of the expression
 *   axes[n_axes] : sorted, unique, in range (else XTB_ERR_AXIS), n_axes may be 0
 *   acc_type     : register type of the accumulator = decltype(reduce(init, x))
 *   keep_dims    : out has rank ndim with extent 1 on reduced axes, else rank ndim-n_axes
 *   initial      : nullable pointer to one acc_type value merged once at the end
 *                  (xt::initial, xreducer.hpp:552-563)
 *   allreduce    : 1 = combine the per-rank results over the communicator set with
 *                  xtb_comm_init before returning (leading-axis sharded inputs) */
int  xtb_reduce(int op, int acc_type, const xtb_program* program,
                const xtb_operand* leaves, int ndim, const int64_t* shape,
                int n_axes, const int32_t* axes, int keep_dims,
                const void* initial, const xtb_operand* out, int allreduce);

/* xtb_reduce with a finalize step fused into the last store, applied once after every merge (including the
 * cross-GPU one): the `finalize` of the reference's blockwise reducer functors
 * (include/xtensor/reducers/xblockwise_reducer_functors.hpp:146-186, 246-258) and the division of
 * xt::mean / xt::variance (include/xtensor/core/xmath.hpp:1827-1852, 2082-2105).
 *   out = static_cast<out.dtype>( T(acc) / imm )            XTB_FIN_DIV
 *   out = static_cast<out.dtype>( sqrt(T(acc) / imm) )      XTB_FIN_DIV_SQRT   (xt::stddev)
 * with T = `type` (XTB_F32 or XTB_F64) and imm the divisor's bits in T.  fin == NULL: plain xtb_reduce. */
typedef enum { XTB_FIN_NONE = 0, XTB_FIN_DIV = 1, XTB_FIN_DIV_SQRT = 2 } xtb_finalize_op;
typedef struct {
    int32_t  op;     /* xtb_finalize_op */
    int32_t  type;   /* XTB_F32 / XTB_F64 */
    uint64_t imm;    /* divisor, raw bits in `type` (f32 in the low 32 bits) */
} xtb_finalize;
int  xtb_reduce_fin(int op, int acc_type, const xtb_program* program,
                    const xtb_operand* leaves, int ndim, const int64_t* shape,
                    int n_axes, const int32_t* axes, int keep_dims,
                    const void* initial, const xtb_operand* out, int allreduce,
                    const xtb_finalize* fin);

/* inclusive scan (cumsum: op = XTB_RED_SUM, cumprod: XTB_RED_PROD) of `in` along
 * `axis`, or over the flattened row-major traversal when axis < 0.  out has the
 * shape of in (or is 1-D of in's size when axis < 0) and dtype = acc_type.      */
int  xtb_scan(int op, int acc_type, const xtb_operand* in, int axis,
              const xtb_operand* out);

/* Index of the first minimum (op = XTB_RED_MIN) / maximum (XTB_RED_MAX) of `in` along `axis`, or over the
 * flattened row-major traversal when axis < 0 (then `in` must be dense row-major): xt::argmin / xt::argmax and
 * detail::arg_func_impl, include/xtensor/misc/xsort.hpp:1150-1300.  Sequential semantics of the reference:
 * ties keep the first index, a NaN never wins unless it is element 0 of its lane.  out: std::size_t (XTB_U64)
 * indices, rank in.ndim - 1 (0-d when flat). */
int  xtb_argreduce(int op, const xtb_operand* in, int axis, const xtb_operand* out);

/* ---- multi-GPU: one process per GPU ---------------------------------------- */
/* 128-byte NCCL unique id, created on rank 0 and handed to all ranks by the host
 * (torch.distributed / MPI / a file).                                           */
int  xtb_comm_unique_id(void* id128);
int  xtb_comm_init(int rank, int world, const void* id128);
int  xtb_comm_destroy(void);
int  xtb_comm_info(int* rank, int* world);
/* Optional: peer-memory allreduce for small payloads (<= 256 KB, 32-/64-bit dtypes).  Every rank
 * calls xtb_comm_p2p_handle (64-byte CUDA IPC handle of its exchange window), the host hands all
 * handles, in rank order, to every rank, which calls xtb_comm_p2p_attach.  From then on
 * xtb_allreduce / xtb_reduce(..., allreduce=1) use one kernel over NVLink peer memory (partials added
 * in rank order: identical bits on every rank) instead of NCCL for such payloads.  Every rank must
 * finish xtb_comm_p2p_attach before any rank's next allreduce.  xtb_comm_p2p_attach(NULL, 0) detaches (every
 * payload goes through NCCL again); a rank whose attach failed must make all ranks detach. */
int  xtb_comm_p2p_handle(void* handle64);
int  xtb_comm_p2p_attach(const void* handles, int world);
/* in-place allreduce of `count` elements of storage dtype `dtype` on the stream */
int  xtb_allreduce(void* buf, size_t count, int dtype, int op);

/* ---- process options ------------------------------------------------------- */
/* The XTB_* environment switches are read once, at first use; afterwards they are changed through this call
 * (name = the variable without the XTB_ prefix, lower case: "no_static", "no_jit", "no_staged", "no_tma",
 * "jit_min_elems", "jit_verbose", "scan_variant", "scan_nv", "tile_variant", "arg_two_pass";
 *   "no_pdl"       1: every kernel launch fully serialised (default: the reduce / merge / map kernels of a pipeline
 *                  use programmatic dependent launch);
 *   "no_decompose" 1: single-pass reductions only (default: narrow and mixed-axis reductions run as two passes);
 *   "reduce_split", "reduce_g": development overrides of the reduction planner (row-split count / lanes per output)).
 * Changing an option does not change results beyond the documented summation-order freedom of split reductions.
 * xtb_get_option returns -1 for unknown names. */
int  xtb_set_option(const char* name, long long value);
long long xtb_get_option(const char* name);

/* ---- introspection (tests, bench.py) --------------------------------------- */
/* kernel nodes of an instantiated graph (each replay counts that many launches) */
int  xtb_graph_kernel_count(void* graph_exec);
/* number of kernels launched by this library since the last reset            */
int64_t xtb_launch_count(int reset);
/* name of the kernel variant the last xtb_assign/xtb_reduce/xtb_scan selected  */
const char* xtb_last_kernel(void);
/* validate a program on the host: returns the register type of its result (>=0)
 * or a negative status                                                         */
int  xtb_program_result_type(const xtb_program* program, const int32_t* leaf_dtypes);

#ifdef __cplusplus
}
#endif
#endif /* XTB200_H */
