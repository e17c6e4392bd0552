// xtb_jit.hpp -- run-time specialisation of the kernel templates (see xtb_jit.cu).
#pragma once
#include "xtb_common.hpp"

namespace xtb {

enum JitKind { JIT_EW = 0, JIT_TILE = 1, JIT_RED_OUTER = 2, JIT_RED_INNER_WARP = 3, JIT_RED_INNER_BLOCK = 4, JIT_RED_ROWS_EXACT = 5 };

struct JitSpec {
    int kind = JIT_EW;
    int w64 = 0;      // slot width: 0 -> uint32_t, 1 -> uint64_t
    int V = 4;        // vector width (elements)
    int nd = 1;       // collapsed rank (JIT_EW)
    int binop = 0;    // reducer functor opcode (reduce kinds)
    int acc_rt = 0;   // accumulator register type (reduce kinds)
};

bool jit_program_ok(const xtb_program* p);
bool jit_worthwhile(int64_t elements);
// Returns XTB_OK and the kernel handle, or XTB_ERR_UNSUPPORTED when NVRTC / libcuda are not
// available or the compile failed (callers then use the interpreter kernels).
int jit_get(const DeviceCtx* ctx, const xtb_program* prog, const JitSpec& spec, void** fn);
int jit_launch(void* fn, unsigned grid_x, unsigned grid_y, unsigned block, size_t smem, cudaStream_t stream, const void* params);
int jit_compile_count();

}  // namespace xtb
