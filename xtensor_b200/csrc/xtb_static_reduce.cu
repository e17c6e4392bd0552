// xtb_static_reduce.cu -- ahead-of-time reduction kernels, 128-bit vector access (see xtb_static_reduce_impl.cuh)
#define XTB_SR_SCALAR 0
#define XTB_SR_TABLE static_reduce_table
#include "xtb_static_reduce_impl.cuh"
