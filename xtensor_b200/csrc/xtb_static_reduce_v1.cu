// xtb_static_reduce_v1.cu -- ahead-of-time reduction kernels, scalar access (see xtb_static_reduce_impl.cuh)
#define XTB_SR_SCALAR 1
#define XTB_SR_TABLE static_reduce_table_v1
#include "xtb_static_reduce_impl.cuh"
