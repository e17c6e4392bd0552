// xtb_scan.cu -- xtb_scan: inclusive scan along an axis (cumsum / cumprod).
// Replaces detail::accumulator_impl (include/xtensor/reducers/xaccumulator.hpp:215-341).
#include "xtb_common.hpp"
#include "xtb_ops.cuh"

using namespace xtb;

extern "C" int xtb_scan(int op, int acc_type, const xtb_operand* in, int axis, const xtb_operand* out) {
    (void) op; (void) acc_type; (void) in; (void) axis; (void) out;
    XTB_FAIL(XTB_ERR_UNSUPPORTED, "xtb_scan is not implemented yet");
}
