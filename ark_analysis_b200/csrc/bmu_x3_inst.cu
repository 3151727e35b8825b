// bmu_x3_inst.cu -- instantiates the split-operand assignment kernel (bmu_x3_kernel.cuh).
#include <float.h>

#define PIXIE_FAMILY_ACC false
#include "bmu_x3_kernel.cuh"

namespace pixie {

cudaError_t launch_bmu_x3(const CUtensorMap &tmX, const TcParams &p, int num_sms,
                          cudaStream_t stream)
{
    if (p.ntiles <= 0) return cudaSuccess;
    int grid = num_sms;
    if ((int64_t)grid > p.ntiles) grid = (int)p.ntiles;
    const TcPlan &pl = p.plan;
    if (!pl.x3) return cudaErrorInvalidValue;
#define PIXIE_X3_VARIANT(a_, b_) \
    if (pl.SL == a_ && pl.spc == b_) return launch_x3_variant<a_, b_>(tmX, p, grid, stream);
    PIXIE_X3_VARIANT(32, 1)
    PIXIE_X3_VARIANT(32, 2)
    PIXIE_X3_VARIANT(48, 2)
    PIXIE_X3_VARIANT(50, 2)
    PIXIE_X3_VARIANT(52, 2)
#undef PIXIE_X3_VARIANT
    return cudaErrorInvalidValue;
}

}  // namespace pixie
