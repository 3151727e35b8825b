// bmu_tc_inst_acc.cu -- instantiates the ACC=true family of bmu_tc_kernel.
#define PIXIE_FAMILY_ACC true
#define PIXIE_FAMILY_NO_T8 1
#include "bmu_tc_kernel.cuh"

namespace pixie {

cudaError_t launch_tc_family_acc(const CUtensorMap &tmX, const TcParams &p, int grid,
                                    cudaStream_t stream)
{
    const TcPlan &pl = p.plan;
    PIXIE_ALL_VARIANTS
    return cudaErrorInvalidValue;
}

}  // namespace pixie
