// bmu_tc_inst_plain.cu -- instantiates the ACC=false family of bmu_tc_kernel.
#define PIXIE_FAMILY_ACC false
#include "bmu_tc_kernel.cuh"

namespace pixie {

cudaError_t launch_tc_family_plain(const CUtensorMap &tmX, const TcParams &p, int grid,
                                    cudaStream_t stream)
{
    const TcPlan &pl = p.plan;
    PIXIE_ALL_VARIANTS
    return cudaErrorInvalidValue;
}

}  // namespace pixie
