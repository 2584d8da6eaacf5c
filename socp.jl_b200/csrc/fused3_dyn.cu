// fused3_dyn.cu -- the generic-layout instantiations of the whole-solve kernel k_fused3 (fused_v3.cuh), one team per
// CTA (several CTAs per SM) or four teams per CTA.
#include "fused_v3.cuh"

namespace socp {
void fused3_launch_dyn(const F3Plan& plan, const F3Args& args, int teams, int grid, cudaStream_t stream) {
    if (teams == 4) {
        if (plan.nb <= 4) fused3_launch<4, 4, 3, 1, Dims3Dyn>(plan, args, grid, stream);          // n <= 32: 10 tiles
        else if (plan.nb <= 7) fused3_launch<4, 4, 7, 1, Dims3Dyn>(plan, args, grid, stream);     // n <= 56: 28 tiles
        else fused3_launch<4, 4, 9, 1, Dims3Dyn>(plan, args, grid, stream);                       // n <= 64: 36 tiles
        return;
    }
    if (plan.nb <= 4) fused3_launch<4, 1, 3, 4, Dims3Dyn>(plan, args, grid, stream);
    else if (plan.nb <= 7) fused3_launch<4, 1, 7, 4, Dims3Dyn>(plan, args, grid, stream);
    else fused3_launch<4, 1, 9, 3, Dims3Dyn>(plan, args, grid, stream);
}
}  // namespace socp
