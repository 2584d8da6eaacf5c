// fused3.cu -- translation unit of the second-generation whole-solve kernel (fused_v3.cuh): the entry point and the
// kernels specialised at compile time for the BASELINE C2 layout.  The generic-layout kernels are in fused3_dyn.cu.
#include "fused_v3.cuh"

namespace socp {
void fused3_launch_c2(const F3Plan& plan, const F3Args& args, int teams, int grid, cudaStream_t stream) {
    if (teams == 4) fused3_launch<4, 4, 7, 1, Dims3C2>(plan, args, grid, stream);
    else fused3_launch<4, 1, 7, 4, Dims3C2>(plan, args, grid, stream);
}
void solve_fused3_ext(const F3Plan& plan, const F3Glob& g, int first, int batch, const LoopParams& prm, int sing_detect,
                      int verify, cudaStream_t stream, bool allow_static, int counter_slot) {
    solve_fused3(plan, g, first, batch, prm, sing_detect, verify, stream, allow_static, counter_slot);
}
}  // namespace socp
