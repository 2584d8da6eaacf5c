// fused3.cu -- translation unit of the second-generation whole-solve kernel (fused_v3.cuh).
#include "fused_v3.cuh"

namespace socp {
void solve_fused3_ext(const F3Plan& plan, const F3Glob& g, int first, int batch, const LoopParams& prm, int sing_detect,
                      int verify, cudaStream_t stream, bool allow_static, int counter_slot) {
    solve_fused3(plan, g, first, batch, prm, sing_detect, verify, stream, allow_static, counter_slot);
}
}  // namespace socp
