// fused3.cu -- translation unit of the second-generation whole-solve kernel (fused_v3.cuh).
#include "fused_v3.cuh"

namespace socp {
void solve_fused3_ext(const F3Plan& plan, const F3Glob& g, int first, int batch, const LoopParams& prm, int sing_detect,
                      int verify, cudaStream_t stream, bool allow_static, int counter_slot) {
    solve_fused3(plan, g, first, batch, prm, sing_detect, verify, stream, allow_static, counter_slot);
}
}  // namespace socp

#ifdef SOCP_PHASE_TIMING
// profiling build only (not declared in include/socp_b200.h)
extern "C" int socp_b200_debug_phase_clocks3(unsigned long long* out16, int reset) {
    if (out16) cudaMemcpyFromSymbol(out16, socp::g_phase_clk3, sizeof(unsigned long long) * 16);
    if (reset) {
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(socp::g_phase_clk3, z, sizeof z);
    }
    return 0;
}
#endif
