// fused2.cu -- translation unit of the first-generation whole-solve kernels (fused_v2.cuh): one-warp teams (n <= 16)
// and the 4- / 8-warp fallback.  Separate from solver.cu so that the three units compile in parallel.
#include "fused_v2.cuh"

namespace socp {
void solve_fused2_ext(F2Plan& plan, const Ws& g, int first, int batch, int max_iter, double tol, double step_damp,
                      double init_eps, cudaStream_t stream, bool allow_static, int counter_slot) {
    solve_fused2(plan, g, first, batch, max_iter, tol, step_damp, init_eps, stream, allow_static, counter_slot);
}
}  // namespace socp
