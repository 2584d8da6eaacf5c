// fused_small.cuh -- whole-solve kernel for layouts that fit in shared memory.
// (placeholder: the plan reports "does not fit" until the kernel lands)
#pragma once
#include "tiled_kernels.cuh"
#include <vector>

namespace socp {

struct FusedPlan {
    bool fits = false;
};

inline void fused_plan(FusedPlan& plan, int n, int p, int k, const std::vector<int>& kind,
                       const std::vector<int>& offs, const std::vector<int>& dim, int device) {
    (void)n; (void)p; (void)k; (void)kind; (void)offs; (void)dim; (void)device;
    plan.fits = false;
}

inline void solve_fused(const FusedPlan&, const Ws&, int, int, double, double, double, cudaStream_t) {}

}  // namespace socp
