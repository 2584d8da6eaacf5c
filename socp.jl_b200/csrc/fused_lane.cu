// fused_lane.cu -- translation unit of the lane-per-problem whole-solve kernel (fused_lane.cuh): tiny problems
// (n <= 16, p = 0; BASELINE.json C3), one problem per lane.
#include "fused_lane.cuh"
#include <cstdlib>

namespace socp {

template <class D>
static void fl_dispatch(const FLPlan& plan, const FLArgs& args, int lpw, cudaStream_t stream) {
    switch (lpw) {
        case 32: fused_lane_launch<D, 32>(plan, args, stream); break;
        case 16: fused_lane_launch<D, 16>(plan, args, stream); break;
        case 4: fused_lane_launch<D, 4>(plan, args, stream); break;
        default: fused_lane_launch<D, 8>(plan, args, stream); break;
    }
}

void solve_fused_lane_ext(FLPlan& plan, const Ws& g, int first, int batch, const LoopParams& lp, cudaStream_t stream,
                          int ws_set) {
    ws_set &= FL_WS_SETS - 1;
    cudaMemsetAsync(plan.d_counter + ws_set, 0, sizeof(int), stream);
    FLArgs a;
    a.c = g.c; a.G = g.G; a.h = g.h; a.sG = g.sG;
    a.x = g.x; a.z = g.z; a.s = g.s; a.pobj = g.pobj; a.dobj = g.dobj;
    a.status = g.status; a.iters = g.iters; a.active = g.active; a.fail = g.fail;
    a.ws = plan.d_ws + (size_t)ws_set * plan.ws_doubles;
    a.counter = plan.d_counter + ws_set;
    a.first = first; a.batch = batch; a.cap = 0; a.deg = plan.deg;
    a.prm = lp;
    int lpw = plan.lpw;
    if (const char* e = getenv("SOCP_B200_LANE_LPW")) lpw = atoi(e);      // experiment switch (profiles/)
    if (plan.shape == 1) fl_dispatch<LaneC3>(plan, a, lpw, stream);
    else fl_dispatch<LaneT1>(plan, a, lpw, stream);
}

}  // namespace socp
