// fused_lane.cu -- translation unit of the lane-per-problem whole-solve kernel (fused_lane.cuh): tiny problems
// (n <= 16, p = 0; BASELINE.json C3), one problem per lane.
#include "fused_lane.cuh"
#include <cstdlib>

namespace socp {

// (lanes per warp, warps per CTA) of the instantiations: full warps by default; half and quarter warps with the
// same number of problems per SM are experiment switches (SOCP_B200_LANE_LPW, profiles/)
template <class D, int PPS>
static void fl_dispatch(const FLPlan& plan, const FLArgs& args, int lpw, cudaStream_t stream) {
    switch (lpw) {
        case 16: fused_lane_launch<D, 16, PPS / 16>(plan, args, stream); break;
        case 8: if (PPS / 8 <= 8) { fused_lane_launch<D, 8, (PPS / 8 <= 8 ? PPS / 8 : 8)>(plan, args, stream); break; }
        default: fused_lane_launch<D, 32, PPS / 32>(plan, args, stream); break;
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
    if (lpw != 16 && lpw != 8) lpw = 32;
    if (lpw == 8 && plan.pps > 64) lpw = 16;             // 255 registers x 8 warps fill the register file
    if (plan.shape == 100) { lane_jit_launch(plan.jit_fn, plan, a, stream); return; }      // full warps only
    if (plan.shape == 1 && plan.pps == 96) fl_dispatch<LaneC3, 96>(plan, a, lpw, stream);
    else if (plan.shape == 1) fl_dispatch<LaneC3, 64>(plan, a, lpw, stream);
    else if (plan.shape == 2) fl_dispatch<LaneC3r2, 64>(plan, a, lpw, stream);
    else fl_dispatch<LaneT1, 128>(plan, a, lpw, stream);
}

}  // namespace socp
