// emu_fused_lane.cpp -- host build of the lane-per-problem whole-solve kernel k_fused_lane
// (socp.jl_b200/csrc/fused_lane.cuh) against the SIMT emulator.  TEST INFRASTRUCTURE ONLY (see emu_fused3.cpp).
#define SOCP_SIMT_EMU 1
#include "fused_lane.cuh"

#include <vector>

using namespace socp;

template <class D, int LPW, int NWARP>
static void run(const FLPlan& P, FLArgs a, int grid_cap, int order) {
    int grid;
    fl_grid(P, a.batch, LPW, grid, a.cap);
    if (grid_cap > 0) grid = std::min(grid, grid_cap);
    std::vector<double> ws((size_t)D::WS_PER_LANE * P.pps * grid, std::nan(""));
    a.ws = ws.data();
    simt_emu::LaunchCfg cfg;
    cfg.grid = (unsigned)grid;
    cfg.block = NWARP * 32;
    cfg.smem = P.smem;
    cfg.order = order;
    simt_emu::launch(cfg, [&]() { k_fused_lane<D, LPW, NWARP>(a); });
}

// Returns 0, -1 when no instantiation takes the layout.  grid_cap > 0 limits the number of CTAs (more problems per lane).
extern "C" int emu_fused_lane_solve(int n, int k, int ncones, const int* kind, const int* offs, const int* dim, int batch,
                                    const double* c, const double* G, int64_t sG, const double* h, int lpw, int order,
                                    int grid_cap, int max_iter, double tol, double damp, double init_eps, double* x,
                                    double* z, double* s, int* status, int* iters, double* pobj, double* dobj) {
    FLPlan P;
    fl_plan(P, n, 0, k, std::vector<int>(kind, kind + ncones), std::vector<int>(offs, offs + ncones),
            std::vector<int>(dim, dim + ncones), 227 * 1024, 148);
    // (shape 100: specialised at run time with NVRTC on the device only; one mixed-dimension layout of that kind is
    // instantiated here so that the grouped cone loops are covered on the CPU)
    using LaneT2 = LaneDimsG<6, 2, 1, ConeGroup<2, 3>, ConeGroup<1, 5>>;
    const bool t2 = P.fits && P.shape == 100 && P.jn == 6 && P.jkpoc == 2 && P.jrs == 1 &&
                    P.jgroups == std::vector<std::pair<int, int>>{{2, 3}, {1, 5}};
    if (!P.fits || (P.shape == 100 && !t2)) return -1;
    int counter = 0;
    std::vector<int> active(batch, 1), fail(batch, 0);
    FLArgs a;
    a.c = c; a.G = G; a.h = h; a.sG = sG;
    a.x = x; a.z = z; a.s = s; a.pobj = pobj; a.dobj = dobj;
    a.status = status; a.iters = iters; a.active = active.data(); a.fail = fail.data();
    a.ws = nullptr; a.counter = &counter;
    a.first = 0; a.batch = batch; a.cap = 0; a.deg = P.deg;
    a.prm = LoopParams{max_iter, tol, damp, init_eps};
#define FL_RUN(D, PPS)                                                     \
    switch (lpw) {                                                         \
        case 16: run<D, 16, PPS / 16>(P, a, grid_cap, order); break;       \
        case 8: run<D, 8, PPS / 8>(P, a, grid_cap, order); break;          \
        default: run<D, 32, PPS / 32>(P, a, grid_cap, order); break;       \
    }
    if (t2) {
        if (P.pps != 128) return -1;
        FL_RUN(LaneT2, 128)
        return 0;
    }
    if (P.shape == 1 && P.pps == 96) { FL_RUN(LaneC3, 96) }
    else if (P.shape == 1) { FL_RUN(LaneC3, 64) }
    else if (P.shape == 2) { FL_RUN(LaneC3r2, 64) }
    else { FL_RUN(LaneT1, 128) }
    return 0;
}

// the host-side plan of the lane kernel for a layout: out = fits, shape, pps, smem, rs, #groups, then (count, dim) pairs
extern "C" int emu_fused_lane_plan(int n, int p, int k, int ncones, const int* kind, const int* offs, const int* dim, int* out) {
    FLPlan P;
    fl_plan(P, n, p, k, std::vector<int>(kind, kind + ncones), std::vector<int>(offs, offs + ncones),
            std::vector<int>(dim, dim + ncones), 227 * 1024, 148);
    out[0] = P.fits; out[1] = P.shape; out[2] = P.pps; out[3] = (int)P.smem; out[4] = P.jrs; out[5] = (int)P.jgroups.size();
    for (size_t i = 0; i < P.jgroups.size() && i < 8; ++i) { out[6 + 2 * i] = P.jgroups[i].first; out[7 + 2 * i] = P.jgroups[i].second; }
    return 0;
}
