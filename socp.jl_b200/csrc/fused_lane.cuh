// fused_lane.cuh -- whole-solve kernel for batches of TINY problems (n <= 16, p = 0: BASELINE.json C3, n = 12 with ten
// second-order cones of dimension 4): ONE LANE PER PROBLEM.
//
// Reference path: the same Mehrotra predictor-corrector as the other whole-solve kernels (src/solver.jl:68-152 on
// src/densesolver.jl:41-90, src/scalings.jl:22-141, src/mats.jl:1-86, src/vectors.jl:54-125), state machine of
// fused_v2.cuh (initial point = one factor + solve with W = I; affine and combined directions share one code path).
//
// Why another kernel.  A problem of this size is ~13 kFLOP per iteration: a warp (let alone a team of warps) per
// problem spends its instructions on shuffles, partial vectors and shared-memory round trips -- fused_v2 on C3 issues
// ~5 000 warp instructions per problem-iteration for ~200 instructions' worth of FMAs, and its shared-memory pipe is
// the busiest unit.  Here every lane runs the whole scalar algorithm on its own problem: no shuffles,
// no barriers, no reductions; a warp instruction advances LPW problems at once.
//
//   * lane-interleaved storage everywhere: element e of the problem of lane l lives at base[e * LPW + l], so every
//     warp access is one contiguous run (shared memory: conflict free; global: fully coalesced);
//   * shared memory holds what every phase touches (lam, wbar, k0, k2, u, four scalars per cone: 240 doubles for C3)
//     plus the lane's slice of the ring below; everything touched once or twice per slot lives in a per-warp global
//     workspace that stays L2 resident (G, h, c, the Cholesky factor, x, dx, dz, s, z, the corrector term: 6.4 KB per
//     problem, 148 SMs x 96 problems = 90 MB) and is read with batched loads; H (n(n+1)/2 packed) and the n-vectors of
//     a solve live in registers;
//   * G streams through a per-lane cp.async ring in shared memory (fl_stream_rows): a lane has nobody to hide its L2
//     latency behind, and the ring costs no registers;
//   * the reduced KKT matrix is accumulated row by row, H = sum_r d_r g_r g_r' + sum_c h_c h_c' (closed form of
//     G'W^-2 G, src/densesolver.jl:42-43), factored in registers (LL', src/densesolver.jl:47) and solved by
//     substitution; the factor of the affine direction is parked in the workspace for the combined direction;
//   * lanes pull problems from an atomic counter as they finish (iteration counts vary 6..19 inside a batch); the
//     warp copies a new problem into the lane's workspace slot cooperatively (coalesced reads);
//   * slots alternate F (factor: initial point or affine direction) and N (no factor: combined direction), so lanes
//     of one warp stay in step: a lane that has just taken a problem runs its initial point in an F slot, idles
//     through the next N slot and joins the others at the following F slot;
//   * throughput follows the number of problems in flight per SM (measured: two full warps = four half warps at 64
//     problems per SM), which shared memory sets: 2400 B per problem -> 96 problems per SM, three full warps.  LPW
//     (lanes in use per warp) and the ring depth stay template parameters for the measurements in profiles/.
//
// Restrictions: p = 0, no `sing` problems (callers fall back to fused_v2 / the tiled path); layouts of the family
// "positive-orthant rows first, then second-order cones back to back" (n <= 16, cone dimensions 2..8, at most six runs of
// equal cones).  BASELINE.json's C3 (and a small test layout) are instantiated at compile time in fused_lane.cu; any
// other layout of the family is specialised at run time with NVRTC on first use (lane_jit.cu: the device code lives in
// fused_lane_dev.cuh, free of host headers).
#pragma once
#include "fused_lane_dev.cuh"
#include <vector>
#include <algorithm>
#include <cstdlib>
#include <utility>

namespace socp {

// does the caller's layout have the shape LaneDims<N, KPOC, NSOC, SDIM, .> describes (orthant rows first, then equal cones)?
inline bool lane_layout_matches(int N, int KPOC, int NSOC, int SDIM, int n, int p, int k, const std::vector<int>& kind,
                                const std::vector<int>& offs, const std::vector<int>& dim) {
    if (n != N || p != 0 || k != KPOC + NSOC * SDIM) return false;
    int kpoc = 0, nsoc = 0;
    for (size_t i = 0; i < kind.size(); ++i) {
        if (kind[i] == KIND_POC) {
            if (nsoc) return false;                   // orthant rows first
            kpoc += dim[i];
        } else {
            if (dim[i] != SDIM || offs[i] != KPOC + nsoc * SDIM) return false;
            ++nsoc;
        }
    }
    return kpoc == KPOC && nsoc == NSOC;
}
// the family of the kernel: orthant rows first, then second-order cones back to back; returns the cones as groups of
// consecutive equal dimensions (count, dim) and the number of orthant rows
inline bool lane_family(int p, int k, const std::vector<int>& kind, const std::vector<int>& offs, const std::vector<int>& dim,
                        int& kpoc, std::vector<std::pair<int, int>>& groups) {
    kpoc = 0;
    groups.clear();
    if (p != 0) return false;
    int at = 0;
    bool soc_seen = false;
    for (size_t i = 0; i < kind.size(); ++i) {
        if (offs[i] != at) return false;
        if (kind[i] == KIND_POC) {
            if (soc_seen) return false;
            kpoc += dim[i];
        } else {
            soc_seen = true;
            if (!groups.empty() && groups.back().second == dim[i]) ++groups.back().first;
            else groups.push_back({1, dim[i]});
        }
        at += dim[i];
    }
    return at == k;
}
// problems in flight per SM: what the shared memory holds, in whole warps, at most four warps
inline int lane_pps(int sm_per_lane, int dev_smem) { return std::min(128, (int)(dev_smem / (sm_per_lane * sizeof(double))) / 32 * 32); }

using LaneC3 = LaneDims<12, 0, 10, 4, 1>;   // BASELINE.json C3: 2400 B of shared memory per problem, 96 problems per SM
using LaneC3r2 = LaneDims<12, 0, 10, 4, 2>; // the same with a ring of twice the depth: 2880 B, 64 (80) problems per SM
using LaneT1 = LaneDims<6, 5, 3, 3, 2>;     // small mixed layout (orthant block + cones): keeps the generic code honest

constexpr int FL_WS_SETS = 2;               // launches that may overlap use different workspace sets

struct FLPlan {
    bool fits = false;
    int shape = 0;          // 1: LaneC3, 2: LaneC3r2, 3: LaneT1; 100: specialised at run time (lane_jit.cu)
    int jn = 0, jkpoc = 0, jrs = 1;                            // the layout of shape 100: orthant rows, then
    std::vector<std::pair<int, int>> jgroups;                  // groups of (count, dimension) equal cones
    void* jit_fn = nullptr; // ... and its kernel once compiled (CUfunction)
    bool jit_tried = false;
    int pps = 64;           // problems in flight per SM (a multiple of 32)
    int lpw = 32;           // lanes in use per warp (measured on C3: 32 = 16 > 8 -- throughput follows the problems in
                            // flight per SM, not the number of instruction streams; profiles/r02_lane_c3_lpw_sweep.txt)
    int num_sms = 148;
    int deg = 0;
    size_t smem = 0;
    size_t ws_doubles = 0;  // per workspace set
    double* d_ws = nullptr; // FL_WS_SETS sets
    int* d_counter = nullptr;
};

inline void fl_plan(FLPlan& P, int n, int p, int k, const std::vector<int>& kind, const std::vector<int>& offs,
                    const std::vector<int>& dim, int dev_smem, int sms) {
    P.fits = false;
    P.shape = 0;
    int spl = 0, wpl = 0;
    const char* rs2 = getenv("SOCP_B200_LANE_RS2");      // experiment switch: the deeper ring, fewer problems per SM
    if (lane_layout_matches(12, 0, 10, 4, n, p, k, kind, offs, dim)) {
        if (rs2 && atoi(rs2)) { P.shape = 2; spl = LaneC3r2::SM_PER_LANE; wpl = LaneC3r2::WS_PER_LANE; P.pps = lane_pps(LaneC3r2::SM_PER_LANE, dev_smem); }
        else { P.shape = 1; spl = LaneC3::SM_PER_LANE; wpl = LaneC3::WS_PER_LANE; P.pps = lane_pps(LaneC3::SM_PER_LANE, dev_smem); }
    } else if (lane_layout_matches(6, 5, 3, 3, n, p, k, kind, offs, dim)) {
        P.shape = 3; spl = LaneT1::SM_PER_LANE; wpl = LaneT1::WS_PER_LANE; P.pps = lane_pps(LaneT1::SM_PER_LANE, dev_smem);
    } else {
        // no compile-time instantiation: a layout of the same family (orthant rows first, then equal second-order cones,
        // p = 0, n <= 16; up to n = 12 the packed H fits the registers) is specialised at run time with NVRTC (lane_jit.cu)
        if (getenv("SOCP_B200_NO_LANE_JIT") || p != 0 || n < 1 || n > 16 || k > 96) return;
        int kpoc = 0, nsoc = 0;
        std::vector<std::pair<int, int>> groups;
        if (!lane_family(p, k, kind, offs, dim, kpoc, groups) || groups.empty() || groups.size() > 6) return;
        for (const auto& g : groups) {
            if (g.second < 2 || g.second > 8) return;
            nsoc += g.first;
        }
        const int np = (n + 1) / 2 * 2;
        const int state = 5 * k + kpoc + 4 * nsoc;                       // LaneDimsG::SM_STATE
        P.jrs = (k % 2 == 0 && lane_pps(state + 2 * 5 * np, dev_smem) >= 128) ? 2 : 1;
        spl = state + P.jrs * 5 * np;                                    // + LaneDimsG::SM_RING
        wpl = k * np + k + n + n * (n + 1) / 2 + 2 * n + 4 * k;          // LaneDimsG::WS_PER_LANE
        P.pps = lane_pps(spl, dev_smem);
        if (P.pps < 32) return;
        P.jgroups = groups;
        P.shape = 100;
        P.jn = n; P.jkpoc = kpoc;
    }
    // only what fused_lane.cu instantiates: 96 (or 64) problems per SM for C3, 64 with the deeper ring, 128 for T1
    const char* pe = getenv("SOCP_B200_LANE_PPS");       // experiment switch: 64 problems per SM on the C3 layout
    if (P.shape == 1) P.pps = (P.pps >= 96 && !(pe && atoi(pe) == 64)) ? 96 : (P.pps >= 64 ? 64 : 0);
    else if (P.shape == 2) P.pps = P.pps >= 64 ? 64 : 0;
    else if (P.shape == 3) P.pps = P.pps >= 128 ? 128 : 0;
    if (P.pps == 0) return;
    P.num_sms = sms;
    P.deg = 0;
    for (size_t i = 0; i < kind.size(); ++i) P.deg += kind[i] == KIND_POC ? dim[i] : 1;
    P.smem = (size_t)spl * P.pps * sizeof(double);
    if (P.smem > (size_t)dev_smem) return;
    P.ws_doubles = (size_t)wpl * P.pps * sms;
    P.fits = true;
}

// grid and per-warp lane cap of a launch over `batch` problems: a small batch is spread over the SMs instead of
// filling the first CTAs
inline void fl_grid(const FLPlan& plan, int batch, int lpw, int& grid, int& cap) {
    const int nwarp = plan.pps / lpw;
    grid = std::max(1, std::min(plan.num_sms, (batch + plan.pps - 1) / plan.pps));
    cap = std::min(lpw, (batch + grid * nwarp - 1) / (grid * nwarp));
}

#ifndef SOCP_SIMT_EMU
template <class D, int LPW, int NWARP>
inline void fused_lane_launch(const FLPlan& plan, FLArgs args, cudaStream_t stream) {
    int grid;
    fl_grid(plan, args.batch, LPW, grid, args.cap);
    cudaFuncSetAttribute(k_fused_lane<D, LPW, NWARP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
    k_fused_lane<D, LPW, NWARP><<<grid, NWARP * 32, plan.smem, stream>>>(args);
}

// run-time specialisation (lane_jit.cu): the kernel of a shape-100 plan, or nullptr; its launch
void* lane_jit_get(const FLPlan& plan, int device);
bool lane_jit_launch(void* fn, const FLPlan& plan, FLArgs args, cudaStream_t stream);

// Solves problems [first, first + batch) of the shard.  ws_set: which workspace set / counter this launch uses
// (launches that may overlap need different ones).
void solve_fused_lane_ext(FLPlan& plan, const Ws& g, int first, int batch, const LoopParams& lp, cudaStream_t stream,
                          int ws_set);
#endif

}  // namespace socp
