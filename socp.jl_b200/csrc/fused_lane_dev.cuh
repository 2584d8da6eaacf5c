// fused_lane_dev.cuh -- device side of the lane-per-problem whole-solve kernel (see fused_lane.cuh for the design).
// No host code and no host headers: this file is also what lane_jit.cu hands to NVRTC to specialise the kernel for a
// layout at run time -- the GPU analogue of the reference's compile-time specialisation on the cone tuple.
#pragma once
#include "fused_common.cuh"

namespace socp {

// COUNT consecutive second-order cones of dimension DIM
template <int COUNT_, int DIM_>
struct ConeGroup { static constexpr int COUNT = COUNT_, DIM = DIM_; };
// what the kernel's cone loops see of one group: its first row in a k-vector and the index of its first cone
template <int COUNT_, int DIM_, int OFF_, int C0_>
struct ConeGroupAt { static constexpr int COUNT = COUNT_, DIM = DIM_, OFF = OFF_, C0 = C0_; };

template <int N_, int KPOC_, int RS_, class... Gs>
struct LaneDimsG {
    static constexpr int N = N_, KPOC = KPOC_;
    static constexpr int NSOC = (0 + ... + Gs::COUNT);    // second-order cones: groups of equal cones, in order
    static constexpr int K = KPOC_ + (0 + ... + (Gs::COUNT * Gs::DIM));
    static constexpr int NH = N_ * (N_ + 1) / 2;
    static constexpr int NP = (N_ + 1) / 2 * 2;           // row stride of G in the workspace (rows are runs of double2)
    static constexpr int RS = RS_;                        // rows per stage of the G ring
    static_assert(K % RS_ == 0, "ring stages must tile the rows of G");
    static constexpr int NS = 5;                          // stages of the G ring
    // shared memory, doubles per lane: k-vectors, then iwb (orthant rows), cone scalars; then the lane's slice of the
    // ring that G streams through (RS * NS rows of NP doubles)
    enum { V_LAM = 0, V_WB, V_K0, V_K2, V_U, NVEC };
    static constexpr int O_IWB = NVEC * K;
    static constexpr int O_CS = O_IWB + KPOC_;            // 4 per cone: eta, 1/eta, 1/(1+w0), |lam_1|^2
    static constexpr int SM_STATE = O_CS + 4 * NSOC;
    static constexpr int SM_RING = RS * NS * NP;
    static constexpr int SM_PER_LANE = SM_STATE + SM_RING;
    // global workspace (L2 resident), doubles per lane: G (row-major k x NP, as double2 pairs), h, c, the packed factor,
    // and the vectors touched once or twice per slot: x, dx, dz, the corrector term of ds, s, z
    static constexpr int W_G = 0, W_H = K * NP, W_C = W_H + K, W_L = W_C + N_, W_X = W_L + NH, W_DX = W_X + N_,
                         W_DZ = W_DX + N_, W_DSC = W_DZ + K, W_S = W_DSC + K, W_Z = W_S + K;
    static constexpr int WS_PER_LANE = W_Z + K;
    // f(ConeGroupAt<...>{}) for every group, in order
    template <class F>
    __device__ __forceinline__ static void for_groups(F&& f) { walk<KPOC_, 0, Gs...>(f); }
    template <int OFF, int C0, class F>
    __device__ __forceinline__ static void walk(F&) {}
    template <int OFF, int C0, class G, class... Rest, class F>
    __device__ __forceinline__ static void walk(F& f) {
        f(ConeGroupAt<G::COUNT, G::DIM, OFF, C0>{});
        walk<OFF + G::COUNT * G::DIM, C0 + G::COUNT, Rest...>(f);
    }
};
// the common case: NSOC equal cones of dimension SDIM
template <int N_, int KPOC_, int NSOC_, int SDIM_, int RS_>
using LaneDims = LaneDimsG<N_, KPOC_, RS_, ConeGroup<NSOC_, SDIM_>>;



struct FLArgs {
    // problem data and results of the shard (batch slowest), as in Ws
    const double *c, *G, *h;
    int64_t sG;
    double *x, *z, *s, *pobj, *dobj;
    int *status, *iters, *active, *fail;
    double* ws;
    int* counter;
    int first, batch, cap, deg;
    LoopParams prm;
};

enum { FL_FREE = -1, FL_DONE = -2 };

#ifdef SOCP_SIMT_EMU
inline unsigned __ballot_sync(unsigned, int pred) {      // a lane that has returned reads back as the caller's own value:
    unsigned r = 0;                                      // every lane votes with its own number so that it cannot pass
    const int mine = pred ? simt_emu::cur_thread()->lane + 1 : 0;      // for another lane's vote
    for (int l = 0; l < 32; ++l)
        if (__shfl_sync(0xffffffffu, mine, l) == l + 1) r |= 1u << l;
    return r;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
#define FL_SMEM() reinterpret_cast<double*>(emu_dyn_smem())
#else
#define FL_SMEM() fl_sm
#endif

// 16-byte asynchronous copy global -> shared (LDGSTS): no registers in between, completion by commit groups
__device__ __forceinline__ void fl_cp16(double2* dst, const double2* src) {
#ifdef SOCP_SIMT_EMU
    *dst = *src;
#else
    const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void fl_commit() {
#ifndef SOCP_SIMT_EMU
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int PENDING>
__device__ __forceinline__ void fl_wait() {
#ifndef SOCP_SIMT_EMU
    asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
#endif
}
// block `blk` (RS rows) of this lane's G into stage blk % NS of its ring slice
template <int NQ, int LPW, int RS, int NS>
__device__ __forceinline__ void fl_ring_issue(const double2* __restrict__ G2, double2* ring, int blk, int stage) {
#pragma unroll
    for (int i = 0; i < RS; ++i)
#pragma unroll
        for (int q = 0; q < NQ; ++q) fl_cp16(ring + ((stage * RS + i) * NQ + q) * LPW, G2 + ((blk * RS + i) * NQ + q) * LPW);
}
// the first NS - 1 blocks of a pass: issued as early as possible (right after the previous pass) so that their latency
// hides behind whatever runs between the passes
template <int K, int NP, int LPW, int RS, int NS>
__device__ __forceinline__ void fl_ring_prime(const double2* __restrict__ G2, double2* ring) {
    constexpr int NQ = NP / 2, NBLK = K / RS;
#pragma unroll
    for (int blk = 0; blk < NS - 1; ++blk) {
        if (blk < NBLK) fl_ring_issue<NQ, LPW, RS, NS>(G2, ring, blk, blk);
        fl_commit();
    }
}
// Streams the K rows of this lane's G (double2 pairs, element q of row r at G2[(r * NQ + q) * LPW]) through the lane's
// slice of a shared-memory ring (NS stages of RS rows, cp.async): up to NS - 1 stages are in flight while f(r, g) runs
// on the rows of the current one -- a lane has nobody to hide its L2 latency behind (one or two warps per SM
// sub-partition), and the ring costs no registers (H alone takes 156 of the 255).  The ring slice is private to the
// lane, so there is no synchronisation with other lanes.  The ring is primed on entry (by the previous pass, or by
// the lane when it took its problem) and primed again on exit.
template <int K, int NP, int LPW, int RS, int NS, class F>
__device__ __forceinline__ void fl_stream_rows(const double2* __restrict__ G2, double2* ring, F f) {
    static_assert(K % RS == 0, "stages must tile the rows");
    constexpr int NQ = NP / 2, NBLK = K / RS;
    int stage = 0, nstage = NS - 1;
#pragma unroll 1
    for (int blk = 0; blk < NBLK; ++blk) {
        if (blk + NS - 1 < NBLK) fl_ring_issue<NQ, LPW, RS, NS>(G2, ring, blk + NS - 1, nstage);
        fl_commit();
        fl_wait<NS - 1>();
#pragma unroll
        for (int i = 0; i < RS; ++i) {
            double g[NP];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const double2 v = ring[((stage * RS + i) * NQ + q) * LPW];
                g[2 * q] = v.x; g[2 * q + 1] = v.y;
            }
            f(blk * RS + i, g);
        }
        stage = stage + 1 == NS ? 0 : stage + 1;
        nstage = nstage + 1 == NS ? 0 : nstage + 1;
    }
    fl_ring_prime<K, NP, LPW, RS, NS>(G2, ring);      // every pass starts at row 0: the next one is on its way
}

template <class D, int LPW, int NWARP>
__global__ void __launch_bounds__(NWARP * 32, 1) k_fused_lane(const FLArgs a) {
#ifndef SOCP_SIMT_EMU
    extern __shared__ __align__(16) double fl_sm[];
#endif
    constexpr int N = D::N, K = D::K, KPOC = D::KPOC, NH = D::NH, NP = D::NP, RS = D::RS, NS = D::NS;
    constexpr unsigned MASK = LPW == 32 ? 0xffffffffu : ((1u << LPW) - 1u);
    // measured on C3: interleaving five cones per loop body instead of two, and four partial sums per row instead of
    // two, made the kernel slower (16.5M -> 13.4M problems/s at 16 lanes per warp): code size, not chain length
    constexpr int CU = 2;
    const int tid = (int)(unsigned)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (lane >= LPW) return;
    const int gwarp = (int)(unsigned)blockIdx.x * NWARP + warp;
    double* const S = FL_SMEM() + (size_t)warp * D::SM_PER_LANE * LPW + lane;
    double* const Wb = a.ws + (size_t)gwarp * D::WS_PER_LANE * LPW;
    double* const W = Wb + lane;
    const double2* const G2 = reinterpret_cast<const double2*>(Wb) + lane;
    double2* const ring = reinterpret_cast<double2*>(S - lane + D::SM_STATE * LPW) + lane;
    const LoopParams prm = a.prm;
#define SV(v, i) S[((v) * K + (i)) * LPW]
#define SO(o, i) S[((o) + (i)) * LPW]
#define WO(o, i) W[((o) + (i)) * LPW]
#define TRI(i, j) ((i) * ((i) + 1) / 2 + (j))

    int phase = FL_FREE, b = 0, iters = 0, status = ST_RUNNING;
    bool need_top = false, dead = false, exhausted = false, fslot = true;
    double sc = 1.0, ll = 0.0, smu_l = 0.0;

    for (;;) {
        if (fslot) {
            // ------------------------------------------------ top of a Mehrotra iteration, src/solver.jl:105-126
            if (phase == 1 && need_top) {
                need_top = false;
                if (iters >= prm.max_iter) { status = ST_MAXITER; phase = FL_DONE; }
                else {
                    double gap = 0.0, llacc = 0.0;
                    int fl = 0;
#pragma unroll 2
                    for (int i = 0; i < KPOC; ++i) {                                    // src/scalings.jl:22-30
                        const double si = WO(D::W_S, i), zi = WO(D::W_Z, i);
                        SV(D::V_U, i) = zi;                           // z and h - s for the residual pass below
                        SV(D::V_K2, i) = WO(D::W_H, i) - si;
                        const double q = si * fast_rcp(zi), qi = zi * fast_rcp(si), pz = si * zi;
                        fl |= !(q >= 0.0) | !(pz >= 0.0);
                        const double lv = fast_sqrt(pz);
                        SV(D::V_WB, i) = fast_sqrt(q);
                        SO(D::O_IWB, i) = fast_sqrt(qi);
                        SV(D::V_LAM, i) = lv;
                        gap = fma(si, zi, gap);
                        llacc = fma(lv, lv, llacc);
                    }
                    // s, z and h come from the L2 workspace: the loads of the next two cones are in flight under this one
                    D::for_groups([&](auto gtag) {
                        constexpr int SDIM = decltype(gtag)::DIM, GCNT = decltype(gtag)::COUNT, GOFF = decltype(gtag)::OFF, GC0 = decltype(gtag)::C0;
                        double sn[SDIM], zn[SDIM], hn[SDIM], sm2[SDIM], zm2[SDIM], hm2[SDIM];
#pragma unroll
                        for (int e = 0; e < SDIM; ++e) {
                            sn[e] = WO(D::W_S, GOFF + e); zn[e] = WO(D::W_Z, GOFF + e); hn[e] = WO(D::W_H, GOFF + e);
                            sm2[e] = GCNT > 1 ? WO(D::W_S, GOFF + SDIM + e) : 0.0; zm2[e] = GCNT > 1 ? WO(D::W_Z, GOFF + SDIM + e) : 0.0;
                            hm2[e] = GCNT > 1 ? WO(D::W_H, GOFF + SDIM + e) : 0.0;
                        }
#pragma unroll CU
                        for (int cc = 0; cc < GCNT; ++cc) {                                    // src/scalings.jl:32-99
                        const int c = GC0 + cc, o = GOFF + cc * SDIM;
                        double sv[SDIM], zv[SDIM];
#pragma unroll
                        for (int e = 0; e < SDIM; ++e) {
                            sv[e] = sn[e]; zv[e] = zn[e];
                            SV(D::V_U, o + e) = zn[e];                 // z and h - s for the residual pass below
                            SV(D::V_K2, o + e) = hn[e] - sn[e];
                            sn[e] = sm2[e]; zn[e] = zm2[e]; hn[e] = hm2[e];
                        }
                        if (cc + 2 < GCNT) {
#pragma unroll
                            for (int e = 0; e < SDIM; ++e) {
                                sm2[e] = WO(D::W_S, o + 2 * SDIM + e); zm2[e] = WO(D::W_Z, o + 2 * SDIM + e);
                                hm2[e] = WO(D::W_H, o + 2 * SDIM + e);
                            }
                        }
                        double ss = 0.0, zz = 0.0, sz = 0.0;
#pragma unroll
                        for (int e = 1; e < SDIM; ++e) { ss = fma(sv[e], sv[e], ss); zz = fma(zv[e], zv[e], zz); sz = fma(sv[e], zv[e], sz); }
                        const double onrms = sv[0] * sv[0] - ss, onrmz = zv[0] * zv[0] - zz;      // :39-45
                        fl |= !(onrms >= 0.0) | !(onrmz >= 0.0);
                        const double is = fast_rsqrt(onrms), iz = fast_rsqrt(onrmz);    // :46-49
                        const double nrms = onrms * is, nrmz = onrmz * iz;
                        const double sb0 = sv[0] * is, zb0 = zv[0] * iz;
                        const double ns = sz * (is * iz) + zb0 * sb0;                   // :53-56
                        const double g2 = (1.0 + ns) / 2.0;
                        fl |= !(g2 >= 0.0);
                        const double rg = fast_rsqrt(g2);
                        const double gamma = g2 * rg, ig = 0.5 * rg;                    // :57, :64
                        const double eta = fast_sqrt(nrms * iz);                        // :68
                        const double ie = fast_rcp(eta);
                        const double tmv1 = fast_sqrt(nrms * nrmz);                     // :91
                        const double mult = tmv1 * fast_rcp(zb0 + sb0 + 2.0 * gamma);   // :92
                        const double csf = gamma + zb0, czf = gamma + sb0;              // :93-94
                        double llt = 0.0;
#pragma unroll
                        for (int e = 1; e < SDIM; ++e) {
                            const double sb = sv[e] * is, zb = zv[e] * iz;
                            const double lv = (sb * csf + zb * czf) * mult;             // :95-97
                            SV(D::V_WB, o + e) = (sb - zb) * ig;                        // :62, :64
                            SV(D::V_LAM, o + e) = lv;
                            llt = fma(lv, lv, llt);
                        }
                        const double w0 = (sb0 + zb0) * ig, l0 = gamma * tmv1;
                        SV(D::V_WB, o) = w0;                                            // :60
                        SV(D::V_LAM, o) = l0;                                           // :98
                        SO(D::O_CS, 4 * c + 0) = eta;
                        SO(D::O_CS, 4 * c + 1) = ie;
                        SO(D::O_CS, 4 * c + 2) = fast_rcp(1.0 + w0);
                        SO(D::O_CS, 4 * c + 3) = llt;
                        gap += sv[0] * zv[0] + sz;
                        llacc += l0 * l0 + llt;
                    }
                    });
                    // negated residuals (:110-118, :125) in one pass over G: dx = -G'z - c, dz = -G x - s + h
                    double rx[N], xr[N];
#pragma unroll
                    for (int j = 0; j < N; ++j) { rx[j] = -WO(D::W_C, j); xr[j] = WO(D::W_X, j); }
                    fl_stream_rows<K, NP, LPW, RS, NS>(G2, ring, [&](int r, const double (&g)[NP]) {
                        const double zr = SV(D::V_U, r);
                        double a0 = 0.0, a1 = 0.0;
#pragma unroll
                        for (int j = 0; j < N; ++j) {
                            rx[j] = fma(-g[j], zr, rx[j]);
                            if (j & 1) a1 = fma(g[j], xr[j], a1); else a0 = fma(g[j], xr[j], a0);
                        }
                        WO(D::W_DZ, r) = SV(D::V_K2, r) - (a0 + a1);
                    });
                    double rxn = 0.0;
#pragma unroll
                    for (int j = 0; j < N; ++j) { rxn = fma(rx[j], rx[j], rxn); WO(D::W_DX, j) = rx[j]; }
                    if (fl) { status = ST_NUMERICAL; phase = FL_DONE; }                 // compute_scaling threw
                    else if (sqrt(rxn) + gap < prm.tol) { status = ST_CONVERGED; phase = FL_DONE; }   // :122-124 (p = 0)
                    else { ll = llacc; sc = 1.0; }
                }
            }
            // ------------------------------------------------ a finished problem: iterate and objectives
            if (phase == FL_DONE) {
                double po = 0.0, dob = 0.0;
                // workspace loads first, in batches (the compiler cannot move them above the stores to a.x / a.z / a.s)
                {
                    double xv[N], cv[N];
#pragma unroll
                    for (int j = 0; j < N; ++j) { xv[j] = WO(D::W_X, j); cv[j] = WO(D::W_C, j); }
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        const double xj = dead ? 0.0 : xv[j];
                        a.x[(int64_t)b * N + j] = xj;
                        po = fma(cv[j], xj, po);
                    }
                }
                constexpr int FB = 20;
#pragma unroll 1
                for (int r0 = 0; r0 < K; r0 += FB) {
                    double hv[FB], zv[FB], sv[FB];
#pragma unroll
                    for (int q = 0; q < FB; ++q) {
                        const bool in = r0 + q < K;
                        hv[q] = in ? WO(D::W_H, r0 + q) : 0.0;
                        zv[q] = in && !dead ? WO(D::W_Z, r0 + q) : 0.0;
                        sv[q] = in && !dead ? WO(D::W_S, r0 + q) : 0.0;
                    }
#pragma unroll
                    for (int q = 0; q < FB; ++q) {
                        const int r = r0 + q;
                        if (r < K) {
                            a.z[(int64_t)b * K + r] = zv[q];
                            a.s[(int64_t)b * K + r] = sv[q];
                            dob = fma(-hv[q], zv[q], dob);
                        }
                    }
                }
                a.pobj[b] = po;
                a.dobj[b] = dob;
                a.status[b] = status;
                a.iters[b] = iters;
                a.active[b] = 0;
                a.fail[b] = (status == ST_NUMERICAL);
                phase = FL_FREE;
            }
            // ------------------------------------------------ free lanes take the next problems of the batch
            {
                const bool want = phase == FL_FREE && lane < a.cap && !exhausted;
                const unsigned need = __ballot_sync(MASK, want);
                if (need) {
                    const int cnt = __popc(need);
                    int base = 0;
                    if (lane == 0) base = atomicAdd(a.counter, cnt);
                    base = __shfl_sync(MASK, base, 0);
                    int myb = want ? base + __popc(need & ((1u << lane) - 1u)) : -1;
                    if (base + cnt >= a.batch) exhausted = true;
                    if (myb >= a.batch) myb = -1;
                    // warp-uniform: the problems of up to NF lanes are copied into their slots together -- all the loads
                    // of the group are issued before its first store, so the warp waits one memory latency per group
                    // and not one per element or per problem (the data comes from DRAM: first touch)
                    constexpr int NF = 4, CH = 16, NHC = (K + N + LPW - 1) / LPW;
                    unsigned m = need;
                    while (m) {
                        int ls[NF], nf = 0;
                        int64_t gbs[NF];
#pragma unroll
                        for (int f = 0; f < NF; ++f) { ls[f] = 0; gbs[f] = 0; }
                        while (m && nf < NF) {
                            const int l = __ffs(m) - 1;
                            m &= m - 1;
                            const int pb = __shfl_sync(MASK, myb, l);
                            if (pb < 0) continue;
#pragma unroll
                            for (int f = 0; f < NF; ++f)
                                if (f == nf) { ls[f] = l; gbs[f] = (int64_t)a.first + pb; }
                            ++nf;
                        }
                        double hc[NF][NHC];
#pragma unroll
                        for (int f = 0; f < NF; ++f)
#pragma unroll
                            for (int i = 0; i < NHC; ++i) {
                                const int e = i * LPW + lane;
                                hc[f][i] = (f < nf && e < K) ? a.h[gbs[f] * K + e]
                                                             : ((f < nf && e < K + N) ? a.c[gbs[f] * N + (e - K)] : 0.0);
                            }
#pragma unroll 1
                        for (int e0 = 0; e0 < K * N; e0 += CH * LPW) {
                            double t[NF][CH];
#pragma unroll
                            for (int f = 0; f < NF; ++f) {
                                const double* Gg = a.G + gbs[f] * a.sG;
#pragma unroll
                                for (int i = 0; i < CH; ++i) {
                                    const int e = e0 + i * LPW + lane;
                                    t[f][i] = (f < nf && e < K * N) ? Gg[e] : 0.0;
                                }
                            }
#pragma unroll
                            for (int f = 0; f < NF; ++f)
#pragma unroll
                                for (int i = 0; i < CH; ++i) {
                                    const int e = e0 + i * LPW + lane;
                                    const int j = e / K, r = e - j * K;
                                    if (f < nf && e < K * N) Wb[((r * (NP / 2) + (j >> 1)) * LPW + ls[f]) * 2 + (j & 1)] = t[f][i];
                                }
                        }
#pragma unroll
                        for (int f = 0; f < NF; ++f) {
                            if (f < nf) {
                                if (NP != N)
                                    for (int r = lane; r < K; r += LPW) Wb[((r * (NP / 2) + (N >> 1)) * LPW + ls[f]) * 2 + 1] = 0.0;
#pragma unroll
                                for (int i = 0; i < NHC; ++i) {
                                    const int e = i * LPW + lane;
                                    if (e < K + N) Wb[(D::W_H + e) * LPW + ls[f]] = hc[f][i];      // h, then c (W_C = W_H + K)
                                }
                            }
                        }
                    }
                    __syncwarp(MASK);
                    if (myb >= 0) {
                        b = a.first + myb; phase = 0; iters = 0; status = ST_RUNNING; need_top = false; dead = false;
                        fl_wait<0>();                                   // whatever the previous problem left in flight
                        fl_ring_prime<K, NP, LPW, RS, NS>(G2, ring);
                    }
                }
            }
            if (!__ballot_sync(MASK, phase != FL_FREE)) break;
        }

        const bool act = fslot ? (phase == 0 || phase == 1) : (phase == 2);
        if (__ballot_sync(MASK, act) && act) {
            // ------------------------------------------------ head of solve_kkt, src/densesolver.jl:61-66 (+ W^-2 of :86)
            if (phase == 0) {
                // initial point (src/solver.jl:68-104): W = I, u = h, k2 = h, dx = -c
#pragma unroll 20
                for (int r = 0; r < K; ++r) {
                    const double hr = WO(D::W_H, r);
                    SV(D::V_U, r) = hr;
                    SV(D::V_K2, r) = hr;
                    SV(D::V_WB, r) = 0.0;
                }
                for (int i = 0; i < KPOC; ++i) { SV(D::V_WB, i) = 1.0; SO(D::O_IWB, i) = 1.0; }
                D::for_groups([&](auto gtag) {
                    constexpr int SDIM = decltype(gtag)::DIM, GCNT = decltype(gtag)::COUNT, GOFF = decltype(gtag)::OFF, GC0 = decltype(gtag)::C0;
                    for (int cc = 0; cc < GCNT; ++cc) {
                        const int c = GC0 + cc;
                    SV(D::V_WB, GOFF + cc * SDIM) = 1.0;
                    SO(D::O_CS, 4 * c + 0) = 1.0; SO(D::O_CS, 4 * c + 1) = 1.0; SO(D::O_CS, 4 * c + 2) = 0.5; SO(D::O_CS, 4 * c + 3) = 0.0;
                }
                });
#pragma unroll
                for (int j = 0; j < N; ++j) WO(D::W_DX, j) = -WO(D::W_C, j);
                sc = 1.0;
            } else {
                const bool comb = phase == 2;          // ds = -lam o lam (:120) [+ sigma mu e - kt2 o kt3 (:137-139)]
                // dz and the corrector term live in the L2 workspace: all their loads go out before the first cone
                double dzr[K], dscr[K];
#pragma unroll
                for (int r = 0; r < K; ++r) { dzr[r] = WO(D::W_DZ, r); dscr[r] = WO(D::W_DSC, r); }
#pragma unroll
                for (int i = 0; i < KPOC; ++i) {
                    const double w = SV(D::V_WB, i), iw = SO(D::O_IWB, i), lv = SV(D::V_LAM, i);
                    const double dsc = dscr[i] + smu_l;
                    const double dsv = -(lv * lv) + (comb ? dsc : 0.0);
                    const double kk = dsv * fast_rcp(lv);
                    const double kz = sc * dzr[i] - w * kk;
                    SV(D::V_K0, i) = kk; SV(D::V_K2, i) = kz; SV(D::V_U, i) = iw * iw * kz;
                }
                D::for_groups([&](auto gtag) {
                    constexpr int SDIM = decltype(gtag)::DIM, GCNT = decltype(gtag)::COUNT, GOFF = decltype(gtag)::OFF, GC0 = decltype(gtag)::C0;
#pragma unroll
                    for (int cc = 0; cc < GCNT; ++cc) {
                    const int c = GC0 + cc, o = GOFF + cc * SDIM;
                    const double eta = SO(D::O_CS, 4 * c + 0), ie = SO(D::O_CS, 4 * c + 1), r1w = SO(D::O_CS, 4 * c + 2),
                                 llt = SO(D::O_CS, 4 * c + 3);
                    const double ie2 = ie * ie;
                    double lv[SDIM], wv[SDIM], dsv[SDIM], k0v[SDIM], k2v[SDIM];
#pragma unroll
                    for (int e = 0; e < SDIM; ++e) { lv[e] = SV(D::V_LAM, o + e); wv[e] = SV(D::V_WB, o + e); }
                    const double l0 = lv[0], w0 = wv[0], aa = l0 * l0 - llt;
                    {
                        const double d0 = dscr[o] + smu_l;
                        dsv[0] = -(llt + l0 * l0) + (comb ? d0 : 0.0);                  // src/vectors.jl:66-69
                    }
                    double beta = 0.0;
#pragma unroll
                    for (int e = 1; e < SDIM; ++e) {
                        const double de = dscr[o + e];
                        dsv[e] = -(l0 * lv[e] + l0 * lv[e]) + (comb ? de : 0.0);        // :73-75
                        beta = fma(lv[e], dsv[e], beta);
                    }
                    const double ia = fast_rcp(aa), il0 = fast_rcp(l0);
                    k0v[0] = (l0 * dsv[0] - beta) * ia;                                 // src/vectors.jl:105-125, O(d) form
                    double dl = 0.0;
#pragma unroll
                    for (int e = 1; e < SDIM; ++e) {
                        k0v[e] = (-dsv[0] * lv[e] + (aa * dsv[e] + beta * lv[e]) * il0) * ia;
                        dl = fma(wv[e], k0v[e], dl);
                    }
                    const double cst = k0v[0] + dl * r1w;                               // src/scalings.jl:135
                    k2v[0] = dzr[o] * sc - eta * (w0 * k0v[0] + dl);            // :136, densesolver :65
                    double qv = w0 * k2v[0];
#pragma unroll
                    for (int e = 1; e < SDIM; ++e) {
                        k2v[e] = dzr[o + e] * sc - eta * (k0v[e] + cst * wv[e]);        // :137-139
                        qv = fma(-wv[e], k2v[e], qv);                                   // W^-2 = eta^-2 (2 q q' - J)
                    }
#pragma unroll
                    for (int e = 1; e < SDIM; ++e) {
                        SV(D::V_K0, o + e) = k0v[e];
                        SV(D::V_K2, o + e) = k2v[e];
                        SV(D::V_U, o + e) = ie2 * (k2v[e] - 2.0 * wv[e] * qv);
                    }
                    SV(D::V_K0, o) = k0v[0];
                    SV(D::V_K2, o) = k2v[0];
                    SV(D::V_U, o) = ie2 * (2.0 * w0 * qv - k2v[0]);
                }
                });
            }
            // ------------------------------------------------ n0 = G'u + sc dx                src/densesolver.jl:66-67
            double n0[N];
#pragma unroll
            for (int j = 0; j < N; ++j) n0[j] = 0.0;
            if (!fslot) {          // (in an F slot the SYRK pass below accumulates G'u as well: one pass over G less)
                fl_stream_rows<K, NP, LPW, RS, NS>(G2, ring, [&](int r, const double (&g)[NP]) {
                    const double ur = SV(D::V_U, r);
#pragma unroll
                    for (int j = 0; j < N; ++j) n0[j] = fma(g[j], ur, n0[j]);
                });
            }

            double Lr[NH], dxv[N];
            bool ok = true;
            if (!fslot) {
#pragma unroll
                for (int j = 0; j < N; ++j) dxv[j] = WO(D::W_DX, j);
            }
            if (fslot) {
                // -------------------------------------------- KKT factor, src/densesolver.jl:41-47
                const bool init = phase == 0;
#pragma unroll
                for (int e = 0; e < NH; ++e) Lr[e] = 0.0;
                // rows [0, KPOC): d = iwb^2; cone rows: head -eta^-2, tail eta^-2, and after the cone's last row the
                // rank-one term h_c h_c'
                double hq[N];
#pragma unroll
                for (int j = 0; j < N; ++j) hq[j] = 0.0;
                fl_stream_rows<K, NP, LPW, RS, NS>(G2, ring, [&](int r, const double (&g)[NP]) {
                    double d, wq = 0.0, f = 0.0;
                    int e = -1;
                    bool last = false;
                    if (r < KPOC) {
                        const double iw = SO(D::O_IWB, r);
                        d = iw * iw;
                    } else {
                        int c = 0;
                        D::for_groups([&](auto gtag) {
                            constexpr int SDIM = decltype(gtag)::DIM, GCNT = decltype(gtag)::COUNT, GOFF = decltype(gtag)::OFF, GC0 = decltype(gtag)::C0;
                            if (r >= GOFF && r < GOFF + GCNT * SDIM) {
                                const int cc = (r - GOFF) / SDIM;
                                e = (r - GOFF) - cc * SDIM;
                                c = GC0 + cc;
                                last = e == SDIM - 1;
                            }
                        });
                        const double ie = SO(D::O_CS, 4 * c + 1);
                        const double ie2 = ie * ie;
                        const double wb = SV(D::V_WB, r);
                        wq = e == 0 ? wb : -wb;
                        d = (e == 0 && !init) ? -ie2 : ie2;                       // -eta^-2 on the head, eta^-2 on the tail
                        f = init ? 0.0 : 1.4142135623730951 * ie;                 // h_c = sqrt(2)/eta G_c'q, q = J wbar
                    }
                    const double ur = SV(D::V_U, r);
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        n0[j] = fma(g[j], ur, n0[j]);
                        hq[j] = fma(wq, g[j], hq[j]);
                        const double t = d * g[j];
#pragma unroll
                        for (int q = j; q < N; ++q) Lr[TRI(q, j)] = fma(g[q], t, Lr[TRI(q, j)]);
                    }
                    if (last) {
#pragma unroll
                        for (int j = 0; j < N; ++j) hq[j] *= f;
#pragma unroll
                        for (int j = 0; j < N; ++j)
#pragma unroll
                            for (int q = j; q < N; ++q) Lr[TRI(q, j)] = fma(hq[q], hq[j], Lr[TRI(q, j)]);
#pragma unroll
                        for (int j = 0; j < N; ++j) hq[j] = 0.0;
                    }
                });
#pragma unroll
                for (int j = 0; j < N; ++j) dxv[j] = WO(D::W_DX, j);        // in flight under the factorisation below
                // in-register LL' (src/densesolver.jl:47); the diagonal keeps 1 / l_jj
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    const double d = Lr[TRI(j, j)];
                    ok = ok && (d > 0.0);
                    const double rj = fast_rsqrt(d);
                    Lr[TRI(j, j)] = rj;
#pragma unroll
                    for (int i = j + 1; i < N; ++i) Lr[TRI(i, j)] *= rj;
#pragma unroll
                    for (int i = j + 1; i < N; ++i)
#pragma unroll
                        for (int q = j + 1; q <= i; ++q) Lr[TRI(i, q)] = fma(-Lr[TRI(i, j)], Lr[TRI(q, j)], Lr[TRI(i, q)]);
                }
                if (!init) {
#pragma unroll
                    for (int e = 0; e < NH; ++e) WO(D::W_L, e) = Lr[e];
                }
            } else {
#pragma unroll
                for (int e = 0; e < NH; ++e) Lr[e] = WO(D::W_L, e);
            }
            if (!ok) { status = ST_NUMERICAL; dead = phase == 0; phase = FL_DONE; }     // cholesky! threw
            else {
#pragma unroll
                for (int j = 0; j < N; ++j) n0[j] = fma(sc, dxv[j], n0[j]);
                // -------------------------------------------- cx = H^-1 n0 by substitution     src/densesolver.jl:83
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    n0[j] *= Lr[TRI(j, j)];
#pragma unroll
                    for (int i = j + 1; i < N; ++i) n0[i] = fma(-Lr[TRI(i, j)], n0[j], n0[i]);
                }
#pragma unroll
                for (int j = N - 1; j >= 0; --j) {
                    n0[j] *= Lr[TRI(j, j)];
#pragma unroll
                    for (int m = 0; m < j; ++m) n0[m] = fma(-Lr[TRI(j, m)], n0[j], n0[m]);
                }
                // -------------------------------------------- u = G cx - k2                     :84-85
                fl_stream_rows<K, NP, LPW, RS, NS>(G2, ring, [&](int r, const double (&g)[NP]) {
                    double a0 = 0.0, a1 = 0.0;
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        if (j & 1) a1 = fma(g[j], n0[j], a1); else a0 = fma(g[j], n0[j], a0);
                    }
                    SV(D::V_U, r) = (a0 + a1) - SV(D::V_K2, r);
                });

                if (phase == 0) {
                    // ---------------------------------------- initial iterate, src/solver.jl:86-101 (max_step: src/mats.jl:1-28)
#pragma unroll
                    for (int j = 0; j < N; ++j) WO(D::W_X, j) = n0[j];
                    double mp = -INFINITY, md = -INFINITY;
                    for (int i = 0; i < KPOC; ++i) { const double v = SV(D::V_U, i); mp = fmax(mp, v); md = fmax(md, -v); }
                    D::for_groups([&](auto gtag) {
                        constexpr int SDIM = decltype(gtag)::DIM, GCNT = decltype(gtag)::COUNT, GOFF = decltype(gtag)::OFF, GC0 = decltype(gtag)::C0;
                        for (int cc = 0; cc < GCNT; ++cc) {
                        const int c = GC0 + cc, o = GOFF + cc * SDIM;
                        double sq = 0.0;
#pragma unroll
                        for (int e = 1; e < SDIM; ++e) { const double v = SV(D::V_U, o + e); sq = fma(v, v, sq); }
                        const double nr = fast_sqrt(sq), z0 = SV(D::V_U, o);
                        mp = fmax(mp, nr + z0);
                        md = fmax(md, nr - z0);
                    }
                    });
                    const bool shp = !(fabs(mp) < prm.init_eps), shd = !(fabs(md) < prm.init_eps);
                    const double shs = shp ? 1.0 + mp : 0.0, shz = shd ? 1.0 + md : 0.0;
                    for (int i = 0; i < KPOC; ++i) {
                        const double z0 = SV(D::V_U, i);
                        WO(D::W_S, i) = shp ? -z0 + shs : -z0;
                        WO(D::W_Z, i) = shd ? z0 + shz : z0;
                    }
                    D::for_groups([&](auto gtag) {
                        constexpr int SDIM = decltype(gtag)::DIM, GCNT = decltype(gtag)::COUNT, GOFF = decltype(gtag)::OFF, GC0 = decltype(gtag)::C0;
                        for (int cc = 0; cc < GCNT; ++cc) {
                        const int c = GC0 + cc, o = GOFF + cc * SDIM;
#pragma unroll
                        for (int e = 0; e < SDIM; ++e) {
                            const double z0 = SV(D::V_U, o + e);
                            WO(D::W_S, o + e) = (e == 0 && shp) ? -z0 + shs : -z0;
                            WO(D::W_Z, o + e) = (e == 0 && shd) ? z0 + shz : z0;
                        }
                    }
                    });
                    need_top = true;
                    phase = 1;
                } else {
                    // ---------------------------------------- tail of solve_kkt (src/densesolver.jl:86-89), scale!/iscale!
                    // (src/solver.jl:128-129), scmax of both results (src/mats.jl:53-86); u <- cz, k0 <- cs, dsc <- -(kt2 o kt3)
                    const bool chk = phase == 2;
                    double dotacc = 0.0, mx = -INFINITY;
                    int fl = 0;
#pragma unroll 2
                    for (int i = 0; i < KPOC; ++i) {
                        const double w = SV(D::V_WB, i), iw = SO(D::O_IWB, i), il = fast_rcp(SV(D::V_LAM, i));
                        const double cz = iw * iw * SV(D::V_U, i);
                        const double kt3 = w * cz;
                        const double kk = SV(D::V_K0, i) - kt3;
                        const double csx = w * kk;
                        const double kt2 = iw * csx;
                        mx = fmax(mx, fmax(-kt3 * il, -kt2 * il));
                        dotacc = fma(kt2, kt3, dotacc);
                        SV(D::V_U, i) = cz;
                        SV(D::V_K0, i) = csx;
                        WO(D::W_DSC, i) = -(kt2 * kt3);
                        if (chk) fl |= !isfinite(cz) | !isfinite(csx);
                    }
                    D::for_groups([&](auto gtag) {
                        constexpr int SDIM = decltype(gtag)::DIM, GCNT = decltype(gtag)::COUNT, GOFF = decltype(gtag)::OFF, GC0 = decltype(gtag)::C0;
#pragma unroll CU
                        for (int cc = 0; cc < GCNT; ++cc) {
                        const int c = GC0 + cc, o = GOFF + cc * SDIM;
                        const double eta = SO(D::O_CS, 4 * c + 0), ie = SO(D::O_CS, 4 * c + 1), r1w = SO(D::O_CS, 4 * c + 2),
                                     llt = SO(D::O_CS, 4 * c + 3);
                        const double ie2 = ie * ie;
                        double wv[SDIM], lv[SDIM], uv[SDIM], k0v[SDIM], czv[SDIM], csv[SDIM], kt2v[SDIM], kt3v[SDIM];
#pragma unroll
                        for (int e = 0; e < SDIM; ++e) {
                            wv[e] = SV(D::V_WB, o + e); lv[e] = SV(D::V_LAM, o + e);
                            uv[e] = SV(D::V_U, o + e); k0v[e] = SV(D::V_K0, o + e);
                        }
                        const double w0 = wv[0], l0 = lv[0], aa = l0 * l0 - llt;
                        double qv = w0 * uv[0];                                         // cz = W^-2 u          :86
#pragma unroll
                        for (int e = 1; e < SDIM; ++e) qv = fma(-wv[e], uv[e], qv);
                        czv[0] = ie2 * (2.0 * w0 * qv - uv[0]);
                        double dl = 0.0;
#pragma unroll
                        for (int e = 1; e < SDIM; ++e) { czv[e] = ie2 * (uv[e] - 2.0 * wv[e] * qv); dl = fma(wv[e], czv[e], dl); }
                        double cst = czv[0] + dl * r1w;                                 // kt3 = W cz           :87, solver :128
                        kt3v[0] = eta * (w0 * czv[0] + dl);
                        k0v[0] -= kt3v[0];                                              // k0 -= W cz           :88
                        dl = 0.0;
#pragma unroll
                        for (int e = 1; e < SDIM; ++e) {
                            kt3v[e] = eta * (czv[e] + cst * wv[e]);
                            k0v[e] -= kt3v[e];
                            dl = fma(wv[e], k0v[e], dl);
                        }
                        cst = k0v[0] + dl * r1w;                                        // cs = W k0            :89
                        csv[0] = eta * (w0 * k0v[0] + dl);
                        dl = 0.0;
#pragma unroll
                        for (int e = 1; e < SDIM; ++e) { csv[e] = eta * (k0v[e] + cst * wv[e]); dl = fma(wv[e], csv[e], dl); }
                        cst = -csv[0] + dl * r1w;                                       // kt2 = W^-1 cs        solver :129
                        kt2v[0] = ie * (w0 * csv[0] - dl);
                        double lx3 = 0.0, lx2 = 0.0, dot = 0.0;
#pragma unroll
                        for (int e = 1; e < SDIM; ++e) {
                            kt2v[e] = ie * (csv[e] + cst * wv[e]);
                            lx3 = fma(lv[e], kt3v[e], lx3);
                            lx2 = fma(lv[e], kt2v[e], lx2);
                            dot = fma(kt2v[e], kt3v[e], dot);
                        }
                        dot += kt2v[0] * kt3v[0];
                        fl |= !(aa >= 0.0);
                        const double as = fast_rsqrt(aa);                               // src/mats.jl:67-71
                        const double r13 = as * l0 * kt3v[0] - as * lx3, r12 = as * l0 * kt2v[0] - as * lx2;   // :74-77
                        const double den = fast_rcp(as * l0 + 1.0);
                        const double c3 = (r13 + kt3v[0]) * den, c2 = (r12 + kt2v[0]) * den;      // :80
                        double q3 = 0.0, q2 = 0.0;
#pragma unroll
                        for (int e = 1; e < SDIM; ++e) {
                            const double v3 = as * (kt3v[e] - c3 * as * lv[e]);         // :83
                            const double v2 = as * (kt2v[e] - c2 * as * lv[e]);
                            q3 = fma(v3, v3, q3);
                            q2 = fma(v2, v2, q2);
                        }
                        mx = fmax(mx, fmax(fast_sqrt(q3) - as * r13, fast_sqrt(q2) - as * r12));  // :85
                        dotacc += dot;
#pragma unroll
                        for (int e = 1; e < SDIM; ++e) {
                            SV(D::V_U, o + e) = czv[e];
                            SV(D::V_K0, o + e) = csv[e];
                            WO(D::W_DSC, o + e) = -(kt2v[0] * kt3v[e] + kt3v[0] * kt2v[e]);      // src/vectors.jl:73-75
                            if (chk) fl |= !isfinite(czv[e]) | !isfinite(csv[e]);
                        }
                        SV(D::V_U, o) = czv[0];
                        SV(D::V_K0, o) = csv[0];
                        WO(D::W_DSC, o) = -dot;                                         // src/vectors.jl:66-69
                        if (chk) fl |= !isfinite(czv[0]) | !isfinite(csv[0]);
                    }
                    });
                    const double tstep = step_from_t(mx);                               // src/solver.jl:130 / :145
                    if (phase == 1) {
                        // centering parameter (:130-134) and the combined right-hand side (:136-140)
                        const double rho = 1.0 - tstep - tstep * tstep * dotacc * fast_rcp(ll);   // :132 (minus: reference quirk)
                        const double cl = fmax(0.0, fmin(1.0, rho));
                        const double sig = cl * cl * cl;                                // :133
                        const double mu = ll / (double)a.deg;                           // :134
                        if (fl) { status = ST_NUMERICAL; phase = FL_DONE; }
                        else {
                            const double smu = sig * mu;
                            sc = 1.0 - sig;                                             // :136
                            smu_l = smu;                // :137-139: added to the orthant rows and the cone heads of
                                                        // the corrector term when the combined head reads it
                            phase = 2;
                        }
                    } else {
                        // step length (:143-146), iterate update (:147-150)
                        const double step = tstep * prm.step_damp;
#pragma unroll
                        for (int j = 0; j < N; ++j) fl |= !isfinite(n0[j]);
                        fl |= !isfinite(step);
                        if (fl) { status = ST_NUMERICAL; phase = FL_DONE; }
                        else {
#pragma unroll
                            for (int j = 0; j < N; ++j) WO(D::W_X, j) = fma(n0[j], step, WO(D::W_X, j));      // :147
                            constexpr int UB = 20;
#pragma unroll 1
                            for (int r0 = 0; r0 < K; r0 += UB) {       // workspace loads of a batch before its stores
                                double zv[UB], sv[UB];
#pragma unroll
                                for (int q = 0; q < UB; ++q) {
                                    zv[q] = r0 + q < K ? WO(D::W_Z, r0 + q) : 0.0;
                                    sv[q] = r0 + q < K ? WO(D::W_S, r0 + q) : 0.0;
                                }
#pragma unroll
                                for (int q = 0; q < UB; ++q) {
                                    const int r = r0 + q;
                                    if (r < K) {
                                        WO(D::W_Z, r) = fma(SV(D::V_U, r), step, zv[q]);                     // :149
                                        WO(D::W_S, r) = fma(SV(D::V_K0, r), step, sv[q]);                    // :150
                                    }
                                }
                            }
                            ++iters;
                            need_top = true;
                            phase = 1;
                        }
                    }
                }
            }
        }
        fslot = !fslot;
    }
#undef SV
#undef SO
#undef WO
#undef TRI
}

}  // namespace socp
