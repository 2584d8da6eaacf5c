// fused_v2.cuh -- whole-solve kernel for layouts that fit in shared memory.
//
// One CTA ("team" of NW warps; a single warp for n <= 16) owns one problem from the initial point
// to the last Mehrotra step (reference src/solver.jl:68-152).  G, the reduced KKT matrix H (later
// its explicit inverse), the inverse X = L^-1 of its Cholesky factor and every work vector live in
// shared memory; global traffic is the problem in and the iterate out.  CTAs are persistent and
// pull problems from an atomic counter.
//
// What bounds this kernel is the serial dependency chain of one interior-point iteration, so the
// design minimises team barriers, reduction rounds and instructions on the critical warp:
//   * cone chains (compute_scaling; solve_kkt head: iprod -> W -> W^-2; tail: W^-2 -> W -> W ->
//     W^-1 -> scmax) run register-resident on a group of LPC lanes per second-order cone (LPC = 1
//     for SOC(4): one thread per cone, no shuffles; LPC = 16 for SOC(51)), reductions are xor
//     shuffles inside the group; positive-orthant rows are elementwise over the whole team
//   * H = G'W^-2 G = G' diag(dw) G + sum_c hc hc' straight from G (the reference's product order,
//     src/densesolver.jl:42-43, with the block-diagonal iWiW in closed form): mma.sync m8n8k4 f64
//     (SASS DMMA.8x8x4) with the row weight folded into the A fragment; no scaled copy of G
//   * blocked right-looking Cholesky on 8x8 tiles that carries the inverse along: one warp factors
//     the diagonal tile redundantly in registers (no shuffles, no barriers), the panel and the
//     trailing update (of H and of the partially built inverse) are DMMA tile products driven by
//     a descriptor table; two team barriers per block column, and the diagonal warp only arrives
//     at the second one
//   * explicit H^-1 = X'X by DMMA, as the reference does (Li, src/densesolver.jl:48,83); equality
//     rows are folded into one solve matrix per factorisation (Pm = H^-1 - K (H^-1 A')', K = H^-1 A'
//     M^-1), so a solve is three phases: n0 = G'u + dx, (cx, cy) = (Pm n0 + K dy, K'n0 - M^-1 dy),
//     u = G cx - k2
//   * one copy of factor / solve / head / tail for the initial point and both directions (state
//     machine), compile-time specialised layouts for the BASELINE.json shapes (DimsStatic)
//   * stop-test norms, sigma/mu and step-length reductions share one scratch exchange per phase
//
// Restrictions (the tiled path takes everything else): no `sing` problems, n <= 64, p <= 32,
// second-order cones of dimension <= 128, at most 64 of them.
#pragma once
#include "fused_common.cuh"
#ifndef __CUDACC_RTC__          // (NVRTC compiles the device code of this file alone: run-time specialisation, lane_jit.cu)
#include <vector>
#include <algorithm>
#include <cstdlib>
#endif

namespace socp {


struct F2Plan {
    bool fits = false;
    int variant = 0;       // 0: 1 warp (n<=16), 1: 4 warps (n<=32), 2: 8 warps (n<=56), 3: 8 warps (n<=64)
    int nw = 1;
    int ctas_per_sm = 1, num_sms = 148;
    size_t smem = 0;
    int* d_counter = nullptr;
    unsigned long long* d_clk = nullptr;   // 16 phase-cycle counters (SOCP_PHASE_TIMING builds)
    void* jit_fn = nullptr;                // the kernel specialised for this layout at run time (lane_jit.cu), if any
    int jit_tried = 0, device = 0;
    // layout
    int n = 0, p = 0, k = 0, kpoc = 0, nsoc = 0, lpc = 1;
    int npad = 0, nb = 0, kpad = 0, ldg = 0, ldh = 0, ppad = 0, pb = 0, ldm = 0;
    int shape = 0;         // 0: generic kernel; > 0: index of a compile-time specialised layout (F2_SHAPES)
    int split_k = 1, split_n = 1, split_p = 1;    // lanes per output row of the row gemvs (k, n, p rows)
    int soc_offs[F2_MAX_SOC], soc_dim[F2_MAX_SOC];
    // offsets into the dynamic shared memory, in doubles
    int oG, oR, oX, oA, oAt, oHiAt, oK, oM, oMX, oMinv, oDinv;
    int oc, ob, oh, ox, oy, oz, os, olam, owb, oiwb, ocs, odx, ody, odz, ods, ok0, ok2, ou;
    int on0, ocx, ocy, okd, omd, odw, ohc, oscr, odesc, odescm;
    int total = 0;
};

#ifndef __CUDACC_RTC__
// kind/offs/dim: the CALLER's cones (POC blocks first).
inline void f2_plan(F2Plan& P, int n, int p, int k, const std::vector<int>& kind, const std::vector<int>& offs,
                    const std::vector<int>& dim, int device) {
    P.fits = false;
    P.device = device;
    P.n = n; P.p = p; P.k = k;
    P.kpoc = 0; P.nsoc = 0;
    int maxd = 1;
    for (size_t i = 0; i < kind.size(); ++i) {
        if (kind[i] == KIND_POC) P.kpoc += dim[i];
        else {
            if (P.nsoc >= F2_MAX_SOC) return;
            P.soc_offs[P.nsoc] = offs[i];
            P.soc_dim[P.nsoc] = dim[i];
            maxd = std::max(maxd, dim[i]);
            ++P.nsoc;
        }
    }
    if (maxd > 128 || n > 64 || p > 32) return;
    P.lpc = f2_lpc(maxd);
    if (n <= 16 && p <= 16) { P.variant = 0; P.nw = 1; }
    else if (n <= 32) { P.variant = 1; P.nw = 4; }
    else if (n <= 56) { P.variant = 2; P.nw = 8; }
    else { P.variant = 3; P.nw = 8; }
    P.npad = (n + 7) / 8 * 8; P.nb = P.npad / 8;
    P.kpad = (k + 3) / 4 * 4;
    P.ldg = f2_ld(P.kpad);
    P.ldh = f2_ld(P.npad);
    P.ppad = (std::max(p, 1) + 7) / 8 * 8; P.pb = P.ppad / 8;
    P.ldm = f2_ld(P.ppad);
    P.split_k = f2_split((k + 1) / 2, P.nw);      // row PAIRS per lane in the k-row gemv
    P.split_n = f2_split(n, P.nw);
    P.split_p = f2_split(std::max(p, 1), P.nw);
    int at = 0;
    auto take = [&](int cnt) { int r = at; at += (cnt + 1) / 2 * 2; return r; };
    P.oG = take(P.ldg * (n + 1));                 // pad rows and column n stay zero (pad columns of the SYRK read column n)
    P.oR = take(P.ldh * P.npad + f2_xsize(P.nb)); // H (later H^-1, then the solve matrix) | X (trapezoid)
    P.oX = P.oR + P.ldh * P.npad;
    P.oA = take(p * n); P.oAt = take(p * P.npad); P.oHiAt = take(n * p); P.oK = take(n * p);
    P.oM = take(p ? P.ldm * P.ppad : 0); P.oMX = take(p ? f2_xsize(P.pb) : 0); P.oMinv = take(p * p);
    P.oDinv = take((P.nw > 1 ? 2 : 1) * 8 * 12);
    P.oc = take(n); P.ox = take(n); P.odx = take(n); P.on0 = take(P.npad); P.ocx = take(n); P.okd = take(n);
    P.ob = take(p); P.oy = take(p); P.ody = take(p); P.ocy = take(p); P.omd = take(p);
    P.odw = take(P.kpad); P.ohc = take(std::max(P.nsoc, 1) * P.npad);
    P.oh = take(k); P.oz = take(k); P.os = take(k); P.olam = take(k); P.owb = take(k); P.oiwb = take(P.kpoc);
    P.odz = take(k); P.ods = take(k); P.ok0 = take(k); P.ok2 = take(k); P.ou = take(k);
    P.ocs = take(F2_CS * std::max(P.nsoc, 1));
    P.oscr = take(P.nw > 1 ? 2 * 8 * 8 : 0);     // team_reduce scratch: a one-warp team reduces in registers
    P.odesc = take(f2_trail_base(P.nb, P.nb)); P.odescm = take(p ? f2_trail_base(P.pb, P.pb) : 0);   // uint2 = 8 bytes each
    P.total = at;
    P.smem = (size_t)at * sizeof(double);
    int dev_smem = 0, sms = 148;
    if (cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    P.num_sms = sms;
    if (P.smem + 512 > (size_t)dev_smem) return;
    const int per_sm = 228 * 1024;
    const int threads = P.nw * 32;
    const int reg_cap = P.variant == 0 ? 16 : (P.variant == 1 ? 4 : 2);     // matches the __launch_bounds__ below
    P.ctas_per_sm = std::max(1, std::min({(int)(per_sm / (P.smem + 1024 + 256)), 2048 / threads, 32, reg_cap}));
    P.fits = true;
}

#endif  // __CUDACC_RTC__

// ------------------------------------------------------------------------------------------------ SYRK
// Tile t (row-major over the lower triangle) -> (ti, tj), ti >= tj.
__device__ __forceinline__ void f2_tile(int t, int& ti, int& tj) {
    int a = 0;
    while ((a + 1) * (a + 2) / 2 <= t) ++a;
    ti = a;
    tj = t - a * (a + 1) / 2;
}
// H (lower 8x8 tiles incl. full diagonal tiles; n x n padded to npad with a unit pad diagonal; ld ldh)
//   = G' diag(dw) G + sum_c hc[c] hc[c]'
// which is G'W^-2 G of src/densesolver.jl:42-43 for W^-2 = blockdiag(iwb^2 | eta^-2 (2 q q' - J)) (SURVEY.md appendix
// A.1): dw = iwb^2 on positive-orthant rows, eta^-2 on cone tails, -eta^-2 on cone heads, hc[c] = sqrt(2)/eta G_c'q.
// G: kpad x (n+1), ld ldg (pad rows and column n zero).  mma.sync m8n8k4 f64 from shared memory; accumulators in
// registers.  The caller syncs afterwards.
template <int NW, int MAXT>
__device__ __forceinline__ void f2_my_tiles(int nb, int warp, int (&tl)[MAXT]) {
    const int ntl = nb * (nb + 1) / 2;
#pragma unroll
    for (int q = 0; q < MAXT; ++q) {
        int ti, tj;
        f2_tile(min(warp + q * NW, ntl - 1), ti, tj);
        tl[q] = ti | (tj << 8);
    }
}
template <int NW, int MAXT>
__device__ __forceinline__ void f2_syrk_w(const double* G, int ldg, int kpad, int n, int nb, const double* dw,
                                          const double* hc, int nsoc, int npad, double* H, int ldh, const int (&tl)[MAXT],
                                          int lane, int warp) {
    const int ntl = nb * (nb + 1) / 2;
    const int fr = lane >> 2, fk = lane & 3;
    const double* pa[MAXT];
    const double* pb[MAXT];
    int ti[MAXT], tj[MAXT];
    double acc[MAXT][2];
#pragma unroll
    for (int q = 0; q < MAXT; ++q) {
        ti[q] = tl[q] & 255; tj[q] = tl[q] >> 8;
        pa[q] = G + min(ti[q] * 8 + fr, n) * ldg + fk;      // column n is the zero column
        pb[q] = G + min(tj[q] * 8 + fr, n) * ldg + fk;
        acc[q][0] = acc[q][1] = 0.0;
    }
    const double* pw = dw + fk;
#pragma unroll 4
    for (int kk = 0; kk < kpad; kk += 4) {
        const double w = pw[kk];
#pragma unroll
        for (int q = 0; q < MAXT; ++q)
            if (warp + q * NW < ntl) dmma884(acc[q][0], acc[q][1], pa[q][kk] * w, pb[q][kk]);
    }
#pragma unroll
    for (int q = 0; q < MAXT; ++q) {
        if (warp + q * NW < ntl) {
            const int gi = ti[q] * 8 + fr;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gj = tj[q] * 8 + 2 * fk + e;
                double v = acc[q][e];
                for (int c = 0; c < nsoc; ++c) v = fma(hc[c * npad + gi], hc[c * npad + gj], v);
                if (gi == gj && gi >= n) v = 1.0;
                H[gj * ldh + gi] = v;
            }
        }
    }
}
// Out (full symmetric, ld) = X' X for the lower-triangular X (trapezoid storage; the part of a diagonal tile above
// the diagonal is zero): H^-1 = L^-T L^-1.  Out must not alias X nor anything another warp still reads.
template <int NW, int MAXT>
__device__ __forceinline__ void f2_xtx(const double* X, int ld, int nb, double* Out, const int (&tl)[MAXT], int lane,
                                       int warp) {
    const int ntl = nb * (nb + 1) / 2;
    const int fr = lane >> 2, fk = lane & 3;
    int ti[MAXT], tj[MAXT];
    double acc[MAXT][2];
#pragma unroll
    for (int q = 0; q < MAXT; ++q) {
        ti[q] = tl[q] & 255; tj[q] = tl[q] >> 8;
        acc[q][0] = acc[q][1] = 0.0;
        if (warp + q * NW < ntl) {
            // rows kk of column block cb live at offset kk - 8*cb of that block
            const double* pa = X + f2_xbase(nb, ti[q]) + fr * f2_xld(nb, ti[q]) + fk - ti[q] * 8;
            const double* pb = X + f2_xbase(nb, tj[q]) + fr * f2_xld(nb, tj[q]) + fk - tj[q] * 8;
            for (int kk = ti[q] * 8; kk < nb * 8; kk += 4) dmma884(acc[q][0], acc[q][1], pa[kk], pb[kk]);
        }
    }
#pragma unroll
    for (int q = 0; q < MAXT; ++q) {
        if (warp + q * NW < ntl) {
            const int gi = ti[q] * 8 + fr;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gj = tj[q] * 8 + 2 * fk + e;
                Out[gj * ld + gi] = acc[q][e];
                Out[gi * ld + gj] = acc[q][e];
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------ blocked Cholesky + inverse
// Trailing-update tile descriptors of the blocked factorisation: for block column b, tile (i, j), i > b, j <= i
// (flat index over rows i = b+1.. with i+1 tiles each), packed offsets (in doubles, relative to H) of the three
// operands -- the same table serves every factorisation of the kernel's lifetime.
//   x: aoff | boff << 16      a = H[aoff + (kk+fk)*ld + fr]        (L_ib, negated)
//   y: coff | cl << 16 | flags << 24
//        flags bit 0: b operand in X form  b = H[boff + fr*cl + kk + fk]   (else H form, like a)
//        flags bit 1: destination starts from zero (X(i,b) = -L_ib X(b,b))
//      C[fr][2fk+e] = H[coff + (2fk+e)*cl + fr]
__device__ __forceinline__ uint2 f2_trail_desc(int nbl, int ld, int xoff, int b, int tf) {
    int i = b + 1, j = tf;
    while (j > i) { j -= i + 1; ++i; }
    const int b0 = b * 8, i0 = i * 8, j0 = j * 8;
    const int aoff = b0 * ld + i0;
    int boff, coff, cl, flags;
    if (j > b) {
        boff = b0 * ld + j0; coff = j0 * ld + i0; cl = ld; flags = 0;
    } else {
        const int xj = xoff + f2_xbase(nbl, j);
        cl = f2_xld(nbl, j);
        boff = xj + (b0 - j0); coff = xj + (i0 - j0); flags = 1 | (j == b ? 2 : 0);
    }
    return make_uint2((unsigned)aoff | ((unsigned)boff << 16), (unsigned)coff | ((unsigned)cl << 16) | ((unsigned)flags << 24));
}
// fills desc[0 .. f2_trail_base(nbl, nbl)) -- call once, by the whole team, before the first factorisation
__device__ __forceinline__ void f2_build_trail(uint2* desc, int nbl, int ld, int xoff, int tid, int nthreads) {
    for (int b = 0; b < nbl; ++b) {
        const int nt = f2_ntrail(nbl, b), base = f2_trail_base(nbl, b);
        for (int t = tid; t < nt; t += nthreads) desc[base + t] = f2_trail_desc(nbl, ld, xoff, b, t);
    }
}

// H: nbl x nbl lower tiles (column-major, ld), unit pad diagonal.  On exit X (trapezoid storage, zero-initialised by
// the caller) holds L^-1; H is destroyed.  Dinv: 2 x (8 x 12) scratch, zero above the diagonal.  desc: the table of
// f2_build_trail for (nbl, ld, X - H).  *fail is set (and 0 returned) on a non-positive pivot: cholesky!'s
// PosDefException, src/densesolver.jl:47,51.
template <int NW, bool WRITE_L = false>
__device__ __forceinline__ int f2_chol_inv(double* H, double* X, double* Dinv, const uint2* desc, int nbl, int ld,
                                           int* fail, int lane, int warp, unsigned long long* clk = nullptr) {
    const int fr = lane >> 2, fk = lane & 3;
    const int la = fk * ld + fr;
#ifdef SOCP_PHASE_TIMING
    const bool pt_on = clk && (NW == 8) && !WRITE_L && threadIdx.x == 0 && blockIdx.x == 0;
    long long pt_c = clock64();
#endif
    for (int b = 0; b < nbl; ++b) {
        const int b0 = b * 8;
        double* Db = Dinv + ((NW > 1) ? (b & 1) * 96 : 0);      // double-buffered only when warp 0 runs ahead
#ifdef SOCP_PHASE_TIMING
        if (pt_on) pt_c = clock64();
#endif
        if (warp == 0) {
            __syncwarp();
            double xc[8];
            const int ok = f2_diag_factor<WRITE_L>(H + b0 * ld + b0, ld, lane, xc);
            if (!ok && lane == 0) *fail = 1;
            // publish column c = lane of the tile's inverse: Dinv (row-major, ld 12) and the diagonal tile of X
            if (lane < 8) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    Db[i * 12 + lane] = xc[i];
                    X[f2_xbase(nbl, b) + lane * f2_xld(nbl, b) + i] = xc[i];
                }
            }
        }
#ifdef SOCP_PHASE_TIMING
        if (pt_on) { const long long t = clock64(); atomicAdd(&clk[13], (unsigned long long)(t - pt_c)); pt_c = t; }
#endif
        tsync<NW>();                                   // (1) Dinv_b visible; trailing update of step b-1 complete
#ifdef SOCP_PHASE_TIMING
        if (pt_on) { const long long t = clock64(); atomicAdd(&clk[15], (unsigned long long)(t - pt_c)); pt_c = t; }
#endif
        if (*fail) return 0;
        // ---- panel: H(i,b) <- H(i,b) Dinv_b'  (i > b);   X(b,c) <- Dinv_b X(b,c)  (c < b)
        for (int t = warp; t < nbl - 1; t += NW) {
            double c0 = 0.0, c1 = 0.0;
            const double d0 = Db[fr * 12 + fk], d1 = Db[fr * 12 + 4 + fk];
            if (t < nbl - 1 - b) {
                double* T = H + b0 * ld + (b + 1 + t) * 8;
                const double a0 = T[la], a1 = T[la + 4 * ld];
                dmma884(c0, c1, a0, d0);
                dmma884(c0, c1, a1, d1);
                T[(2 * fk) * ld + fr] = c0;
                T[(2 * fk + 1) * ld + fr] = c1;
            } else {
                const int cb = t - (nbl - 1 - b), xl = f2_xld(nbl, cb);
                double* T = X + f2_xbase(nbl, cb) + (b - cb) * 8;
                const double x0 = T[fr * xl + fk], x1 = T[fr * xl + 4 + fk];
                dmma884(c0, c1, d0, x0);
                dmma884(c0, c1, d1, x1);
                T[(2 * fk) * xl + fr] = c0;
                T[(2 * fk + 1) * xl + fr] = c1;
            }
        }
        // (2) panel visible.  Warp 0 only needs the panel tile it wrote itself (row b+1) for the one trailing tile it
        // owns, so it merely arrives (named barrier 1) and runs ahead into the next diagonal factorisation.
        if (NW > 1) {
            if (warp == 0) { asm volatile("bar.arrive 1, %0;" ::"n"(NW * 32) : "memory"); __syncwarp(); }
            else asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
        } else __syncwarp();
        // ---- trailing update: rows i > b, tiles j = 0..i:  j > b: H(i,j) -= L_ib L_jb';  j <= b: X(i,j) -= L_ib X(b,j)
        // Warp 0 takes the next diagonal tile only (flat index b+1; it then goes on to factor it); the others share
        // the rest.
        {
            const int ntile = f2_ntrail(nbl, b);
            const uint2* dtab = desc + f2_trail_base(nbl, b);
            const int stride = (NW > 1) ? NW - 1 : 1;
            for (int t = (NW > 1) ? warp - 1 : 0;; t += stride) {
                int tf;
                if (NW > 1) {
                    if (warp == 0) { if (t != -1) break; tf = b + 1; }
                    else { tf = t + (t >= b + 1); }
                } else tf = t;
                if (tf >= ntile) break;
                const uint2 d = dtab[tf];
                const int aoff = d.x & 0xffff, boff = d.x >> 16, coff = d.y & 0xffff, cl = (d.y >> 16) & 0xff;
                const bool xform = (d.y >> 24) & 1, zero = (d.y >> 25) & 1;
                const double a0 = -H[aoff + la], a1 = -H[aoff + la + 4 * ld];
                const int lb = xform ? fr * cl + fk : la, sb = xform ? 4 : 4 * ld;
                const double q0 = H[boff + lb], q1 = H[boff + lb + sb];
                double* C = H + coff + (2 * fk) * cl + fr;
                double c0 = zero ? 0.0 : C[0], c1 = zero ? 0.0 : C[cl];
                dmma884(c0, c1, a0, q0);
                dmma884(c0, c1, a1, q1);
                C[0] = c0;
                C[cl] = c1;
            }
        }
    }
    tsync<NW>();
    return 1;
}

// ------------------------------------------------------------------------------------------------ layout providers
// The kernel reads every dimension through a provider.  DimsDyn takes them from the plan (any layout that fits);
// DimsStatic<...> makes them compile-time constants for the BASELINE.json shapes -- the GPU analogue of the
// reference's compile-time specialisation on the cone tuple (`@unroll`, Cone{D}; src/scalings.jl:101-110): loops
// unroll, masks and index arithmetic fold away.
struct DimsDyn {
    static constexpr bool is_static = false;
    __device__ __forceinline__ static int n(const F2Plan& P) { return P.n; }
    __device__ __forceinline__ static int p(const F2Plan& P) { return P.p; }
    __device__ __forceinline__ static int k(const F2Plan& P) { return P.k; }
    __device__ __forceinline__ static int kpoc(const F2Plan& P) { return P.kpoc; }
    __device__ __forceinline__ static int nsoc(const F2Plan& P) { return P.nsoc; }
    __device__ __forceinline__ static int lpc(const F2Plan& P) { return P.lpc; }
    __device__ __forceinline__ static int npad(const F2Plan& P) { return P.npad; }
    __device__ __forceinline__ static int kpad(const F2Plan& P) { return P.kpad; }
    __device__ __forceinline__ static int ldg(const F2Plan& P) { return P.ldg; }
    __device__ __forceinline__ static int ldh(const F2Plan& P) { return P.ldh; }
    __device__ __forceinline__ static int ppad(const F2Plan& P) { return P.ppad; }
    __device__ __forceinline__ static int ldm(const F2Plan& P) { return P.ldm; }
    __device__ __forceinline__ static int split_k(const F2Plan& P) { return P.split_k; }
    __device__ __forceinline__ static int split_n(const F2Plan& P) { return P.split_n; }
    __device__ __forceinline__ static int split_p(const F2Plan& P) { return P.split_p; }
    __device__ __forceinline__ static int soc_offs(const F2Plan& P, int slot) { return P.soc_offs[slot]; }
    __device__ __forceinline__ static int soc_dim(const F2Plan& P, int slot) { return P.soc_dim[slot]; }
};
// N variables, PE equality rows, one POC block of KPOC rows followed by NSOC second-order cones of dimension SDIM
template <int NW, int N, int PE, int KPOC, int NSOC, int SDIM>
struct DimsStatic {
    static constexpr bool is_static = true;
    static constexpr int K = KPOC + NSOC * SDIM;
    static constexpr int NPAD = (N + 7) / 8 * 8, KPAD = (K + 3) / 4 * 4, PPAD = ((PE > 0 ? PE : 1) + 7) / 8 * 8;
    __device__ __forceinline__ static constexpr int n(const F2Plan&) { return N; }
    __device__ __forceinline__ static constexpr int p(const F2Plan&) { return PE; }
    __device__ __forceinline__ static constexpr int k(const F2Plan&) { return K; }
    __device__ __forceinline__ static constexpr int kpoc(const F2Plan&) { return KPOC; }
    __device__ __forceinline__ static constexpr int nsoc(const F2Plan&) { return NSOC; }
    __device__ __forceinline__ static constexpr int lpc(const F2Plan&) { return f2_lpc(SDIM); }
    __device__ __forceinline__ static constexpr int npad(const F2Plan&) { return NPAD; }
    __device__ __forceinline__ static constexpr int kpad(const F2Plan&) { return KPAD; }
    __device__ __forceinline__ static constexpr int ldg(const F2Plan&) { return f2_ld(KPAD); }
    __device__ __forceinline__ static constexpr int ldh(const F2Plan&) { return f2_ld(NPAD); }
    __device__ __forceinline__ static constexpr int ppad(const F2Plan&) { return PPAD; }
    __device__ __forceinline__ static constexpr int ldm(const F2Plan&) { return f2_ld(PPAD); }
    __device__ __forceinline__ static constexpr int split_k(const F2Plan&) { return f2_split((K + 1) / 2, NW); }
    __device__ __forceinline__ static constexpr int split_n(const F2Plan&) { return f2_split(N, NW); }
    __device__ __forceinline__ static constexpr int split_p(const F2Plan&) { return f2_split(PE > 0 ? PE : 1, NW); }
    __device__ __forceinline__ static constexpr int soc_offs(const F2Plan&, int slot) { return KPOC + slot * SDIM; }
    __device__ __forceinline__ static constexpr int soc_dim(const F2Plan&, int) { return SDIM; }
    static bool matches(const F2Plan& P) {
        if (P.n != N || P.p != PE || P.kpoc != KPOC || P.nsoc != NSOC || P.k != K || P.nw != NW) return false;
        for (int i = 0; i < NSOC; ++i)
            if (P.soc_dim[i] != SDIM || P.soc_offs[i] != KPOC + i * SDIM) return false;
        return true;
    }
};

// ------------------------------------------------------------------------------------------------ kernel
struct F2Args {
    Ws g;                 // global arrays of the shard
    F2Plan P;
    LoopParams prm;
    int first, batch;     // this launch solves problems [first, first + batch) of the shard
    int* counter;
};

#ifdef SOCP_PHASE_TIMING
#define PT2_DECL() long long pt_t0 = 0
#define PT2_INIT() pt_t0 = clock64()
#define PT2_MARK(idx)                                                             \
    do {                                                                          \
        if (tid == 0 && blockIdx.x == 0) {                                        \
            const long long t_ = clock64();                                       \
            atomicAdd(&a.P.d_clk[idx], (unsigned long long)(t_ - pt_t0));      \
            pt_t0 = t_;                                                           \
        }                                                                         \
    } while (0)
#else
#define PT2_DECL()
#define PT2_INIT()
#define PT2_MARK(idx)
#endif
enum { P2_LOAD = 0, P2_RESID, P2_HEAD_GT, P2_SYRK, P2_CHOL, P2_EQ, P2_SOLVE, P2_INIT, P2_TAIL, P2_MIDPOST, P2_OUT,
       P2_XTX, P2_SOLVE_A, P2_SOLVE_B, P2_N0 };


template <int NW, int MAXT, int MINB, class D>
__global__ void __launch_bounds__(NW * 32, MINB) k_fused2(const F2Args a) {
    extern __shared__ __align__(16) double sm[];
    __shared__ int s_prob, s_fail;
    const F2Plan& P = a.P;
    int tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int T = NW * 32;
    const int n = D::n(P), p = D::p(P), k = D::k(P), kpoc = D::kpoc(P), nsoc = D::nsoc(P), lpc = D::lpc(P);
    const int ldg = D::ldg(P), ldh = D::ldh(P), npad = D::npad(P), nb = npad / 8, kpad = D::kpad(P);
    const int ppad = D::ppad(P), pb = ppad / 8, ldm = D::ldm(P);
    double* G = sm + P.oG;
    double* H = sm + P.oR;           // G'W^-2 G, then H^-1, then the solve matrix Pm = H^-1 - K (H^-1 A')'
    double* X = sm + P.oR + ldh * npad;
    double* A = sm + P.oA;
    double* At = sm + P.oAt;         // A' (rows of A contiguous, stride npad)
    double* HiAt = sm + P.oHiAt;     // H^-1 A'         (n x p)
    double* Km = sm + P.oK;          // H^-1 A' M^-1    (n x p)
    double* Mm = sm + P.oM;
    double* MX = sm + P.oMX;
    double* Minv = sm + P.oMinv;
    double* Dinv = sm + P.oDinv;
    double* cv = sm + P.oc;
    double* bv = sm + P.ob;
    double* hv = sm + P.oh;
    double* x = sm + P.ox; double* y = sm + P.oy; double* z = sm + P.oz; double* s = sm + P.os;
    double* lam = sm + P.olam; double* wb = sm + P.owb; double* iwb = sm + P.oiwb; double* cs = sm + P.ocs;
    double* dx = sm + P.odx; double* dy = sm + P.ody; double* dz = sm + P.odz; double* ds = sm + P.ods;
    double* k0 = sm + P.ok0; double* k2 = sm + P.ok2; double* u = sm + P.ou;
    double* n0 = sm + P.on0; double* cx = sm + P.ocx; double* cy = sm + P.ocy;
    double* kd = sm + P.okd;         // K dy            (n)
    double* md = sm + P.omd;         // M^-1 dy         (p)
    double* dw = sm + P.odw;         // row weights of the SYRK (kpad)
    double* hc = sm + P.ohc;         // sqrt(2)/eta G_c'q per cone (nsoc x npad)
    double* scr = sm + P.oscr;
    int scr_par = 0;
    const LoopParams prm = a.prm;
    PT2_DECL();

    for (int q = tid; q < P.total; q += T) sm[q] = 0.0;
    tsync<NW>();
    // tables that serve every factorisation of this kernel: trailing-update tile descriptors, SYRK tile lists
    uint2* desc = reinterpret_cast<uint2*>(sm + P.odesc);
    uint2* descm = reinterpret_cast<uint2*>(sm + P.odescm);
    f2_build_trail(desc, nb, ldh, ldh * npad, tid, T);
    if (p > 0) f2_build_trail(descm, pb, ldm, P.oMX - P.oM, tid, T);
    int tl[MAXT];
    f2_my_tiles<NW, MAXT>(nb, warp, tl);
    for (int i = p + tid; i < ppad; i += T) Mm[i * ldm + i] = 1.0;       // unit pad diagonal of M (never overwritten)

    // ------------------------------------------------------------------ building blocks (lambdas over the work set)
    // calls f(SocLane, slot) for every cone slot of this lane's group; all 32 lanes take part in every pass
    auto make_lane = [&](int slot, bool valid) {
        SocLane L;
        L.valid = valid;
        L.offs = valid ? D::soc_offs(P, slot) : 0;
        const int dim = valid ? D::soc_dim(P, slot) : 0;
        L.g = lane & (lpc - 1);
        L.lpc = lpc;
        L.tm = 0;
        F2_FOR_E { const int i = L.g + e * lpc; if (i > 0 && i < dim) L.tm |= 1u << e; }
        return L;
    };
    auto for_each_slot = [&](auto&& f) {
        const int spw = 32 / lpc;
        for (int base = warp * spw; base < nsoc; base += NW * spw) {
            const int slot = base + lane / lpc;
            const bool valid = slot < nsoc;
            f(make_lane(valid ? slot : 0, valid), valid ? slot : 0);
        }
    };
    auto load_tail = [&](const SocLane& L, const double* v, double (&r)[4]) {
        F2_FOR_E r[e] = L.tail(e) ? v[L.at(e)] : 0.0;
    };
    auto wdot = [&](const double (&w)[4], const double (&v)[4]) {     // tails only: masked entries are zero
        double d = 0.0;
        F2_FOR_E d = fma(w[e], v[e], d);
        return grp_sum(d, lpc);
    };

    // compute_scaling for one second-order cone, src/scalings.jl:32-99 (closed forms)
    auto soc_scaling = [&](const SocLane& L, int slot, double& gap, double& ll) -> int {
        double sv[4], zv[4];
        load_tail(L, s, sv);
        load_tail(L, z, zv);
        const double s0 = L.valid ? s[L.offs] : 1.0, z0 = L.valid ? z[L.offs] : 1.0;
        double ss = 0.0, zz = 0.0, sz = 0.0;
        F2_FOR_E { ss = fma(sv[e], sv[e], ss); zz = fma(zv[e], zv[e], zz); sz = fma(sv[e], zv[e], sz); }
        ss = grp_sum(ss, lpc); zz = grp_sum(zz, lpc); sz = grp_sum(sz, lpc);
        const double onrms = s0 * s0 - ss, onrmz = z0 * z0 - zz;          // :39-45
        int fail = !(onrms >= 0.0) | !(onrmz >= 0.0);
        const double is = fast_rsqrt(onrms), iz = fast_rsqrt(onrmz);      // :46-49
        const double nrms = onrms * is, nrmz = onrmz * iz;
        const double sb0 = s0 * is, zb0 = z0 * iz;
        const double ns = sz * (is * iz) + zb0 * sb0;                     // :53-56
        const double g2 = (1.0 + ns) / 2.0;
        fail |= !(g2 >= 0.0);
        const double rg = fast_rsqrt(g2);
        const double gamma = g2 * rg, ig = 0.5 * rg;                      // :57, :64
        const double eta = fast_sqrt(nrms * iz);                          // :68
        const double ie = fast_rcp(eta), ie2 = ie * ie;
        const double tmv1 = fast_sqrt(nrms * nrmz);                       // :91
        const double mult = tmv1 * fast_rcp(zb0 + sb0 + 2.0 * gamma);     // :92
        const double csf = gamma + zb0, czf = gamma + sb0;                // :93-94
        double llt = 0.0;
        F2_FOR_E if (L.tail(e)) {
            const double sb = sv[e] * is, zb = zv[e] * iz;
            const double lv = (sb * csf + zb * czf) * mult;               // :95-97
            wb[L.at(e)] = (sb - zb) * ig;                                 // :62,:64
            lam[L.at(e)] = lv;
            dw[L.at(e)] = ie2;                                            // W^-2 = eta^-2 (2 q q' - J): tail weight
            llt = fma(lv, lv, llt);
        }
        llt = grp_sum(llt, lpc);
        if (L.head()) {
            const double w0 = (sb0 + zb0) * ig, l0 = gamma * tmv1;
            wb[L.offs] = w0;                                              // :60
            lam[L.offs] = l0;                                             // :98
            double* c = cs + slot * F2_CS;
            c[CS_ETA] = eta; c[CS_IE] = ie; c[CS_IE2] = ie2; c[CS_R1W] = fast_rcp(1.0 + w0);
            c[CS_W0] = w0; c[CS_LAM0] = l0; c[CS_A] = l0 * l0 - llt; c[CS_LLT] = llt;
            dw[L.offs] = -ie2;                                            // head weight
            gap += s0 * z0 + sz;
            ll += l0 * l0 + llt;
        }
        return L.valid ? fail : 0;
    };

    // solve_kkt head for one cone (src/densesolver.jl:61-66, then the W^-2 of :86 applied to k2), from ds (shared
    // memory) and dz scaled by dzs:  k0 = lam \ ds, k2 = dzs*dz - W k0, u = W^-2 k2.
    auto soc_head = [&](const SocLane& L, int slot, double dzs) {
        const double* c = cs + slot * F2_CS;
        const double eta = c[CS_ETA], ie2 = c[CS_IE2], r1w = c[CS_R1W], w0 = c[CS_W0], l0 = c[CS_LAM0], aa = c[CS_A];
        double lv[4], wv[4], dsv[4], k0v[4], k2v[4];
        load_tail(L, lam, lv);
        load_tail(L, wb, wv);
        load_tail(L, ds, dsv);
        const double ds0 = L.valid ? ds[L.offs] : 0.0;
        const double beta = wdot(lv, dsv);
        const double ia = fast_rcp(aa), il0 = fast_rcp(l0);
        const double k00 = (l0 * ds0 - beta) * ia;                                   // src/vectors.jl:105-125, O(d) form
        F2_FOR_E k0v[e] = L.tail(e) ? (-ds0 * lv[e] + (aa * dsv[e] + beta * lv[e]) * il0) * ia : 0.0;
        const double dl = wdot(wv, k0v);
        const double cst = k00 + dl * r1w;                                           // src/scalings.jl:135
        const double k20 = (L.valid ? dz[L.offs] * dzs : 0.0) - eta * (w0 * k00 + dl);   // :136, densesolver :65
        F2_FOR_E k2v[e] = L.tail(e) ? dz[L.at(e)] * dzs - eta * (k0v[e] + cst * wv[e]) : 0.0;   // :137-139
        const double qv = w0 * k20 - wdot(wv, k2v);                                  // W^-2 = eta^-2 (2 q q' - J)
        F2_FOR_E if (L.tail(e)) {
            k0[L.at(e)] = k0v[e];
            k2[L.at(e)] = k2v[e];
            u[L.at(e)] = ie2 * (k2v[e] - 2.0 * wv[e] * qv);
        }
        if (L.head()) {
            k0[L.offs] = k00;
            k2[L.offs] = k20;
            u[L.offs] = ie2 * (2.0 * w0 * qv - k20);
        }
    };

    // solve_kkt tail (src/densesolver.jl:86-89), the driver's scale!/iscale! (src/solver.jl:128-129) and scmax of both
    // results (src/mats.jl:64-86) for one cone.  On exit u <- cz, k0 <- cs, k2 <- kt2 o kt3 (Jordan product).
    auto soc_tail = [&](const SocLane& L, int slot, double& mx, double& dotacc, int& fl, bool chk) {
        const double* c = cs + slot * F2_CS;
        const double eta = c[CS_ETA], ie = c[CS_IE], ie2 = c[CS_IE2], r1w = c[CS_R1W], w0 = c[CS_W0], l0 = c[CS_LAM0],
                     aa = c[CS_A];
        double wv[4], lv[4], uv[4], k0v[4], czv[4], csv[4], kt2v[4], kt3v[4];
        load_tail(L, wb, wv);
        load_tail(L, lam, lv);
        load_tail(L, u, uv);
        load_tail(L, k0, k0v);
        const double u0 = L.valid ? u[L.offs] : 0.0;
        const double k00 = L.valid ? k0[L.offs] : 0.0;
        const double qv = w0 * u0 - wdot(wv, uv);                                    // cz = W^-2 u          :86
        const double cz0 = ie2 * (2.0 * w0 * qv - u0);
        F2_FOR_E czv[e] = L.tail(e) ? ie2 * (uv[e] - 2.0 * wv[e] * qv) : 0.0;
        double dl = wdot(wv, czv);                                                   // kt3 = W cz           :87, solver :128
        double cst = cz0 + dl * r1w;
        const double kt30 = eta * (w0 * cz0 + dl);
        F2_FOR_E kt3v[e] = L.tail(e) ? eta * (czv[e] + cst * wv[e]) : 0.0;
        const double kk0 = k00 - kt30;                                               // k0 -= W cz           :88
        F2_FOR_E k0v[e] -= kt3v[e];
        dl = wdot(wv, k0v);                                                          // cs = W k0            :89
        cst = kk0 + dl * r1w;
        const double cs0 = eta * (w0 * kk0 + dl);
        F2_FOR_E csv[e] = L.tail(e) ? eta * (k0v[e] + cst * wv[e]) : 0.0;
        dl = wdot(wv, csv);                                                          // kt2 = W^-1 cs        solver :129
        cst = -cs0 + dl * r1w;
        const double kt20 = ie * (w0 * cs0 - dl);
        F2_FOR_E kt2v[e] = L.tail(e) ? ie * (csv[e] + cst * wv[e]) : 0.0;
        double lx3 = 0.0, lx2 = 0.0, dot = 0.0;
        F2_FOR_E {
            lx3 = fma(lv[e], kt3v[e], lx3);
            lx2 = fma(lv[e], kt2v[e], lx2);
            dot = fma(kt2v[e], kt3v[e], dot);
        }
        lx3 = grp_sum(lx3, lpc); lx2 = grp_sum(lx2, lpc); dot = grp_sum(dot, lpc);
        dot += kt20 * kt30;
        fl |= L.valid && !(aa >= 0.0);
        const double as = fast_rsqrt(aa);                                            // src/mats.jl:67-71
        const double r13 = as * l0 * kt30 - as * lx3, r12 = as * l0 * kt20 - as * lx2;      // :74-77
        const double den = fast_rcp(as * l0 + 1.0);
        const double c3 = (r13 + kt30) * den, c2 = (r12 + kt20) * den;               // :80
        double q3 = 0.0, q2 = 0.0;
        F2_FOR_E if (L.tail(e)) {
            const double v3 = as * (kt3v[e] - c3 * as * lv[e]);                      // :83
            const double v2 = as * (kt2v[e] - c2 * as * lv[e]);
            q3 = fma(v3, v3, q3);
            q2 = fma(v2, v2, q2);
        }
        q3 = grp_sum(q3, lpc); q2 = grp_sum(q2, lpc);
        if (L.valid) mx = fmax(mx, fmax(fast_sqrt(q3) - as * r13, fast_sqrt(q2) - as * r12));   // :85
        F2_FOR_E if (L.tail(e)) {
            u[L.at(e)] = czv[e];
            k0[L.at(e)] = csv[e];
            k2[L.at(e)] = kt20 * kt3v[e] + kt30 * kt2v[e];                           // src/vectors.jl:73-75
            if (chk) fl |= !isfinite(czv[e]) | !isfinite(csv[e]);
        }
        if (chk && L.valid) fl |= !isfinite(cz0) | !isfinite(cs0);
        if (L.head()) {
            dotacc += dot;
            u[L.offs] = cz0;
            k0[L.offs] = cs0;
            k2[L.offs] = dot;                                                        // src/vectors.jl:66-69
        }
    };

    for (;;) {
        if (tid == 0) s_prob = atomicAdd(a.counter, 1);
        tsync<NW>();
        const int b = s_prob + a.first;
        tsync<NW>();
        if (b >= a.first + a.batch) break;
        PT2_INIT();

        // ---- load the problem (global -> shared)
        {
            const double* Gg = a.g.G + (int64_t)b * a.g.sG;
            for (int col = warp; col < n; col += NW)
                for (int r = lane; r < k; r += 32) G[col * ldg + r] = Gg[(int64_t)col * k + r];
            const double* Ag = a.g.A + (int64_t)b * a.g.sA;
            for (int q = tid; q < p * n; q += T) { const double v = Ag[q]; A[q] = v; At[(q % p) * npad + q / p] = v; }
            for (int i = tid; i < n; i += T) cv[i] = a.g.c[(int64_t)b * n + i];
            for (int i = tid; i < p; i += T) bv[i] = a.g.b[(int64_t)b * p + i];
            for (int i = tid; i < k; i += T) hv[i] = a.g.h[(int64_t)b * k + i];
        }
        int status = ST_RUNNING, iters = 0;
        tsync<NW>();

        // The initial point (src/solver.jl:68-104, W = I) is the same factor + solve as a loop iteration with
        // u = h, dx = -c, dy = b, k2 = h: then cx = x, cy = y and u = G x - h = z0.
        for (int i = tid; i < k; i += T) { u[i] = hv[i]; k2[i] = hv[i]; dw[i] = 1.0; }
        for (int q = tid; q < nsoc * npad; q += T) hc[q] = 0.0;
        for (int i = tid; i < n; i += T) dx[i] = -cv[i];
        for (int i = tid; i < p; i += T) dy[i] = bv[i];
        tsync<NW>();
        PT2_MARK(P2_LOAD);

        int phase = 0;              // 0: initial point, 1: affine direction (solve #1), 2: combined direction (solve #2)
        double sc = 1.0;            // (1 - sigma) applied to dx, dy, dz in solve #2 (src/solver.jl:140)
        double ll = 0.0;            // lambda'lambda of the current iteration
        for (;;) {
            // ---- n0 = G'u + sc*dx                                                   src/densesolver.jl:66-67
            gemv_cols_v<NW>(G, ldg, k, n, u, lane, warp, [&](int c, double acc) { n0[c] = acc + sc * dx[c]; });
            PT2_MARK(P2_N0);
            if (phase != 2) {
                // ---- KKT factor, src/densesolver.jl:41-52
                f2_syrk_w<NW, MAXT>(G, ldg, kpad, n, nb, dw, hc, nsoc, npad, H, ldh, tl, lane, warp);   // :42-43
                for (int q = tid; q < f2_xsize(nb); q += T) X[q] = 0.0;
                if (tid == 0) s_fail = 0;
                tsync<NW>();
                PT2_MARK(P2_SYRK);
                int ok = f2_chol_inv<NW>(H, X, Dinv, desc, nb, ldh, &s_fail, lane, warp, a.P.d_clk);          // :47
                PT2_MARK(P2_XTX);              // slot 11 = f2_chol_inv; slot P2_CHOL below = f2_xtx (tools/phase_timing.py)
                if (ok) f2_xtx<NW, MAXT>(X, ldh, nb, H, tl, lane, warp);                         // :48  Li = H^-1 (explicit)
                tsync<NW>();
                PT2_MARK(P2_CHOL);
                if (ok && p > 0) {
                    for (int q = 0; q < p; ++q)          // HiAt = H^-1 A'                           :49
                        gemv_cols_v<NW>(H, ldh, n, n, At + q * npad, lane, warp, [&](int c, double acc) { HiAt[q * n + c] = acc; });
                    if (pb > 1) for (int q = tid; q < f2_xsize(pb); q += T) MX[q] = 0.0;
                    tsync<NW>();
                    if (pb == 1) {
                        // p <= 8: warp 0 alone forms M = A HiAt (:50), factors it (:51) and inverts it -- no team barrier
                        if (p == 1) {
                            // one equality row (the budget row of C2): M is a scalar, no 8 x 8 factorisation
                            if (warp == 0) {
                                double acc = 0.0;
                                for (int c = lane; c < n; c += 32) acc = fma(A[c], HiAt[c], acc);
                                acc = warp_sum(acc);
                                if (lane == 0) {
                                    if (!(acc > 0.0)) s_fail = 1;
                                    Minv[0] = 1.0 / acc;
                                }
                            }
                        } else if (warp == 0) {
                            for (int e = 0; e < p * p; ++e) {
                                const int i = e % p, j = e / p;
                                double acc = 0.0;
                                for (int c = lane; c < n; c += 32) acc = fma(A[c * p + i], HiAt[j * n + c], acc);
                                acc = warp_sum(acc);
                                if (lane == 0) Mm[j * ldm + i] = acc;
                            }
                            __syncwarp();
                            double xc[8];
                            const int okm = f2_diag_factor<false>(Mm, ldm, lane, xc);
                            if (!okm && lane == 0) s_fail = 1;
                            if (lane < 8) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) Dinv[i * 12 + lane] = xc[i];      // X = chol(M)^-1, row-major
                            }
                            __syncwarp();
                            for (int e = lane; e < p * p; e += 32) {                        // Minv = X'X
                                const int i = e % p, j = e / p;
                                double acc = 0.0;
                                for (int m = max(i, j); m < p; ++m) acc = fma(Dinv[m * 12 + i], Dinv[m * 12 + j], acc);
                                Minv[j * p + i] = acc;
                            }
                        }
                        tsync<NW>();
                        ok = !s_fail;
                    } else {
                        for (int j = 0; j < p; ++j)          // M = A HiAt                              :50
                            gemv_rows<NW, false>(A, p, p, n, HiAt + j * n, 1, D::split_p(P), lane, warp,
                                                 [&](int r, double acc) { Mm[j * ldm + r] = acc; });
                        tsync<NW>();
                        ok = f2_chol_inv<NW>(Mm, MX, Dinv, descm, pb, ldm, &s_fail, lane, warp);        // :51
                        if (ok) {
                            for (int q = tid; q < p * p; q += T) {          // Minv = MX' MX
                                const int i = q % p, j = q / p;
                                double acc = 0.0;
                                const double* xi = MX + f2_xbase(pb, i >> 3) + (i & 7) * f2_xld(pb, i >> 3) - (i & ~7);
                                const double* xj = MX + f2_xbase(pb, j >> 3) + (j & 7) * f2_xld(pb, j >> 3) - (j & ~7);
                                for (int m = max(i, j); m < p; ++m) acc = fma(xi[m], xj[m], acc);
                                Minv[j * p + i] = acc;
                            }
                            tsync<NW>();
                        }
                    }
                    if (ok) {
                        // K = HiAt Minv, md = Minv dy, kd = K dy and Pm = H^-1 - K HiAt' (in place): then cx = Pm n0 + kd,
                        // cy = K'n0 - md, which is src/densesolver.jl:73-83 (m0 = A H^-1 n0 - dy, cy = M^-1 m0,
                        // cx = H^-1 (n0 - A'cy)) in one pass.  Every lane keeps the rows of K it needs in registers.
                        for (int r0 = 0; r0 < n; r0 += 32) {
                            const int r = r0 + lane;
                            for (int q = 0; q < p; ++q) {
                                double kq = 0.0;
                                if (r < n)
                                    for (int t = 0; t < p; ++t) kq = fma(HiAt[t * n + r], Minv[q * p + t], kq);
                                for (int col = warp; col < n; col += NW)
                                    if (r < n) H[col * ldh + r] = fma(-kq, HiAt[q * n + col], H[col * ldh + r]);
                                if (warp == 0 && r < n) Km[q * n + r] = kq;
                            }
                        }
                        for (int i = tid; i < p; i += T) {              // md = Minv dy
                            double acc = 0.0;
                            for (int r = 0; r < p; ++r) acc = fma(Minv[r * p + i], dy[r], acc);
                            md[i] = acc;
                        }
                        if (warp == NW - 1) {                           // kd = K dy = HiAt (Minv dy)
                            for (int i = lane; i < n; i += 32) {
                                double acc = 0.0;
                                for (int q = 0; q < p; ++q) {
                                    double mq = 0.0;
                                    for (int r = 0; r < p; ++r) mq = fma(Minv[r * p + q], dy[r], mq);
                                    acc = fma(HiAt[q * n + i], mq, acc);
                                }
                                kd[i] = acc;
                            }
                        }
                    }
                }
                PT2_MARK(P2_EQ);
                if (!ok) { status = ST_NUMERICAL; break; }                                   // cholesky! threw
            }
            // ---- middle of solve_kkt, src/densesolver.jl:66-85: out cx, cy, u = G cx - k2
            tsync<NW>();
            gemv_cols_v<NW>(H, ldh, n, n, n0, lane, warp, [&](int c, double acc) { cx[c] = p > 0 ? acc + sc * kd[c] : acc; });
            if (p > 0)
                gemv_cols<NW, false>(Km, n, n, p, n0, lane, warp, [&](int c, double acc) { cy[c] = acc - sc * md[c]; });
            tsync<NW>();
            PT2_MARK(P2_SOLVE_A);
            const double* cxv = cx;
            gemv_rows_v<NW>(G, ldg, k, n, cxv, D::split_k(P), lane, warp, [&](int r, double acc) { u[r] = acc - k2[r]; });   // :84-85
            tsync<NW>();
            PT2_MARK(P2_SOLVE);

            bool new_iter;
            if (phase == 0) {
                for (int i = tid; i < n; i += T) x[i] = cxv[i];
                for (int i = tid; i < p; i += T) y[i] = cy[i];
                // max_step(-z0), max_step(z0), src/mats.jl:1-28, then the shift of src/solver.jl:91-101
                double r2[2] = {-INFINITY, -INFINITY};
                int fl = 0;
                for (int i = tid; i < kpoc; i += T) { const double v = u[i]; r2[0] = fmax(r2[0], v); r2[1] = fmax(r2[1], -v); }
                for_each_slot([&](const SocLane& L, int) {
                    double zv[4];
                    load_tail(L, u, zv);
                    double sq = 0.0;
                    F2_FOR_E sq = fma(zv[e], zv[e], sq);
                    const double nr = fast_sqrt(grp_sum(sq, lpc));
                    if (L.valid) {
                        const double z0 = u[L.offs];
                        r2[0] = fmax(r2[0], nr + z0);          // ||-z1|| - (-z0)
                        r2[1] = fmax(r2[1], nr - z0);
                    }
                });
                team_reduce<NW, 0, 2>(r2, fl, scr + (scr_par ^= 1) * 64, lane, warp);
                const double mp = r2[0], md = r2[1];
                const bool shp = !(fabs(mp) < prm.init_eps), shd = !(fabs(md) < prm.init_eps);
                for (int i = tid; i < k; i += T) {
                    const double z0 = u[i];
                    s[i] = -z0;
                    z[i] = z0;
                }
                tsync<NW>();
                for (int i = tid; i < kpoc; i += T) {
                    if (shp) s[i] += 1.0 + mp;
                    if (shd) z[i] += 1.0 + md;
                }
                for (int c = tid; c < nsoc; c += T) {
                    const int o = D::soc_offs(P, c);
                    if (shp) s[o] += 1.0 + mp;
                    if (shd) z[o] += 1.0 + md;
                }
                tsync<NW>();
                PT2_MARK(P2_INIT);
                new_iter = true;
            } else {
                // ---- tail of solve_kkt + scale!/iscale! + scmax for every cone (both solves)
                double r2[2] = {0.0, -INFINITY};       // kt2'kt3, max scmax
                int fl = 0;
                for_each_slot([&](const SocLane& L, int slot) { soc_tail(L, slot, r2[1], r2[0], fl, phase == 2); });
                for (int i = tid; i < kpoc; i += T) {
                    const double w = wb[i], iw = iwb[i], il = fast_rcp(lam[i]);
                    const double cz = iw * iw * u[i];
                    const double kt3 = w * cz;
                    const double kk = k0[i] - kt3;
                    const double csx = w * kk;
                    const double kt2 = iw * csx;
                    r2[1] = fmax(r2[1], fmax(-kt3 * il, -kt2 * il));                 // src/mats.jl:53-62
                    r2[0] = fma(kt2, kt3, r2[0]);
                    u[i] = cz;
                    k0[i] = csx;
                    k2[i] = kt2 * kt3;
                }
                if (phase == 2) {      // the reference would carry NaN/Inf into the next cholesky! and throw there
                    for (int i = tid; i < n; i += T) fl |= !isfinite(cxv[i]);
                    for (int i = tid; i < p; i += T) fl |= !isfinite(cy[i]);
                    for (int i = tid; i < kpoc; i += T) fl |= !isfinite(u[i]) | !isfinite(k0[i]);
                }
                PT2_MARK(P2_TAIL);
                team_reduce<NW, 1, 1>(r2, fl, scr + (scr_par ^= 1) * 64, lane, warp);
                if (NW == 1) __syncwarp();
                const double tstep = step_from_t(r2[1]);                             // src/solver.jl:130 / :145
                if (phase == 1) {
                    // centering parameter (:130-134), combined right-hand side (:136-140)
                    const double rho = 1.0 - tstep - tstep * tstep * r2[0] * fast_rcp(ll);   // :132 (minus: reference quirk)
                    const double cl = fmax(0.0, fmin(1.0, rho));
                    const double sig = cl * cl * cl;                                 // :133
                    const double mu = ll / (double)a.g.L.deg;                        // :134
                    if (fl) { status = ST_NUMERICAL; break; }
                    const double smu = sig * mu;
                    sc = 1.0 - sig;                                                  // :136
                    for (int i = tid; i < kpoc; i += T) ds[i] += smu - k2[i];        // :137-139
                    for_each_slot([&](const SocLane& L, int) {
                        F2_FOR_E if (L.tail(e)) ds[L.at(e)] -= k2[L.at(e)];
                        if (L.head()) ds[L.offs] += smu - k2[L.offs];
                    });
                    phase = 2;
                    new_iter = false;
                } else {
                    // step length (:143-146), iterate update (:147-150)
                    const double step = tstep * prm.step_damp;
                    fl |= !isfinite(step);
                    if (fl) { status = ST_NUMERICAL; break; }
                    for (int i = tid; i < n; i += T) x[i] = fma(cxv[i], step, x[i]);     // :147
                    for (int i = tid; i < p; i += T) y[i] = fma(cy[i], step, y[i]);      // :148
                    for (int i = tid; i < k; i += T) {
                        z[i] = fma(u[i], step, z[i]);                                    // :149
                        s[i] = fma(k0[i], step, s[i]);                                   // :150
                    }
                    ++iters;
                    tsync<NW>();
                    new_iter = true;
                }
                PT2_MARK(P2_MIDPOST);
            }

            if (new_iter) {
                // ---- top of a Mehrotra iteration, src/solver.jl:105-126
                if (iters >= prm.max_iter) break;
                // compute_scaling (:106) and the negated residuals (:110-118,:125) in one phase
                double r4[4] = {0.0, 0.0, 0.0, 0.0};        // |rx|^2, |ry|^2, z's, lambda'lambda
                int fl = 0;
                for_each_slot([&](const SocLane& L, int slot) { fl |= soc_scaling(L, slot, r4[2], r4[3]); });
                for (int i = tid; i < kpoc; i += T) {                                   // src/scalings.jl:22-30
                    const double si = s[i], zi = z[i];
                    const double q = si * fast_rcp(zi), qi = zi * fast_rcp(si), pz = si * zi;
                    fl |= !(q >= 0.0) | !(pz >= 0.0);
                    const double lv = fast_sqrt(pz);
                    const double iw = fast_sqrt(qi);
                    wb[i] = fast_sqrt(q);
                    iwb[i] = iw;
                    dw[i] = iw * iw;
                    lam[i] = lv;
                    r4[2] = fma(si, zi, r4[2]);
                    r4[3] = fma(lv, lv, r4[3]);
                }
                gemv_cols_v<NW>(G, ldg, k, n, z, lane, warp, [&](int c, double acc) {
                    double v = -acc - cv[c];
                    for (int q = 0; q < p; ++q) v = fma(-A[c * p + q], y[q], v);
                    dx[c] = v;
                    r4[0] = fma(v, v, r4[0]);
                });
                gemv_rows_v<NW>(G, ldg, k, n, x, D::split_k(P), lane, warp, [&](int r, double acc) { dz[r] = -acc - s[r] + hv[r]; });
                if (p > 0)
                    gemv_rows<NW, false>(A, p, p, n, x, 1, D::split_p(P), lane, warp, [&](int r, double acc) {
                        const double v = -acc + bv[r];
                        dy[r] = v;
                        r4[1] = fma(v, v, r4[1]);
                    });
                team_reduce<NW, 4, 0>(r4, fl, scr + (scr_par ^= 1) * 64, lane, warp);
                if (NW == 1) __syncwarp();
                PT2_MARK(P2_RESID);
                if (fl) { status = ST_NUMERICAL; break; }                              // compute_scaling threw
                const double resid = sqrt(r4[0]) + sqrt(r4[1]) + r4[2];
                if (resid < prm.tol) { status = ST_CONVERGED; break; }                  // :122-124
                ll = r4[3];
                // affine right-hand side ds = -lam o lam (:120,:125)
                for (int i = tid; i < kpoc; i += T) { const double lv = lam[i]; ds[i] = -(lv * lv); }
                for_each_slot([&](const SocLane& L, int slot) {
                    const double* c = cs + slot * F2_CS;
                    const double l0 = c[CS_LAM0];
                    F2_FOR_E if (L.tail(e)) { const double lv = lam[L.at(e)]; ds[L.at(e)] = -(l0 * lv + l0 * lv); }   // src/vectors.jl:73-75
                    if (L.head()) ds[L.offs] = -(c[CS_LLT] + l0 * l0);                  // :66-69
                });
                // hc[c] = sqrt(2)/eta G_c'q, q = J wbar: the rank-one part of G'W^-2 G per cone (densesolver :41-43)
                {
                    const int spw = 32 / lpc, npairs = n * nsoc;
                    for (int base = warp * spw; base < npairs; base += NW * spw) {
                        const int pr = base + lane / lpc;
                        const bool valid = pr < npairs;
                        const int col = valid ? pr / nsoc : 0;
                        const int slot = valid ? pr - col * nsoc : 0;
                        const SocLane L = make_lane(slot, valid);
                        const double* c = cs + slot * F2_CS;
                        const double* gc = G + col * ldg;
                        double wv[4], gv[4];
                        load_tail(L, wb, wv);
                        load_tail(L, gc, gv);
                        const double dl = wdot(wv, gv);
                        if (L.head()) hc[slot * npad + col] = 1.4142135623730951 * c[CS_IE] * (c[CS_W0] * gc[L.offs] - dl);
                    }
                }
                sc = 1.0;
                phase = 1;
            }
            // ---- head of solve_kkt (src/densesolver.jl:61-66 + W^-2) from ds and sc*dz: k0, k2, u
            __syncwarp();       // ds[head] was written by the cone's head lane, every lane of the group reads it
            for_each_slot([&](const SocLane& L, int slot) { soc_head(L, slot, sc); });
            for (int i = tid; i < kpoc; i += T) {
                const double w = wb[i], iw = iwb[i];
                const double kk = ds[i] * fast_rcp(lam[i]);
                const double kz = sc * dz[i] - w * kk;
                k0[i] = kk; k2[i] = kz; u[i] = iw * iw * kz;
            }
            tsync<NW>();
            PT2_MARK(P2_HEAD_GT);
        }
        if (status == ST_RUNNING) status = ST_MAXITER;
        tsync<NW>();

        // ---- results: iterate and objectives (pobj = c'x, dobj = -b'y - h'z)
        {
            const bool dead = (status == ST_NUMERICAL && phase == 0);      // the initial factorisation failed
            double r2[2] = {0.0, 0.0};
            int fl = 0;
            for (int i = tid; i < n; i += T) {
                const double xi = dead ? 0.0 : x[i];
                a.g.x[(int64_t)b * n + i] = xi;
                r2[0] = fma(cv[i], xi, r2[0]);
            }
            for (int i = tid; i < p; i += T) {
                const double yi = dead ? 0.0 : y[i];
                a.g.y[(int64_t)b * p + i] = yi;
                r2[1] = fma(-bv[i], yi, r2[1]);
            }
            for (int i = tid; i < k; i += T) {
                const double zi = dead ? 0.0 : z[i];
                a.g.z[(int64_t)b * k + i] = zi;
                a.g.s[(int64_t)b * k + i] = dead ? 0.0 : s[i];
                r2[1] = fma(-hv[i], zi, r2[1]);
            }
            team_reduce<NW, 2, 0>(r2, fl, scr + (scr_par ^= 1) * 64, lane, warp);
            if (tid == 0) {
                a.g.pobj[b] = r2[0];
                a.g.dobj[b] = r2[1];
                a.g.status[b] = status;
                a.g.iters[b] = iters;
                a.g.active[b] = 0;
                a.g.fail[b] = (status == ST_NUMERICAL);
            }
        }
        tsync<NW>();
        PT2_MARK(P2_OUT);
    }
}

// compile-time specialised layouts (BASELINE.json): C2 = portfolio n=50, p=1, POC 50 + SOC 51; C3 = n=12, 10 x SOC 4
using DimsC2 = DimsStatic<8, 50, 1, 50, 1, 51>;
using DimsC3 = DimsStatic<1, 12, 0, 0, 10, 4>;

#ifndef __CUDACC_RTC__
template <int NW, int MAXT, int MINB, class D>
inline void fused2_launch(const F2Plan& plan, const F2Args& args, int grid, cudaStream_t stream) {
    cudaFuncSetAttribute(k_fused2<NW, MAXT, MINB, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
    k_fused2<NW, MAXT, MINB, D><<<grid, NW * 32, plan.smem, stream>>>(args);
}

// run-time specialisation of the kernel for the plan's layout (lane_jit.cu): DimsStatic per layout instead of DimsDyn
void* fused2_jit_get(const F2Plan& plan);
bool fused2_jit_launch(void* fn, const F2Plan& plan, const F2Args& args, int grid, cudaStream_t stream);
// tries it once per plan (callers outside their timed regions)
inline void fused2_prepare(F2Plan& plan) {
    if (plan.jit_tried || !plan.fits) return;
    plan.jit_tried = 1;
    if (getenv("SOCP_B200_NO_F2_JIT") || getenv("SOCP_B200_GENERIC_ONLY")) return;
    plan.jit_fn = fused2_jit_get(plan);
}

// Solves problems [first, first + batch) of the shard.
// counter_slot: which of the plan's work counters this launch uses (launches that may overlap need different ones).
inline void solve_fused2(F2Plan& plan, const Ws& g, int first, int batch, int max_iter, double tol, double step_damp,
                         double init_eps, cudaStream_t stream, bool allow_static = true, int counter_slot = 0) {
    cudaMemsetAsync(plan.d_counter + counter_slot, 0, sizeof(int), stream);
    F2Args args;
    args.g = g;
    args.P = plan;
    args.prm = LoopParams{max_iter, tol, step_damp, init_eps};
    args.first = first;
    args.batch = batch;
    args.counter = plan.d_counter + counter_slot;
    const int grid = std::min(batch, plan.num_sms * plan.ctas_per_sm);
    if (allow_static && DimsC2::matches(plan)) { fused2_launch<8, 4, 2, DimsC2>(plan, args, grid, stream); return; }
    if (allow_static && DimsC3::matches(plan)) { fused2_launch<1, 3, 16, DimsC3>(plan, args, grid, stream); return; }
    if (allow_static && plan.jit_fn && fused2_jit_launch(plan.jit_fn, plan, args, grid, stream)) return;
    switch (plan.variant) {
        case 0: fused2_launch<1, 3, 16, DimsDyn>(plan, args, grid, stream); break;     // n <= 16: one warp per problem
        case 1: fused2_launch<4, 3, 4, DimsDyn>(plan, args, grid, stream); break;      // n <= 32: 10 tiles over 4 warps
        case 2: fused2_launch<8, 4, 2, DimsDyn>(plan, args, grid, stream); break;      // n <= 56: 28 tiles over 8 warps
        default: fused2_launch<8, 5, 2, DimsDyn>(plan, args, grid, stream); break;     // n <= 64: 36 tiles over 8 warps
    }
}

// non-inline entry of solve_fused2, compiled once in fused2.cu (the kernels are instantiated there only)
void solve_fused2_ext(F2Plan& plan, const Ws& g, int first, int batch, int max_iter, double tol, double step_damp,
                      double init_eps, cudaStream_t stream, bool allow_static, int counter_slot);

#endif  // __CUDACC_RTC__

}  // namespace socp
