// fused_v3.cuh -- whole-solve kernel, second generation: four problems in flight per SM for the n = 50 shapes.
//
// Same algorithm and the same cone code as fused_v2.cuh (one team of NW warps owns one problem from the initial point
// to the last Mehrotra step, reference src/solver.jl:68-152, everything resident in shared memory), with a memory
// plan that is half the size, so that twice as many independent dependency chains share an SM:
//
//   * compressed G.  The reference keeps G sparse (src/Socp.jl:29) and multiplies it sparsely.  Rows of G that hold at
//     most one nonzero (bounds such as x >= 0: rows +-e_j; empty rows such as the head row of a norm constraint) are
//     kept as (column, value) pairs; only the rows from the first to the last row with two or more nonzeros are kept
//     as a dense block G_D.  A singleton row i = (j, g) contributes g u_i to (G'u)_j, g x_j to (G x)_i and
//     d_i g^2 to H_jj -- the SYRK, the six gemv passes and the footprint of G shrink by the share of such rows (C2:
//     101 rows -> 50 dense).  The pattern is one per batch (detected on the device at set_data, or taken from the CSC
//     pattern); a problem that does not conform is reported back (status ST_PATTERN) and re-solved on the dense plan.
//   * H as packed 8x8 tiles of the lower triangle (64 contiguous doubles per tile, swizzled so that the A/B fragment,
//     transposed fragment and accumulator accesses of mma.sync.m8n8k4 are all bank-conflict free) instead of a padded
//     square, and everything on it in place: blocked Cholesky carrying the inverse (X = L^-1 overwrites L tile by
//     tile; the current panel L_ib lives in a side buffer of nb-1 tiles), H^-1 = X'X over X.  The solves use the
//     symmetric packed matrix directly.
//   * `sing` problems (src/Socp.jl:49-56, src/densesolver.jl:44-46,68-80) are solved by the same kernel: A'A is added
//     in the SYRK epilogue, A'dy in the right-hand side; with `sing` unknown the initial factorisation of G'G *is* the
//     reference's test, so a failure there switches the problem to the sing formulation and repeats the initial point.
//
// Restrictions (fused_v2 / the tiled path take everything else): 16 < n <= 64, p <= 32, second-order cones of
// dimension <= 128, at most 64 of them.  Compiles for the host against tests/simt_emu/simt_emu.h (test
// infrastructure) as well.
#pragma once
#include "fused_common.cuh"
#ifndef __CUDACC_RTC__          // (NVRTC compiles the device code of this file alone: run-time specialisation, fused_jit.cu)
#include <vector>
#include <algorithm>
#endif

namespace socp {

constexpr int ST_PATTERN = -2;       // internal: the problem's G does not conform to the batch's row pattern

// ------------------------------------------------------------------------------------------------ packed tiles
// element (r, c) of an 8x8 tile; conflict free for the fragment (r = lane>>2, c = (lane&3) + 4h), transposed fragment
// (r = (lane&3) + 4h, c = lane>>2) and accumulator (r = lane>>2, c = 2(lane&3) + e) accesses of a warp
__host__ __device__ constexpr int t3_g(int c) { return 2 * ((c >> 1) & 1) + ((c & 1) ^ (c >> 2)); }
__host__ __device__ constexpr int t3_off(int r, int c) { return (r & 3) + 32 * (r >> 2) + 4 * t3_g(c) + 16 * (c >> 2); }
__host__ __device__ constexpr int t3_idx(int i, int j) { return i * (i + 1) / 2 + j; }      // tile (i, j), i >= j
struct T3Lane {
    int a0, a1;      // (fr, fk), (fr, fk + 4)
    int c0, c1;      // (fk, fr), (fk + 4, fr)
    int d0, d1;      // (fr, 2 fk), (fr, 2 fk + 1)
};
__device__ __forceinline__ T3Lane t3_lane(int lane) {
    const int fr = lane >> 2, fk = lane & 3;
    T3Lane L;
    L.a0 = t3_off(fr, fk); L.a1 = t3_off(fr, fk + 4);
    L.c0 = t3_off(fk, fr); L.c1 = t3_off(fk + 4, fr);
    L.d0 = t3_off(fr, 2 * fk); L.d1 = t3_off(fr, 2 * fk + 1);
    return L;
}
// element (i, j) of a symmetric matrix held as packed lower tiles with full diagonal tiles
__device__ __forceinline__ double t3_sym(const double* Tt, int i, int j) {
    const int r = max(i, j), c = min(i, j);
    return Tt[t3_idx(r >> 3, c >> 3) * 64 + t3_off(r & 7, c & 7)];
}

// offsets into one team's slice of the dynamic shared memory, in doubles -- a pure function of the layout, so that the
// kernels specialised on a layout at compile time address shared memory with immediates
struct F3Offs {
    int oG, oT, oLp, oA, oAt, oHiAt, oK, oMt, oMLp, oMinv;
    int oc, ox, odx, on0, ocx, okd, osd, oatdy, ob, oy, ody, ocy, omd, ohc;
    int oh, oz, os, olam, owb, oiwb, odz, ods, ok0, ok2, ou, odw, osval, ocs, oscr;
    int otab, otij, otijm;
    int total;
};
__host__ __device__ constexpr F3Offs f3_offsets(int n, int p, int k, int kpoc, int nsoc, int kd, int nsing, int nw) {
    const int npad = (n + 7) / 8 * 8, nb = npad / 8, ntl = nb * (nb + 1) / 2;
    const int pb = ((p > 1 ? p : 1) + 7) / 8, ntlm = pb * (pb + 1) / 2;
    const int kdpad = (kd + 3) / 4 * 4, ldg = f2_ld(kdpad > 4 ? kdpad : 4);
    const int kvl = (k + 5 + 1) / 2 * 2;
    const int ns1 = nsoc > 1 ? nsoc : 1;
    F3Offs O{};
    int at = 0;
#define F3_TAKE(field, cnt) do { O.field = at; at += ((cnt) + 1) / 2 * 2; } while (0)
    F3_TAKE(oG, ldg * n);
    F3_TAKE(oT, ntl * 64);
    F3_TAKE(oLp, (nb - 1 > 3 ? nb - 1 : 3) * 64);      // panel L_ib of the current block column; 2 x (8x12) scratch when p <= 8
    F3_TAKE(oA, p * n); F3_TAKE(oAt, p * npad); F3_TAKE(oHiAt, p * npad); F3_TAKE(oK, p * npad);
    F3_TAKE(oMt, p > 8 ? ntlm * 64 : 0); F3_TAKE(oMLp, p > 8 ? (pb - 1) * 64 : 0); F3_TAKE(oMinv, p * p);
    F3_TAKE(oc, npad); F3_TAKE(ox, npad); F3_TAKE(odx, npad); F3_TAKE(on0, npad); F3_TAKE(ocx, npad);
    F3_TAKE(okd, p ? npad : 0); F3_TAKE(osd, npad); F3_TAKE(oatdy, p ? npad : 0);
    F3_TAKE(ob, p); F3_TAKE(oy, p); F3_TAKE(ody, p); F3_TAKE(ocy, p); F3_TAKE(omd, p);
    F3_TAKE(ohc, ns1 * npad);
    F3_TAKE(oh, kvl); F3_TAKE(oz, kvl); F3_TAKE(os, kvl); F3_TAKE(olam, kvl); F3_TAKE(owb, kvl);
    F3_TAKE(odz, kvl); F3_TAKE(ods, kvl); F3_TAKE(ok0, kvl); F3_TAKE(ok2, kvl); F3_TAKE(ou, kvl);
    F3_TAKE(odw, kvl); F3_TAKE(osval, kvl);
    F3_TAKE(oiwb, kpoc);
    F3_TAKE(ocs, F2_CS * ns1);
    F3_TAKE(oscr, 2 * 8 * nw);
    F3_TAKE(otab, (k + n + 1 + nsing + 1) / 2);                          // ints: srow_col[k] | scol_ptr[n+1] | scol_rows[nsing]
    F3_TAKE(otij, (ntl + 3) / 4); F3_TAKE(otijm, p > 8 ? (ntlm + 3) / 4 : 0);   // unsigned shorts
#undef F3_TAKE
    O.total = at;
    return O;
}

struct F3Plan : F3Offs {
    bool fits = false;
    int nw = 4;
    int ctas_per_sm = 1, num_sms = 148;
    bool teams4 = false;             // four teams (problems) fit into one CTA's shared memory
    size_t smem = 0;                 // per team
    int* d_counter = nullptr;
    unsigned long long* d_clk = nullptr;   // 16 phase-cycle counters (SOCP_PHASE_TIMING builds, tools/phase_timing.py)
    const int* d_tables = nullptr;   // device copy of `tables` (srow_col[k] | scol_ptr[n+1] | scol_rows[nsing])
    void* jit_fn = nullptr;          // the kernel specialised for this layout at run time (lane_jit.cu), if any,
    int jit_teams = 0;               // and the number of teams per CTA it was built for
    int n = 0, p = 0, k = 0, kpoc = 0, nsoc = 0, lpc = 1;
    int npad = 0, nb = 0, ntl = 0;
    int d0 = 0, kd = 0, kdpad = 0, ldg = 0, kshift = 0, kvl = 0;
    int nsing = 0, ident = 0;        // ident: the singleton rows are exactly rows 0..n-1, row i <-> column i
    int ppad = 0, pb = 0, ntlm = 0;
    int split_kd = 1, split_p = 1;
    int shape = 0;
    int soc_offs[F2_MAX_SOC], soc_dim[F2_MAX_SOC];
};

#ifndef __CUDACC_RTC__
// rowcol[i]: -1 = row i of G is empty in every problem, j >= 0 = its only nonzeros sit in column j, -2 = dense.
// `tables` receives the device tables.  No CUDA calls (the simulator tests build plans too).
inline void f3_plan(F3Plan& P, int n, int p, int k, const std::vector<int>& kind, const std::vector<int>& offs,
                    const std::vector<int>& dim, const std::vector<int>& rowcol, int dev_smem_optin, int num_sms,
                    std::vector<int>& tables) {
    P.fits = false;
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
    if (maxd > 128 || n > 64 || n <= 16 || p > 32 || (int)rowcol.size() != k) return;
    P.lpc = f2_lpc(maxd);
    P.nw = 4;
    P.npad = (n + 7) / 8 * 8; P.nb = P.npad / 8; P.ntl = P.nb * (P.nb + 1) / 2;
    P.ppad = (std::max(p, 1) + 7) / 8 * 8; P.pb = P.ppad / 8; P.ntlm = P.pb * (P.pb + 1) / 2;
    // dense block: from the first to the last row with two or more nonzeros
    int lo = k, hi = -1;
    for (int i = 0; i < k; ++i)
        if (rowcol[i] == -2) { lo = std::min(lo, i); hi = std::max(hi, i); }
    if (hi < 0) { lo = 0; hi = -1; }
    P.d0 = lo; P.kd = hi - lo + 1;
    P.kdpad = (P.kd + 3) / 4 * 4;
    P.ldg = f2_ld(std::max(P.kdpad, 4));
    P.kshift = P.d0 & 1;
    P.kvl = (k + 5 + 1) / 2 * 2;
    // singleton tables
    std::vector<int> srow(k, -1), cptr(n + 1, 0), crow;
    for (int i = 0; i < k; ++i)
        if ((i < P.d0 || i >= P.d0 + P.kd) && rowcol[i] >= 0 && rowcol[i] < n) srow[i] = rowcol[i];
    for (int c = 0; c < n; ++c) {
        cptr[c] = (int)crow.size();
        for (int i = 0; i < k; ++i)
            if (srow[i] == c) crow.push_back(i);
    }
    cptr[n] = (int)crow.size();
    P.nsing = (int)crow.size();
    P.ident = (P.nsing == n) ? 1 : 0;
    for (int c = 0; c < n && P.ident; ++c) P.ident = (srow[c] == c);
    tables.clear();
    tables.insert(tables.end(), srow.begin(), srow.end());
    tables.insert(tables.end(), cptr.begin(), cptr.end());
    tables.insert(tables.end(), crow.begin(), crow.end());
    P.split_kd = f2_split((std::max(P.kd, 1) + 1) / 2, P.nw);
    P.split_p = f2_split(std::max(p, 1), P.nw);
    static_cast<F3Offs&>(P) = f3_offsets(n, p, k, P.kpoc, P.nsoc, P.kd, P.nsing, P.nw);
    P.smem = (size_t)P.total * sizeof(double);
    P.num_sms = num_sms;
    if (P.smem + 512 > (size_t)dev_smem_optin) return;
    const int per_sm = 228 * 1024;
    P.ctas_per_sm = std::max(1, std::min({(int)(per_sm / (P.smem + 1024)), 2048 / (P.nw * 32), 4}));
    P.teams4 = 4 * P.smem + 256 <= (size_t)dev_smem_optin;
    P.fits = true;
}

#endif  // __CUDACC_RTC__

// global arrays of one shard as the kernel sees them (filled from Ws by the host; the simulator fills it directly)
struct F3Glob {
    const double *c, *A, *b, *G, *h;
    int64_t sA, sG;
    const uint8_t* sing;      // may be null: see sing_detect
    uint8_t* sing_out;        // may be null; receives 1 for problems found `sing` by the kernel
    double *x, *y, *z, *s, *pobj, *dobj;
    int *status, *iters, *active, *fail;
    int deg;
    int* npattern;            // counts problems reported ST_PATTERN (may be null when verify == 0)
    double* dbg;              // debug dump of one step (may be null), see f3_dbg_size
    int dbg_prob, dbg_iter, dbg_phase;   // dbg_phase: 1 = affine solve, 2 = combined solve
};
// layout of the debug dump: s, z (inputs of compute_scaling) | H (n x n, column-major, before the factorisation) |
// dx, dy, dz, ds (right-hand side of the chosen solve_kkt, scaled as the solver passes them) | cx, cy, cz, cs (its result)
__host__ __device__ inline int f3_dbg_size(int n, int p, int k) { return 2 * k + n * n + 2 * (n + p + 2 * k); }

struct F3Args {
    F3Glob g;
    F3Plan P;
    LoopParams prm;
    int first, batch;
    int* counter;
    int sing_detect;          // 1: `sing` is unknown -- a failing factorisation of G'G switches the problem to sing
    int verify;               // 1: check every entry of G outside the pattern while loading
    int align;                // TEAMS > 1: re-align the teams of a CTA at the top of every iteration
};

// ------------------------------------------------------------------------------------------------ layout providers
struct Dims3Dyn {
    static constexpr bool is_static = false;
    __device__ __forceinline__ static int n(const F3Plan& P) { return P.n; }
    __device__ __forceinline__ static int p(const F3Plan& P) { return P.p; }
    __device__ __forceinline__ static int k(const F3Plan& P) { return P.k; }
    __device__ __forceinline__ static int kpoc(const F3Plan& P) { return P.kpoc; }
    __device__ __forceinline__ static int nsoc(const F3Plan& P) { return P.nsoc; }
    __device__ __forceinline__ static int lpc(const F3Plan& P) { return P.lpc; }
    __device__ __forceinline__ static int d0(const F3Plan& P) { return P.d0; }
    __device__ __forceinline__ static int kd(const F3Plan& P) { return P.kd; }
    __device__ __forceinline__ static int ident(const F3Plan& P) { return P.ident; }
    __device__ __forceinline__ static int split_kd(const F3Plan& P) { return P.split_kd; }
    __device__ __forceinline__ static int split_p(const F3Plan& P) { return P.split_p; }
    __device__ __forceinline__ static int soc_offs(const F3Plan& P, int slot) { return P.soc_offs[slot]; }
    __device__ __forceinline__ static int soc_dim(const F3Plan& P, int slot) { return P.soc_dim[slot]; }
    __device__ __forceinline__ static const F3Offs& offs(const F3Plan& P) { return P; }
    __device__ __forceinline__ static int nsing(const F3Plan& P) { return P.nsing; }
};
// N variables, PE equality rows, one POC block of KPOC rows then NSOC cones of dimension SDIM; dense rows [D0, D0+KD),
// identity singleton block (IDENT) or the generic tables
template <int NW, int N, int PE, int KPOC, int NSOC, int SDIM, int D0, int KD, int IDENT>
struct Dims3Static {
    static constexpr bool is_static = true;
    static constexpr int K = KPOC + NSOC * SDIM;
    __device__ __forceinline__ static constexpr int n(const F3Plan&) { return N; }
    __device__ __forceinline__ static constexpr int p(const F3Plan&) { return PE; }
    __device__ __forceinline__ static constexpr int k(const F3Plan&) { return K; }
    __device__ __forceinline__ static constexpr int kpoc(const F3Plan&) { return KPOC; }
    __device__ __forceinline__ static constexpr int nsoc(const F3Plan&) { return NSOC; }
    __device__ __forceinline__ static constexpr int lpc(const F3Plan&) { return f2_lpc(SDIM); }
    __device__ __forceinline__ static constexpr int d0(const F3Plan&) { return D0; }
    __device__ __forceinline__ static constexpr int kd(const F3Plan&) { return KD; }
    __device__ __forceinline__ static constexpr int ident(const F3Plan&) { return IDENT; }
    __device__ __forceinline__ static constexpr int split_kd(const F3Plan&) { return f2_split(((KD > 0 ? KD : 1) + 1) / 2, NW); }
    __device__ __forceinline__ static constexpr int split_p(const F3Plan&) { return f2_split(PE > 0 ? PE : 1, NW); }
    __device__ __forceinline__ static constexpr int soc_offs(const F3Plan&, int slot) { return KPOC + slot * SDIM; }
    __device__ __forceinline__ static constexpr int soc_dim(const F3Plan&, int) { return SDIM; }
    static constexpr int NSING = IDENT ? N : 0;          // static layouts: identity singleton block or none
    __device__ __forceinline__ static constexpr F3Offs offs(const F3Plan&) { return f3_offsets(N, PE, K, KPOC, NSOC, KD, NSING, NW); }
    __device__ __forceinline__ static constexpr int nsing(const F3Plan&) { return NSING; }
    static bool matches(const F3Plan& P) {
        if (P.n != N || P.p != PE || P.kpoc != KPOC || P.nsoc != NSOC || P.k != K || P.nw != NW) return false;
        if (P.d0 != D0 || P.kd != KD || P.ident != IDENT || P.nsing != NSING) return false;
        for (int i = 0; i < NSOC; ++i)
            if (P.soc_dim[i] != SDIM || P.soc_offs[i] != KPOC + i * SDIM) return false;
        return true;
    }
};

// ------------------------------------------------------------------------------------------------ tile kernels
// the tiles warp `warp` owns in the products over the lower triangle: flat indices warp, warp + NW, ...
template <int NW, int MAXT>
__device__ __forceinline__ void f3_my_tiles(int nb, int warp, int (&tl)[MAXT]) {
    const int ntl = nb * (nb + 1) / 2;
#pragma unroll
    for (int q = 0; q < MAXT; ++q) {
        const int t = min(warp + q * NW, ntl - 1);
        int a = 0;
        while ((a + 1) * (a + 2) / 2 <= t) ++a;
        tl[q] = a | ((t - a * (a + 1) / 2) << 8);
    }
}

// Tt (packed lower tiles, full diagonal tiles; n x n padded to 8 nb with a unit pad diagonal)
//   = G_D' diag(dwD) G_D + sum_c hc[c] hc[c]' + diag(sd) (+ A'A when sing)
// i.e. G'W^-2 G of src/densesolver.jl:42-43 (+ AA of :44-46) on the compressed G: G_D = the dense rows (kdpad x n, ld
// ldg, pad rows zero), dwD their weights, sd[j] = sum over the singleton rows (j, g) of dw g^2.
template <int NW, int MAXT>
__device__ __forceinline__ void f3_syrk(const double* G, int ldg, int kdpad, int n, int nb, const double* dwD,
                                        const double* hc, int nsoc, int npad, const double* sd, const double* A, int p,
                                        bool sing, double* Tt, const int (&tl)[MAXT], int lane, int warp, const T3Lane& TL) {
    const int ntl = nb * (nb + 1) / 2;
    const int fr = lane >> 2, fk = lane & 3;
    const double* pa[MAXT];
    const double* pb[MAXT];
    double acc[MAXT][2];
#pragma unroll
    for (int q = 0; q < MAXT; ++q) {
        const int ti = tl[q] & 255, tj = tl[q] >> 8;
        pa[q] = G + min(ti * 8 + fr, n - 1) * ldg + fk;      // pad columns read a valid column; masked below
        pb[q] = G + min(tj * 8 + fr, n - 1) * ldg + fk;
        acc[q][0] = acc[q][1] = 0.0;
    }
    const double* pw = dwD + fk;
#pragma unroll 2
    for (int kk = 0; kk < kdpad; kk += 4) {
        const double w = pw[kk];
#pragma unroll
        for (int q = 0; q < MAXT; ++q)
            if (warp + q * NW < ntl) dmma884(acc[q][0], acc[q][1], pa[q][kk] * w, pb[q][kk]);
    }
#pragma unroll
    for (int q = 0; q < MAXT; ++q) {
        if (warp + q * NW < ntl) {
            const int ti = tl[q] & 255, tj = tl[q] >> 8;
            const int gi = ti * 8 + fr;
            double* T = Tt + (warp + q * NW) * 64;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gj = tj * 8 + 2 * fk + e;
                double v = acc[q][e];
                if (gi < n && gj < n) {
                    for (int c = 0; c < nsoc; ++c) v = fma(hc[c * npad + gi], hc[c * npad + gj], v);
                    if (gi == gj) v += sd[gi];
                    if (sing)
                        for (int r = 0; r < p; ++r) v = fma(A[gi * p + r], A[gj * p + r], v);      // AA, :32,:44-46
                } else {
                    v = (gi == gj) ? 1.0 : 0.0;
                }
                T[e ? TL.d1 : TL.d0] = v;
            }
        }
    }
}

// register-resident factor of one packed diagonal tile (see f2_diag_factor): on exit xc[i] = (L^-1)[i][lane & 7]
__device__ __forceinline__ int f3_diag_factor(const double* T, int lane, double (&xc)[8]) {
    double a[36];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) a[i * (i + 1) / 2 + j] = T[t3_off(i, j)];
    double r[8];
    int ok = 1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double d = a[j * (j + 1) / 2 + j];
        ok &= (d > 0.0);
        const double rj = fast_rsqrt(d);
        r[j] = rj;
#pragma unroll
        for (int i = j + 1; i < 8; ++i) a[i * (i + 1) / 2 + j] *= rj;
#pragma unroll
        for (int i = j + 1; i < 8; ++i)
#pragma unroll
            for (int c = j + 1; c <= i; ++c)
                a[i * (i + 1) / 2 + c] = fma(-a[i * (i + 1) / 2 + j], a[c * (c + 1) / 2 + j], a[i * (i + 1) / 2 + c]);
    }
    const int c = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double sacc = 0.0;
#pragma unroll
        for (int m = 0; m < i; ++m) sacc = fma(a[i * (i + 1) / 2 + m], xc[m], sacc);
        xc[i] = (i == c) ? r[i] : -r[i] * sacc;      // i < c: sacc == 0
    }
    return ok;
}


// In-place blocked Cholesky that carries the inverse along.  Tt: nbl x nbl packed lower tiles (unit pad diagonal).
// On exit every tile holds the corresponding tile of X = L^-1 (diagonal tiles lower triangular, zeros above).
// Lp: nbl-1 tiles of scratch (the panel L_ib of the current block column).  tij[t] = i | j << 8 of packed tile t.
// Step b:  (warp 0) X_bb = L_bb^-1 from the register factor of tile (b,b)           | barrier 1
//          panel: L_ib = H_ib X_bb' -> Lp (i > b);  X_bc <- X_bb X_bc (c < b)        | barrier 2 (warp 0 only arrives)
//          trailing, rows i > b: H_ij -= L_ib L_jb' (j > b);  X_ij -= L_ib X_bj (j <= b; j == b starts from zero and
//          overwrites the dead H_ib).  Warp 0 takes tile (b+1, b+1) only and runs ahead into its factorisation.
// fail[b & 1] is set (and 0 returned) on a non-positive pivot of step b: cholesky!'s PosDefException,
// src/densesolver.jl:47,51.  Two slots because warp 0 runs one step ahead of the warps that read the flag; the caller
// clears both.
template <int NW>
__device__ __forceinline__ int f3_chol_inv(double* Tt, double* Lp, const unsigned short* tij, int nbl, int* fail,
                                           int lane, int warp, const T3Lane& TL, unsigned long long* clk = nullptr,
                                           int bar = 0, int bar2 = 1) {
    const int ntl = nbl * (nbl + 1) / 2;
#ifdef SOCP_PHASE_TIMING
    const bool pt_on = clk && threadIdx.x == 0 && blockIdx.x == 0;
    long long pt_c = clock64();
#endif
    for (int b = 0; b < nbl; ++b) {
        double* Db = Tt + t3_idx(b, b) * 64;
#ifdef SOCP_PHASE_TIMING
        if (pt_on) pt_c = clock64();
#endif
        if (warp == 0) {
            __syncwarp();
            double xc[8];
            const int ok = f3_diag_factor(Db, lane, xc);
            if (!ok && lane == 0) fail[b & 1] = 1;
            __syncwarp();                              // every lane has read the tile before it is overwritten
            if (lane < 8) {
                const int cb = 4 * t3_g(lane) + 16 * (lane >> 2);
#pragma unroll
                for (int i = 0; i < 8; ++i) Db[(i & 3) + 32 * (i >> 2) + cb] = xc[i];
            }
        }
#ifdef SOCP_PHASE_TIMING
        if (pt_on) { const long long t = clock64(); atomicAdd(&clk[13], (unsigned long long)(t - pt_c)); pt_c = t; }
#endif
        tsync_t<NW>(bar);                                   // (1) X_bb visible; trailing update of step b-1 complete
#ifdef SOCP_PHASE_TIMING
        if (pt_on) { const long long t = clock64(); atomicAdd(&clk[15], (unsigned long long)(t - pt_c)); pt_c = t; }
#endif
        if (fail[b & 1]) return 0;
        // ---- panel
        {
            const double d0 = Db[TL.a0], d1 = Db[TL.a1];          // X_bb as A fragment == X_bb' as B fragment
            const int nl = nbl - 1 - b;                           // L tiles; then b tiles of row b of X
            for (int t = (NW > 1) ? (warp == 0 ? 0 : warp) : 0; t < nbl - 1; t += (NW > 1) ? (warp == 0 ? nbl : NW - 1) : 1) {
                double c0 = 0.0, c1 = 0.0;
                if (t < nl) {
                    const double* Hs = Tt + t3_idx(b + 1 + t, b) * 64;
                    dmma884(c0, c1, Hs[TL.a0], d0);
                    dmma884(c0, c1, Hs[TL.a1], d1);
                    Lp[t * 64 + TL.d0] = c0;
                    Lp[t * 64 + TL.d1] = c1;
                } else {
                    double* Xs = Tt + t3_idx(b, t - nl) * 64;
                    const double x0 = Xs[TL.c0], x1 = Xs[TL.c1];
                    dmma884(c0, c1, d0, x0);
                    dmma884(c0, c1, d1, x1);
                    Xs[TL.d0] = c0;
                    Xs[TL.d1] = c1;
                }
            }
        }
        // (2) panel visible.  Warp 0 needs only the panel tile it wrote itself (row b+1), so it merely arrives.
        if (NW > 1) {
            if (warp == 0) { bar_arrive_id<NW * 32>(bar2); __syncwarp(); }
            else bar_sync_id<NW * 32>(bar2);
        } else __syncwarp();
        // ---- trailing update
        if (b + 1 < nbl) {
            const int T0 = (b + 1) * (b + 2) / 2;
            const int tw = T0 + b + 1;                            // tile (b+1, b+1)
            const double* Xrow = Tt + t3_idx(b, 0) * 64;
            auto one = [&](int t, bool valid, double& a0, double& a1, double& q0, double& q1, double& c0, double& c1) {
                const int ij = tij[valid ? t : T0];
                const int i = ij & 255, j = ij >> 8;
                const double* As = Lp + (i - b - 1) * 64;
                a0 = -As[TL.a0]; a1 = -As[TL.a1];
                if (j > b) {
                    const double* Bs = Lp + (j - b - 1) * 64;
                    q0 = Bs[TL.a0]; q1 = Bs[TL.a1];
                } else {
                    const double* Bs = Xrow + j * 64;
                    q0 = Bs[TL.c0]; q1 = Bs[TL.c1];
                }
                const double* C = Tt + (valid ? t : T0) * 64;
                const bool zero = (j == b);
                c0 = zero ? 0.0 : C[TL.d0];
                c1 = zero ? 0.0 : C[TL.d1];
            };
            if (NW > 1 && warp == 0) {
                double a0, a1, q0, q1, c0, c1;
                one(tw, true, a0, a1, q0, q1, c0, c1);
                dmma884(c0, c1, a0, q0);
                dmma884(c0, c1, a1, q1);
                Tt[tw * 64 + TL.d0] = c0;
                Tt[tw * 64 + TL.d1] = c1;
            } else {
                const int stride = (NW > 1) ? NW - 1 : 1;
                const int nrest = ntl - T0 - ((NW > 1) ? 1 : 0);
                for (int e = (NW > 1) ? warp - 1 : 0; e < nrest; e += 2 * stride) {
                    const int e2 = e + stride;
                    const bool v2 = e2 < nrest;
                    int t1 = T0 + e, t2 = T0 + e2;
                    if (NW > 1) { t1 += (t1 >= tw); t2 += (t2 >= tw); }
                    double a0, a1, q0, q1, c0, c1, A0, A1, Q0, Q1, C0, C1;
                    one(t1, true, a0, a1, q0, q1, c0, c1);
                    one(t2, v2, A0, A1, Q0, Q1, C0, C1);
                    dmma884(c0, c1, a0, q0);
                    dmma884(C0, C1, A0, Q0);
                    dmma884(c0, c1, a1, q1);
                    dmma884(C0, C1, A1, Q1);
                    Tt[t1 * 64 + TL.d0] = c0;
                    Tt[t1 * 64 + TL.d1] = c1;
                    if (v2) {
                        Tt[t2 * 64 + TL.d0] = C0;
                        Tt[t2 * 64 + TL.d1] = C1;
                    }
                }
            }
        }
    }
    tsync_t<NW>(bar);
    return 1;
}

// Tt <- X'X in place (X = L^-1 as left by f3_chol_inv): H^-1 = L^-T L^-1, the reference's Li (src/densesolver.jl:48).
// Out(i,j) = sum_{m >= i} X(m,i)' X(m,j).  Every warp keeps its tiles in registers until all reads are done.
template <int NW, int MAXT>
__device__ __forceinline__ void f3_xtx(double* Tt, int nb, const int (&tl)[MAXT], int lane, int warp, const T3Lane& TL,
                                       int bar = 0) {
    const int ntl = nb * (nb + 1) / 2;
    double acc[MAXT][2];
#pragma unroll
    for (int q = 0; q < MAXT; ++q) {
        acc[q][0] = acc[q][1] = 0.0;
        if (warp + q * NW < ntl) {
            const int ti = tl[q] & 255, tj = tl[q] >> 8;
            for (int m = ti; m < nb; ++m) {
                const double* Xa = Tt + t3_idx(m, ti) * 64;
                const double* Xb = Tt + t3_idx(m, tj) * 64;
                dmma884(acc[q][0], acc[q][1], Xa[TL.c0], Xb[TL.c0]);
                dmma884(acc[q][0], acc[q][1], Xa[TL.c1], Xb[TL.c1]);
            }
        }
    }
    tsync_t<NW>(bar);
#pragma unroll
    for (int q = 0; q < MAXT; ++q) {
        if (warp + q * NW < ntl) {
            double* T = Tt + (warp + q * NW) * 64;
            T[TL.d0] = acc[q][0];
            T[TL.d1] = acc[q][1];
        }
    }
}

// y = S x for the symmetric S in packed lower tiles (full diagonal tiles); x has 8 nb entries (pads finite).
// One block row per warp pass; lane (fr, fk) accumulates columns fk, fk + 4 of every tile.  epi(row, value), row < 8 nb.
template <int NW, class Epi>
__device__ __forceinline__ void f3_symv(const double* Tt, int nb, const double* x, int lane, int warp, const T3Lane& TL,
                                        Epi epi) {
    const int fr = lane >> 2, fk = lane & 3;
    // nb <= 2 NW: a warp owns at most two block rows; their chains are interleaved (a missing second row re-does the
    // first and drops the result)
    for (int i0 = warp; i0 < nb; i0 += 2 * NW) {
        const int i1 = i0 + NW;
        const bool two = i1 < nb;
        const int ib = two ? i1 : i0;
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
        const double* Ta = Tt + t3_idx(i0, 0) * 64;
        const double* Tb = Tt + t3_idx(ib, 0) * 64;
        for (int j = 0; j < nb; ++j) {
            const double x0 = x[8 * j + fk], x1 = x[8 * j + fk + 4];
            const double* pa = (j <= i0) ? Ta + j * 64 : Tt + t3_idx(j, i0) * 64;
            const double* pb = (j <= ib) ? Tb + j * 64 : Tt + t3_idx(j, ib) * 64;
            const int oa0 = (j <= i0) ? TL.a0 : TL.c0, oa1 = (j <= i0) ? TL.a1 : TL.c1;
            const int ob0 = (j <= ib) ? TL.a0 : TL.c0, ob1 = (j <= ib) ? TL.a1 : TL.c1;
            a0 = fma(pa[oa0], x0, a0);
            a1 = fma(pa[oa1], x1, a1);
            b0 = fma(pb[ob0], x0, b0);
            b1 = fma(pb[ob1], x1, b1);
        }
        double acc = a0 + a1, bcc = b0 + b1;
        acc += __shfl_xor_sync(FULL_MASK, acc, 1);
        bcc += __shfl_xor_sync(FULL_MASK, bcc, 1);
        acc += __shfl_xor_sync(FULL_MASK, acc, 2);
        bcc += __shfl_xor_sync(FULL_MASK, bcc, 2);
        if (fk == 0) {
            epi(8 * i0 + fr, acc);
            if (two) epi(8 * i1 + fr, bcc);
        }
    }
}

#ifdef SOCP_PHASE_TIMING
#define PT3_DECL() long long pt_t0 = 0
#define PT3_INIT() pt_t0 = clock64()
#define PT3_MARK(idx)                                                             \
    do {                                                                          \
        if (tid_all == 0 && blockIdx.x == 0) {                                    \
            const long long t_ = clock64();                                       \
            atomicAdd(&a.P.d_clk[idx], (unsigned long long)(t_ - pt_t0));      \
            pt_t0 = t_;                                                           \
        }                                                                         \
    } while (0)
#else
#define PT3_DECL()
#define PT3_INIT()
#define PT3_MARK(idx)
#endif
enum { P3_LOAD = 0, P3_RESID, P3_HEAD_GT, P3_SYRK, P3_XTX, P3_EQ, P3_SOLVE, P3_INIT, P3_TAIL, P3_MIDPOST, P3_OUT,
       P3_CHOL, P3_SOLVE_A, P3_SOLVE_B, P3_N0 };

// ------------------------------------------------------------------------------------------------ kernel
// TEAMS teams of NW warps per CTA, every team on its own problem with its own slice of the dynamic shared memory and
// its own pair of hardware barriers (one 512-thread CTA per SM instead of four 128-thread CTAs: the kernel then knows
// which SM sub-partition each of its warps issues from, see the warp rotation below).  a.align (experiments): the
// teams re-align at the top of every Mehrotra iteration with two CTA-wide barriers.
template <int NW, int TEAMS, int MAXT, int MINB, class D>
__global__ void __launch_bounds__(NW * 32 * TEAMS, MINB) k_fused3(const F3Args a) {
#ifdef SOCP_SIMT_EMU
    double* sm_all = reinterpret_cast<double*>(emu_dyn_smem());
    const int tid_all = (int)(unsigned)threadIdx.x;
#else
    extern __shared__ __align__(16) double sm_all[];
    int tid_all;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_all));
#endif
    __shared__ int s_prob_[TEAMS], s_fail_[TEAMS][2], s_bad_[TEAMS], s_live;
    const int team = (TEAMS > 1) ? tid_all / (NW * 32) : 0;
    // warp w of the CTA issues from SM sub-partition w % 4.  The serial stretches of the solve (diagonal-tile factor,
    // cone chains of a single cone) run on the team's warp 0: rotate the warp numbering by the team index so that the
    // four teams' chain warps sit on four different sub-partitions.
    const int tid = (TEAMS > 1) ? ((tid_all + team * 32) & (NW * 32 - 1)) : tid_all;
    int& s_prob = s_prob_[team];
    int* const s_fail = s_fail_[team];
    int& s_bad = s_bad_[team];
    const int bar = 1 + team, bar2 = 1 + TEAMS + team;       // hardware barriers of this team (0 is the CTA's)
    double* sm = sm_all + (size_t)team * D::offs(a.P).total;
    const F3Plan& P = a.P;
    const F3Offs O = D::offs(P);
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int T = NW * 32;
    const int n = D::n(P), p = D::p(P), k = D::k(P), kpoc = D::kpoc(P), nsoc = D::nsoc(P), lpc = D::lpc(P);
    const int npad = (n + 7) / 8 * 8, nb = npad / 8, ntl = nb * (nb + 1) / 2;
    const int d0 = D::d0(P), kd = D::kd(P), kdpad = (kd + 3) / 4 * 4, ldg = f2_ld(max(kdpad, 4));
    const int ksh = d0 & 1;
    const bool ident = D::ident(P) != 0;
    const int ppad = (max(p, 1) + 7) / 8 * 8, pb = ppad / 8;
    const T3Lane TL = t3_lane(lane);
    double* G = sm + O.oG;           // dense rows [d0, d0 + kd) of G: kdpad x n, ld ldg
    double* Tt = sm + O.oT;          // G'W^-2 G, then L^-1, then H^-1
    double* Lp = sm + O.oLp;
    double* A = sm + O.oA;
    double* At = sm + O.oAt;         // A' (rows of A contiguous, stride npad)
    double* HiAt = sm + O.oHiAt;     // H^-1 A'         (p x npad)
    double* Km = sm + O.oK;          // H^-1 A' M^-1    (p x npad)
    double* Mt = sm + O.oMt;
    double* MLp = sm + O.oMLp;
    double* Minv = sm + O.oMinv;
    double* cv = sm + O.oc;
    double* bv = sm + O.ob;
    double* x = sm + O.ox; double* y = sm + O.oy;
    double* dx = sm + O.odx; double* dy = sm + O.ody;
    double* n0 = sm + O.on0; double* cx = sm + O.ocx; double* cy = sm + O.ocy;
    double* kd_ = sm + O.okd;        // K dy (+ H^-1 A' dy for sing problems inside the loop)   (n)
    double* md = sm + O.omd;         // M^-1 dy         (p)
    double* sd = sm + O.osd;         // singleton rows' share of diag(G'W^-2 G)                 (n)
    double* atdy = sm + O.oatdy;     // A'dy            (n)
    double* hc = sm + O.ohc;         // sqrt(2)/eta G_c'q per cone (nsoc x npad)
    // k-vectors start at an odd offset when d0 is odd, so that their dense-row part is 16-byte aligned
    double* hv = sm + O.oh + ksh; double* z = sm + O.oz + ksh; double* s = sm + O.os + ksh;
    double* lam = sm + O.olam + ksh; double* wb = sm + O.owb + ksh; double* iwb = sm + O.oiwb;
    double* dz = sm + O.odz + ksh; double* ds = sm + O.ods + ksh;
    double* k0 = sm + O.ok0 + ksh; double* k2 = sm + O.ok2 + ksh; double* u = sm + O.ou + ksh;
    double* dw = sm + O.odw + ksh;   // row weights of the SYRK
    double* sval = sm + O.osval + ksh;   // value of the only nonzero of a singleton row (0 elsewhere)
    double* cs = sm + O.ocs;
    double* scr = sm + O.oscr;
    const int* srow_col = reinterpret_cast<const int*>(sm + O.otab);
    const int* scol_ptr = srow_col + k;
    const int* scol_rows = scol_ptr + n + 1;
    unsigned short* tij = reinterpret_cast<unsigned short*>(sm + O.otij);
    unsigned short* tijm = reinterpret_cast<unsigned short*>(sm + O.otijm);
    int scr_par = 0;
    const LoopParams prm = a.prm;
    PT3_DECL();

    for (int q = tid; q < O.total; q += T) sm[q] = 0.0;
    tsync_t<NW>(bar);
    {
        int* tab = reinterpret_cast<int*>(sm + O.otab);
        const int ntab = k + n + 1 + D::nsing(P);
        for (int q = tid; q < ntab; q += T) tab[q] = a.P.d_tables[q];
        for (int t = tid; t < ntl; t += T) {
            int i = 0;
            while ((i + 1) * (i + 2) / 2 <= t) ++i;
            tij[t] = (unsigned short)(i | ((t - i * (i + 1) / 2) << 8));
        }
        if (p > 8)
            for (int t = tid; t < pb * (pb + 1) / 2; t += T) {
                int i = 0;
                while ((i + 1) * (i + 2) / 2 <= t) ++i;
                tijm[t] = (unsigned short)(i | ((t - i * (i + 1) / 2) << 8));
            }
    }
    int tl[MAXT];
    f3_my_tiles<NW, MAXT>(nb, warp, tl);
    if (TEAMS > 1) {
        if (tid_all == 0) s_live = TEAMS;
        __syncthreads();
    }
    tsync_t<NW>(bar);

    // contribution of the singleton rows to column c of G'v
    auto sing_col = [&](int c, const double* v) -> double {
        if (ident) return sval[c] * v[c];
        double acc = 0.0;
        for (int e = scol_ptr[c]; e < scol_ptr[c + 1]; ++e) { const int i = scol_rows[e]; acc = fma(sval[i], v[i], acc); }
        return acc;
    };
    // (G v)_i of a row outside the dense block
    auto sing_row = [&](int i, const double* v) -> double {
        const int c = srow_col[i];
        return c >= 0 ? sval[i] * v[c] : 0.0;
    };

    // ------------------------------------------------------------------ cone chains (identical to fused_v2.cuh)
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

    // work distribution: the first problem of a team is fixed (a small batch spreads over the SMs one problem per CTA
    // before any CTA gets a second one), further ones come from the shard's atomic counter
    bool first_round = true;
    for (;;) {
        int b;
        if (first_round) {
            b = a.first + (int)blockIdx.x + team * (int)gridDim.x;
            first_round = false;
        } else {
            if (tid == 0) s_prob = atomicAdd(a.counter, 1) + (int)gridDim.x * TEAMS;
            tsync_t<NW>(bar);
            b = s_prob + a.first;
            tsync_t<NW>(bar);
        }
        if (b >= a.first + a.batch) {
            if (TEAMS > 1 && a.align) {
                // out of work: keep the CTA-wide alignment barriers of the other teams company until every team is
                // here.  Pairs of barriers: s_live changes only between pairs and is read only inside one.
                if (tid == 0) atomicAdd(&s_live, -1);
                for (;;) {
                    __syncthreads();
                    const int live = *(volatile int*)&s_live;
                    __syncthreads();
                    if (live == 0) break;
                }
            }
            break;
        }
        PT3_INIT();

        // ---- load the problem (global -> shared)
        int bad = 0;
        {
            const double* Gg = a.g.G + (int64_t)b * a.g.sG;
            for (int col = warp; col < n; col += NW)
                for (int r = lane; r < kd; r += 32) G[col * ldg + r] = Gg[(int64_t)col * k + d0 + r];
            for (int i = tid; i < k; i += T) {
                const int c = srow_col[i];
                sval[i] = c >= 0 ? Gg[(int64_t)c * k + i] : 0.0;
            }
            if (a.verify) {
                for (int col = warp; col < n; col += NW)
                    for (int r = lane; r < k; r += 32)
                        if (r < d0 || r >= d0 + kd) bad |= (Gg[(int64_t)col * k + r] != 0.0) & (srow_col[r] != col);
            }
            const double* Ag = a.g.A + (int64_t)b * a.g.sA;
            for (int q = tid; q < p * n; q += T) { const double v = Ag[q]; A[q] = v; At[(q % p) * npad + q / p] = v; }
            for (int i = tid; i < n; i += T) cv[i] = a.g.c[(int64_t)b * n + i];
            for (int i = tid; i < p; i += T) bv[i] = a.g.b[(int64_t)b * p + i];
            for (int i = tid; i < k; i += T) hv[i] = a.g.h[(int64_t)b * k + i];
        }
        int status = ST_RUNNING, iters = 0;
        int sing = (a.g.sing && !a.sing_detect) ? (a.g.sing[b] != 0) : 0;
        if (a.verify) {
            if (tid == 0) s_bad = 0;
            tsync_t<NW>(bar);
            if (bad) s_bad = 1;
            tsync_t<NW>(bar);
            if (s_bad) {
                if (tid == 0) {
                    a.g.status[b] = ST_PATTERN;
                    a.g.iters[b] = 0;
                    if (a.g.npattern) atomicAdd(a.g.npattern, 1);
                }
                tsync_t<NW>(bar);
                continue;
            }
        }
        tsync_t<NW>(bar);

        // The initial point (src/solver.jl:68-104, W = I) is the same factor + solve as a loop iteration with
        // u = h, dx = -c, dy = b, k2 = h: then cx = x, cy = y and u = G x - h = z0.
        for (int i = tid; i < k; i += T) { u[i] = hv[i]; k2[i] = hv[i]; dw[i] = 1.0; }
        for (int q = tid; q < nsoc * npad; q += T) hc[q] = 0.0;
        for (int i = tid; i < n; i += T) {
            dx[i] = -cv[i];
            double acc = 0.0, sdv = 0.0;
            for (int q = 0; q < p; ++q) acc = fma(A[i * p + q], bv[q], acc);
            if (p > 0) atdy[i] = acc;
            if (ident) sdv = sval[i] * sval[i];
            else
                for (int e = scol_ptr[i]; e < scol_ptr[i + 1]; ++e) { const double g = sval[scol_rows[e]]; sdv = fma(g, g, sdv); }
            sd[i] = sdv;
        }
        for (int i = tid; i < p; i += T) dy[i] = bv[i];
        tsync_t<NW>(bar);
        PT3_MARK(P3_LOAD);

        int phase = 0;              // 0: initial point, 1: affine direction (solve #1), 2: combined direction (solve #2)
        double sc = 1.0;            // (1 - sigma) applied to dx, dy, dz in solve #2 (src/solver.jl:140)
        double ll = 0.0;            // lambda'lambda of the current iteration
        for (;;) {
            // ---- n0 = G'u + sc*dx (+ sc*A'dy: sing, src/densesolver.jl:68-71)            src/densesolver.jl:66-67
            gemv_cols_vu<NW, 16 / NW>(G, ldg, kd, n, u + d0, lane, warp, [&](int c, double acc) {
                double r = dx[c];
                if (sing) r += atdy[c];
                n0[c] = acc + sing_col(c, u) + sc * r;
            });
            PT3_MARK(P3_N0);
            if (phase != 2) {
                // ---- KKT factor, src/densesolver.jl:41-52
                f3_syrk<NW, MAXT>(G, ldg, kdpad, n, nb, dw + d0, hc, nsoc, npad, sd, A, p, sing != 0, Tt, tl, lane, warp, TL);   // :42-46
                if (tid == 0) s_fail[0] = s_fail[1] = 0;
                tsync_t<NW>(bar);
                if (a.g.dbg && b == a.g.dbg_prob && phase == 1 && iters == a.g.dbg_iter) {
                    double* o = a.g.dbg + 2 * k;
                    for (int q = tid; q < n * n; q += T) o[q] = t3_sym(Tt, q % n, q / n);
                    tsync_t<NW>(bar);
                }
                PT3_MARK(P3_SYRK);
                int ok = f3_chol_inv<NW>(Tt, Lp, tij, nb, s_fail, lane, warp, TL, a.P.d_clk, bar, bar2);          // :47
                PT3_MARK(P3_CHOL);
                if (!ok && phase == 0 && a.sing_detect && !sing && p > 0) {
                    // cholesky(G'G) threw: the reference's `sing` (src/Socp.jl:49-56).  Repeat the initial point with A'A.
                    tsync_t<NW>(bar);
                    sing = 1;
                    if (tid == 0 && a.g.sing_out) a.g.sing_out[b] = 1;
                    continue;
                }
                if (ok) f3_xtx<NW, MAXT>(Tt, nb, tl, lane, warp, TL, bar);                        // :48  Li = H^-1 (explicit)
                tsync_t<NW>(bar);
                PT3_MARK(P3_XTX);
                if (ok && p > 0) {
                    for (int q = 0; q < p; ++q)          // HiAt = H^-1 A'                           :49
                        f3_symv<NW>(Tt, nb, At + q * npad, lane, warp, TL, [&](int r, double acc) { HiAt[q * npad + r] = acc; });
                    tsync_t<NW>(bar);
                    if (p == 1) {
                        // one equality row (the budget row of C2): M is a scalar
                        if (warp == 0) {
                            double acc = 0.0;
                            for (int c = lane; c < n; c += 32) acc = fma(A[c], HiAt[c], acc);
                            acc = warp_sum(acc);
                            if (lane == 0) {
                                if (!(acc > 0.0)) s_fail[0] = 1;
                                Minv[0] = 1.0 / acc;
                            }
                        }
                        tsync_t<NW>(bar);
                        ok = !s_fail[0];
                    } else if (p <= 8) {
                        // warp 0 alone forms M = A HiAt (:50), factors it (:51) and inverts it (8 x 12 scratch in Lp)
                        double* M8 = Lp;
                        double* D8 = Lp + 96;
                        if (warp == 0) {
                            for (int e = lane; e < 96; e += 32) M8[e] = 0.0;
                            __syncwarp();
                            for (int e = 0; e < p * p; ++e) {
                                const int i = e % p, j = e / p;
                                double acc = 0.0;
                                for (int c = lane; c < n; c += 32) acc = fma(A[c * p + i], HiAt[j * npad + c], acc);
                                acc = warp_sum(acc);
                                if (lane == 0) M8[j * 12 + i] = acc;
                            }
                            for (int i = p + lane; i < 8; i += 32) M8[i * 12 + i] = 1.0;
                            __syncwarp();
                            double xc[8];
                            const int okm = f2_diag_factor<false>(M8, 12, lane, xc);
                            if (!okm && lane == 0) s_fail[0] = 1;
                            if (lane < 8) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) D8[i * 12 + lane] = xc[i];      // X = chol(M)^-1, row-major
                            }
                            __syncwarp();
                            for (int e = lane; e < p * p; e += 32) {                        // Minv = X'X
                                const int i = e % p, j = e / p;
                                double acc = 0.0;
                                for (int m = max(i, j); m < p; ++m) acc = fma(D8[m * 12 + i], D8[m * 12 + j], acc);
                                Minv[j * p + i] = acc;
                            }
                        }
                        tsync_t<NW>(bar);
                        ok = !s_fail[0];
                    } else {
                        // M = A HiAt into packed tiles (unit pad diagonal), blocked factor + inverse, Minv = X'X
                        for (int e = tid; e < pb * (pb + 1) / 2 * 64; e += T) Mt[e] = 0.0;
                        tsync_t<NW>(bar);
                        for (int e = tid; e < ppad * ppad; e += T) {
                            const int i = e % ppad, j = e / ppad;
                            if (i < j) continue;
                            double acc = 0.0;
                            if (i < p && j < p) { for (int c = 0; c < n; ++c) acc = fma(A[c * p + i], HiAt[j * npad + c], acc); }
                            else acc = (i == j) ? 1.0 : 0.0;
                            Mt[t3_idx(i >> 3, j >> 3) * 64 + t3_off(i & 7, j & 7)] = acc;
                            if ((i >> 3) == (j >> 3)) Mt[t3_idx(i >> 3, j >> 3) * 64 + t3_off(j & 7, i & 7)] = acc;
                        }
                        tsync_t<NW>(bar);
                        ok = f3_chol_inv<NW>(Mt, MLp, tijm, pb, s_fail, lane, warp, TL, nullptr, bar, bar2);        // :51
                        if (ok) {
                            for (int q = tid; q < p * p; q += T) {          // Minv = X'X, X lower triangular in packed tiles
                                const int i = q % p, j = q / p;
                                double acc = 0.0;
                                for (int m = max(i, j); m < p; ++m)
                                    acc = fma(Mt[t3_idx(m >> 3, i >> 3) * 64 + t3_off(m & 7, i & 7)],
                                              Mt[t3_idx(m >> 3, j >> 3) * 64 + t3_off(m & 7, j & 7)], acc);
                                Minv[j * p + i] = acc;
                            }
                            tsync_t<NW>(bar);
                        }
                    }
                    if (ok) {
                        // K = HiAt Minv and md = Minv dy: then cy = K'n0 - md and cx = H^-1 n0 - HiAt cy, which is
                        // src/densesolver.jl:73-83 (m0 = A H^-1 n0 - dy, cy = M^-1 m0, cx = H^-1 (n0 - A'cy)).  K HiAt'
                        // is NOT folded into the packed symmetric H^-1: mirrored across the diagonal its rounding
                        // would no longer cancel against A cx = dy row by row.  Inside the loop a sing problem adds
                        // H^-1 A'dy to cx (the m0 = by - cy of :76-78).
                        for (int e = tid; e < p * n; e += T) {
                            const int q = e / n, r = e - q * n;
                            double kq = 0.0;
                            for (int t = 0; t < p; ++t) kq = fma(HiAt[t * npad + r], Minv[q * p + t], kq);
                            Km[q * npad + r] = kq;
                        }
                        for (int i = tid; i < p; i += T) {              // md = Minv dy
                            double acc = 0.0;
                            for (int r = 0; r < p; ++r) acc = fma(Minv[r * p + i], dy[r], acc);
                            md[i] = acc;
                        }
                        const bool dbl = sing && phase != 0;
                        for (int i = tid; i < n; i += T) {              // kd = H^-1 A'dy (sing, inside the loop) or 0
                            double acc = 0.0;
                            if (dbl) for (int q = 0; q < p; ++q) acc = fma(HiAt[q * npad + i], dy[q], acc);
                            kd_[i] = acc;
                        }
                    }
                }
                PT3_MARK(P3_EQ);
                if (!ok) { status = ST_NUMERICAL; break; }                                   // cholesky! threw
            }
            // ---- middle of solve_kkt, src/densesolver.jl:66-85: out cx, cy, u = G cx - k2
            tsync_t<NW>(bar);
            if (a.g.dbg && b == a.g.dbg_prob && phase == a.g.dbg_phase && iters == a.g.dbg_iter) {
                double* o = a.g.dbg + 2 * k + n * n;
                for (int i = tid; i < n; i += T) o[i] = sc * dx[i];
                for (int i = tid; i < p; i += T) o[n + i] = sc * dy[i];
                for (int i = tid; i < k; i += T) { o[n + p + i] = sc * dz[i]; o[n + p + k + i] = ds[i]; }
            }
            if (p == 0) {
                f3_symv<NW>(Tt, nb, n0, lane, warp, TL, [&](int r, double acc) { if (r < n) cx[r] = acc; });
            } else if (p <= 8) {
                // every warp forms cy = K'n0 - sc md for itself (p short dot products), so that cx = H^-1 n0 - HiAt cy
                // needs no barrier between the two
                double cyr[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    cyr[q] = 0.0;
                    if (q < p) {
                        double acc = 0.0;
                        for (int c = lane; c < n; c += 32) acc = fma(Km[q * npad + c], n0[c], acc);
                        cyr[q] = warp_sum(acc) - sc * md[q];
                    }
                }
                if (warp == 0 && lane < p) {
                    double v = cyr[0];
#pragma unroll
                    for (int q = 1; q < 8; ++q) v = (lane == q) ? cyr[q] : v;
                    cy[lane] = v;
                }
                f3_symv<NW>(Tt, nb, n0, lane, warp, TL, [&](int r, double acc) {
                    if (r < n) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            if (q < p) acc = fma(-HiAt[q * npad + r], cyr[q], acc);
                        cx[r] = acc + sc * kd_[r];
                    }
                });
            } else {
                f3_symv<NW>(Tt, nb, n0, lane, warp, TL, [&](int r, double acc) { if (r < n) cx[r] = acc + sc * kd_[r]; });
                for (int q = warp; q < p; q += NW) {              // cy = K'n0 - sc md
                    double acc = 0.0;
                    for (int c = lane; c < n; c += 32) acc = fma(Km[q * npad + c], n0[c], acc);
                    acc = warp_sum(acc);
                    if (lane == 0) cy[q] = acc - sc * md[q];
                }
                tsync_t<NW>(bar);
                for (int r = tid; r < n; r += T) {                // cx = H^-1 n0 - HiAt cy
                    double acc = cx[r];
                    for (int q = 0; q < p; ++q) acc = fma(-HiAt[q * npad + r], cy[q], acc);
                    cx[r] = acc;
                }
            }
            tsync_t<NW>(bar);
            PT3_MARK(P3_SOLVE_A);
            const double* cxv = cx;
            gemv_rows_v<NW>(G, ldg, kd, n, cxv, D::split_kd(P), lane, warp, [&](int r, double acc) { u[d0 + r] = acc - k2[d0 + r]; });   // :84-85
            for (int i = tid; i < k - kd; i += T) {
                const int r = i < d0 ? i : i + kd;
                u[r] = sing_row(r, cxv) - k2[r];
            }
            tsync_t<NW>(bar);
            PT3_MARK(P3_SOLVE);

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
                team_reduce<NW, 0, 2>(r2, fl, scr + (scr_par ^= 1) * 8 * NW, lane, warp, bar);
                const double mp = r2[0], mdl = r2[1];
                const bool shp = !(fabs(mp) < prm.init_eps), shd = !(fabs(mdl) < prm.init_eps);
                for (int i = tid; i < k; i += T) {
                    const double z0 = u[i];
                    s[i] = -z0;
                    z[i] = z0;
                }
                tsync_t<NW>(bar);
                for (int i = tid; i < kpoc; i += T) {
                    if (shp) s[i] += 1.0 + mp;
                    if (shd) z[i] += 1.0 + mdl;
                }
                for (int c = tid; c < nsoc; c += T) {
                    const int o = D::soc_offs(P, c);
                    if (shp) s[o] += 1.0 + mp;
                    if (shd) z[o] += 1.0 + mdl;
                }
                tsync_t<NW>(bar);
                PT3_MARK(P3_INIT);
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
                PT3_MARK(P3_TAIL);
                team_reduce<NW, 1, 1>(r2, fl, scr + (scr_par ^= 1) * 8 * NW, lane, warp, bar);
                if (NW == 1) __syncwarp();
                if (a.g.dbg && b == a.g.dbg_prob && phase == a.g.dbg_phase && iters == a.g.dbg_iter) {
                    double* o = a.g.dbg + 2 * k + n * n + n + p + 2 * k;
                    for (int i = tid; i < n; i += T) o[i] = cxv[i];
                    for (int i = tid; i < p; i += T) o[n + i] = cy[i];
                    for (int i = tid; i < k; i += T) { o[n + p + i] = u[i]; o[n + p + k + i] = k0[i]; }
                }
                const double tstep = step_from_t(r2[1]);                             // src/solver.jl:130 / :145
                if (phase == 1) {
                    // centering parameter (:130-134), combined right-hand side (:136-140)
                    const double rho = 1.0 - tstep - tstep * tstep * r2[0] * fast_rcp(ll);   // :132 (minus: reference quirk)
                    const double cl = fmax(0.0, fmin(1.0, rho));
                    const double sig = cl * cl * cl;                                 // :133
                    const double mu = ll / (double)a.g.deg;                          // :134
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
                    tsync_t<NW>(bar);
                    new_iter = true;
                }
                PT3_MARK(P3_MIDPOST);
            }

            if (new_iter) {
                // ---- top of a Mehrotra iteration, src/solver.jl:105-126
                if (iters >= prm.max_iter) break;
                if (TEAMS > 1 && a.align) { __syncthreads(); __syncthreads(); }      // re-align the teams of this CTA
                if (a.g.dbg && b == a.g.dbg_prob && iters == a.g.dbg_iter)
                    for (int i = tid; i < k; i += T) { a.g.dbg[i] = s[i]; a.g.dbg[k + i] = z[i]; }
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
                gemv_cols_vu<NW, 16 / NW>(G, ldg, kd, n, z + d0, lane, warp, [&](int c, double acc) {
                    double v = -(acc + sing_col(c, z)) - cv[c];
                    for (int q = 0; q < p; ++q) v = fma(-A[c * p + q], y[q], v);
                    dx[c] = v;
                    r4[0] = fma(v, v, r4[0]);
                });
                gemv_rows_v<NW>(G, ldg, kd, n, x, D::split_kd(P), lane, warp,
                                [&](int r, double acc) { dz[d0 + r] = -acc - s[d0 + r] + hv[d0 + r]; });
                for (int i = tid; i < k - kd; i += T) {
                    const int r = i < d0 ? i : i + kd;
                    dz[r] = -sing_row(r, x) - s[r] + hv[r];
                }
                if (p > 0)
                    gemv_rows<NW, false>(A, p, p, n, x, 1, D::split_p(P), lane, warp, [&](int r, double acc) {
                        const double v = -acc + bv[r];
                        dy[r] = v;
                        r4[1] = fma(v, v, r4[1]);
                    });
                team_reduce<NW, 4, 0>(r4, fl, scr + (scr_par ^= 1) * 8 * NW, lane, warp, bar);
                if (NW == 1) __syncwarp();
                PT3_MARK(P3_RESID);
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
                // per column: A'dy (sing right-hand side) and the singleton rows' share of diag(G'W^-2 G)
                for (int i = tid; i < n; i += T) {
                    if (p > 0) {
                        double acc = 0.0;
                        for (int q = 0; q < p; ++q) acc = fma(A[i * p + q], dy[q], acc);
                        atdy[i] = acc;
                    }
                    double sdv = 0.0;
                    if (ident) sdv = dw[i] * sval[i] * sval[i];
                    else
                        for (int e = scol_ptr[i]; e < scol_ptr[i + 1]; ++e) { const int r = scol_rows[e]; const double g = sval[r]; sdv = fma(dw[r] * g, g, sdv); }
                    sd[i] = sdv;
                }
                // hc[c] = sqrt(2)/eta G_c'q, q = J wbar: the rank-one part of G'W^-2 G per cone (densesolver :41-43).
                // G_c'q = 2 w0 G[head] - G_c'wbar: a gemv over the cone's dense rows plus its singleton rows.
                for (int c = 0; c < nsoc; ++c) {
                    const int co = D::soc_offs(P, c), cd = D::soc_dim(P, c);
                    const int lo = max(co, d0), hi = min(co + cd, d0 + kd);
                    const double* csc = cs + c * F2_CS;
                    const double f = 1.4142135623730951 * csc[CS_IE], w0 = csc[CS_W0];
                    gemv_cols<NW, false>(G + (lo - d0), ldg, max(hi - lo, 0), n, wb + lo, lane, warp, [&](int col, double acc) {
                        double head;
                        if (co >= d0 && co < d0 + kd) head = G[col * ldg + co - d0];
                        else head = (srow_col[co] == col) ? sval[co] : 0.0;
                        if (!ident || co < n) {      // singleton rows inside this cone
                            for (int e = scol_ptr[col]; e < scol_ptr[col + 1]; ++e) {
                                const int r = scol_rows[e];
                                if (r >= co && r < co + cd) acc = fma(sval[r], wb[r], acc);
                            }
                        }
                        hc[c * npad + col] = f * (2.0 * w0 * head - acc);
                    });
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
            tsync_t<NW>(bar);
            PT3_MARK(P3_HEAD_GT);
        }
        if (status == ST_RUNNING) status = ST_MAXITER;
        tsync_t<NW>(bar);

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
            team_reduce<NW, 2, 0>(r2, fl, scr + (scr_par ^= 1) * 8 * NW, lane, warp, bar);
            if (tid == 0) {
                a.g.pobj[b] = r2[0];
                a.g.dobj[b] = r2[1];
                a.g.status[b] = status;
                a.g.iters[b] = iters;
                a.g.active[b] = 0;
                a.g.fail[b] = (status == ST_NUMERICAL);
            }
        }
        tsync_t<NW>(bar);
        PT3_MARK(P3_OUT);
    }
}

// compile-time specialised layout (BASELINE.json C2 = portfolio n=50, p=1, POC 50 + SOC 51: rows 0..49 = -I, row 50
// empty, rows 51..100 dense)
using Dims3C2 = Dims3Static<4, 50, 1, 50, 1, 51, 51, 50, 1>;

#if !defined(SOCP_SIMT_EMU) && !defined(__CUDACC_RTC__)
template <int NW, int TEAMS, int MAXT, int MINB, class D>
inline void fused3_launch(const F3Plan& plan, const F3Args& args, int grid, cudaStream_t stream) {
    const size_t smem = plan.smem * TEAMS;
    cudaFuncSetAttribute(k_fused3<NW, TEAMS, MAXT, MINB, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_fused3<NW, TEAMS, MAXT, MINB, D><<<grid, NW * 32 * TEAMS, smem, stream>>>(args);
}
void fused3_launch_c2(const F3Plan& plan, const F3Args& args, int teams, int grid, cudaStream_t stream);    // fused3.cu
void fused3_launch_dyn(const F3Plan& plan, const F3Args& args, int teams, int grid, cudaStream_t stream);   // fused3_dyn.cu
// run-time specialisation of the kernel for the plan's layout (lane_jit.cu)
void* fused3_jit_get(const F3Plan& plan, int teams, int device);
bool fused3_jit_launch(void* fn, const F3Plan& plan, const F3Args& args, int teams, int grid, cudaStream_t stream);

// Solves problems [first, first + batch) of the shard.  counter_slot: which of the plan's work counters this launch
// uses (launches that may overlap need different ones).  Batches that give every SM several problems run with four
// teams per CTA (one CTA per SM, teams aligned per iteration); smaller ones with one team per CTA, spread over the SMs.
inline void solve_fused3(const F3Plan& plan, const F3Glob& g, int first, int batch, const LoopParams& prm, int sing_detect,
                         int verify, cudaStream_t stream, bool allow_static = true, int counter_slot = 0) {
    cudaMemsetAsync(plan.d_counter + counter_slot, 0, sizeof(int), stream);
    F3Args args;
    args.g = g;
    args.P = plan;
    args.prm = prm;
    args.first = first;
    args.batch = batch;
    args.counter = plan.d_counter + counter_slot;
    args.sing_detect = sing_detect;
    args.verify = verify;
    // experiment switch (default off): with the chain warps spread over the sub-partitions, re-aligning the teams every
    // iteration costs 2 % (1.107M against 1.126M problems/s on C2); before that spread it gained 16 %
    { const char* e = getenv("SOCP_B200_F3_ALIGN"); args.align = e ? atoi(e) : 0; }
    // four teams per CTA whenever they fit: one kernel per layout whatever the batch size, so that the results do not
    // depend on how a batch is cut into shards or chunks
    const char* e4 = getenv("SOCP_B200_F3_TEAMS");          // experiments only
    const int teams = (plan.teams4 && !(e4 && atoi(e4) == 1)) ? 4 : 1;
    const int grid = teams > 1 ? std::min(batch, plan.num_sms) : std::min(batch, plan.num_sms * plan.ctas_per_sm);
    if (allow_static && Dims3C2::matches(plan)) fused3_launch_c2(plan, args, teams, grid, stream);
    else if (allow_static && plan.jit_fn && plan.jit_teams == teams && fused3_jit_launch(plan.jit_fn, plan, args, teams, grid, stream)) {}
    else fused3_launch_dyn(plan, args, teams, grid, stream);
}
// non-inline entry of solve_fused3, compiled once in fused3.cu (the kernels are instantiated there only)
void solve_fused3_ext(const F3Plan& plan, const F3Glob& g, int first, int batch, const LoopParams& prm, int sing_detect,
                      int verify, cudaStream_t stream, bool allow_static, int counter_slot);
#endif

}  // namespace socp
