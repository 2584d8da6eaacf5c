// fused_small.cuh -- whole-solve kernel for layouts that fit in shared memory.
//
// One CTA owns one problem from the initial point to the last Mehrotra step
// (reference src/solver.jl:68-152): G, the scaled copy Gt = W^-1 G, the reduced
// KKT matrix and every work vector stay in shared memory; the only global
// traffic is the problem data in and the iterate out.  CTAs are persistent and
// pull problems from an atomic counter (iteration counts differ per problem).
//
//   * cone math: the warp-shuffle primitives of cone_ops.cuh on shared vectors
//   * H = Gt'Gt: mma.sync m8n8k4 f64 (SASS DMMA.8x8x4) straight from the
//     column-major Gt in shared memory (leading dimension == 4 or 12 mod 16 so
//     that the fragment loads are bank-conflict free); accumulators in registers;
//     H is then written over Gt's storage (Gt is dead once the SYRK is done)
//   * Cholesky + explicit inverse of the triangular factor in one sweep of n
//     steps with ONE barrier per step (unscaled elimination, scaling deferred);
//     the solves are then two triangular gemvs each -- the reference too applies
//     an explicit inverse (Li, src/densesolver.jl:48,83)
//
// Restrictions (the tiled path takes everything else): no `sing` problems,
// p <= 32, and at most 12 8x8 tiles of H per warp.
#pragma once
#include "tiled_kernels.cuh"
#include "linalg.cuh"
#include <vector>
#include <algorithm>

namespace socp {

struct FusedOffsets {   // offsets (in doubles) into the dynamic shared memory
    int G, GtH, A, HiAt, M, Mi;
    int c, b, h, x, y, z, s, lam, wb, iwb, eta, dx, dy, dz, ds, rx, ry, rz, rs, k0, k2, u, kt2, kt3, t1, t2, dinv;
    int ldg, npad, kpad4, ldh, ldm;
    int Hoff, Lioff;    // inside the GtH region
    int total;
};

struct FusedPlan {
    bool fits = false;
    int threads = 256;
    int variant = 2;       // kernel instantiation: 0: n<=16, 1: n<=32, 2: n<=64, 3: n<=104
    int ctas_per_sm = 1;
    int num_sms = 148;
    size_t smem = 0;
    FusedOffsets off{};
    int* d_counter = nullptr;
};

inline int fused_ld(int k) {    // smallest ld >= k with ld == 4 (mod 8): conflict-free DMMA fragment loads
    int ld = k;
    while (ld % 8 != 4) ++ld;
    return ld;
}

inline void fused_plan(FusedPlan& plan, int n, int p, int k, const std::vector<int>& kind,
                       const std::vector<int>& offs, const std::vector<int>& dim, int device) {
    (void)kind; (void)offs;
    plan.fits = false;
    FusedOffsets& o = plan.off;
    const int nc = (int)dim.size();
    o.ldg = fused_ld(k);
    o.kpad4 = (k + 3) / 4 * 4;
    o.npad = (n + 7) / 8 * 8;
    o.ldh = n | 1;                      // odd: conflict-free row and column walks
    o.ldm = p | 1;
    int at = 0;
    auto take = [&](int cnt) { int r = at; at += (cnt + 1) / 2 * 2; return r; };   // keep 16-byte alignment
    o.G = take(o.ldg * o.npad);
    const int gt_sz = o.ldg * o.npad;
    const int hl_sz = 2 * o.ldh * n;
    o.GtH = take(std::max(gt_sz, hl_sz));
    o.Hoff = 0;
    o.Lioff = o.ldh * n;
    o.A = take(p * n);
    o.HiAt = take(n * p);
    o.M = take(o.ldm * p);
    o.Mi = take(o.ldm * p);
    o.c = take(n); o.x = take(n); o.dx = take(n); o.rx = take(n); o.t1 = take(n); o.t2 = take(n); o.dinv = take(n + p);
    o.b = take(p); o.y = take(p); o.dy = take(p); o.ry = take(p);
    o.h = take(k); o.z = take(k); o.s = take(k); o.lam = take(k); o.wb = take(k); o.iwb = take(k);
    o.dz = take(k); o.ds = take(k); o.rz = take(k); o.rs = take(k); o.k0 = take(k); o.k2 = take(k); o.u = take(k);
    o.kt2 = take(k); o.kt3 = take(k);
    o.eta = take(4 * nc);
    o.total = at;
    plan.smem = (size_t)at * sizeof(double) + 256;   // + static scratch headroom
    int dev_smem = 0, sms = 148;
    if (cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    plan.num_sms = sms;
    if (plan.smem > (size_t)dev_smem) return;
    if (nc > 64) return;                 // FUSED_MAX_CONES
    // kernel variants <MAXT, NP, NWARPS>: NP >= max(n, p) is the register-tiled Cholesky size, MAXT bounds the
    // number of 8x8 tiles of H a warp accumulates: ntl = nt(nt+1)/2 <= MAXT * NWARPS
    if (n <= 16 && p <= 16) { plan.variant = 0; plan.threads = 64; }        // nt<=2: 3 tiles  <= 2*2
    else if (n <= 32 && p <= 32) { plan.variant = 1; plan.threads = 128; }  // nt<=4: 10 tiles <= 4*4
    else if (n <= 64 && p <= 64) { plan.variant = 2; plan.threads = 256; }  // nt<=8: 36 tiles -> needs MAXT 5: see below
    else if (n <= 104 && p <= 104) { plan.variant = 3; plan.threads = 256; }  // nt<=13: 91 tiles <= 12*8
    else return;
    if (p > 32) return;
    if (plan.variant == 2 && o.npad > 56) plan.variant = 3;                 // 57..64 -> 36 tiles > 4*8
    // resident CTAs per SM: shared memory (1 KB reserved per CTA) and threads
    const int per_sm = 228 * 1024;
    const int reg_cap = plan.variant == 0 ? 12 : (plan.variant == 1 ? 4 : (plan.variant == 2 ? 2 : 1));
    plan.ctas_per_sm = std::max(1, std::min({(int)(per_sm / (plan.smem + 1024)), 2048 / plan.threads, reg_cap}));
    plan.fits = true;
}

// ------------------------------------------------------------------ CTA-level dense helpers (shared memory)
// out[c] = alpha * sum_r M[r,c] x[r] + c1*v1[c] + c2*v2[c].  A quarter-warp per column (8 columns per warp
// pass, 4 lanes striding the rows), two shuffle steps to reduce.
__device__ __forceinline__ void cta_gemv_t(const double* M, int ld, int rows, int cols, const double* x, double* out,
                                           double alpha, const double* v1, double c1, const double* v2, double c2) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int cq = lane >> 2, rl = lane & 3;
    for (int c0 = warp * 8; c0 < cols; c0 += nw * 8) {
        const int c = c0 + cq;
        const bool ok = c < cols;
        const double* col = M + (ok ? c : 0) * ld;
        double a0 = 0.0, a1 = 0.0;
        int r = rl;
        for (; r + 4 < rows; r += 8) {
            a0 = fma(col[r], x[r], a0);
            a1 = fma(col[r + 4], x[r + 4], a1);
        }
        if (r < rows) a0 = fma(col[r], x[r], a0);
        double acc = a0 + a1;
        acc += __shfl_xor_sync(FULL_MASK, acc, 1);
        acc += __shfl_xor_sync(FULL_MASK, acc, 2);
        if (rl == 0 && ok) {
            acc *= alpha;
            if (v1) acc += c1 * v1[c];
            if (v2) acc += c2 * v2[c];
            out[c] = acc;
        }
    }
}
// out[r] = alpha * sum_c M[r,c] x[c] + c1*v1[r] + c2*v2[r]     (thread per row)
__device__ __forceinline__ void cta_gemv_n(const double* M, int ld, int rows, int cols, const double* x, double* out,
                                           double alpha, const double* v1, double c1, const double* v2, double c2) {
    for (int r = threadIdx.x; r < rows; r += blockDim.x) {
        double a0 = 0.0, a1 = 0.0;
        int c = 0;
        for (; c + 1 < cols; c += 2) {
            a0 = fma(M[c * ld + r], x[c], a0);
            a1 = fma(M[(c + 1) * ld + r], x[c + 1], a1);
        }
        if (c < cols) a0 = fma(M[c * ld + r], x[c], a0);
        double acc = alpha * (a0 + a1);
        if (v1) acc += c1 * v1[r];
        if (v2) acc += c2 * v2[r];
        out[r] = acc;
    }
}

// H (lower, n x n, ld ldh) = Gt' Gt on the FP64 tensor pipe.  Gt: kpad4 x npad column-major, ld ldg, pad rows
// zero.  Phase 1 accumulates in registers; after the barrier phase 2 stores (H may alias Gt).
template <int MAXT>
__device__ __forceinline__ void cta_syrk_dmma(const double* Gt, int ldg, int kpad4, int n, int npad, double* H, int ldh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int nt = npad >> 3, ntl = nt * (nt + 1) / 2;
    const int fr = lane >> 2, fk = lane & 3;
    const double* pa[MAXT];
    const double* pb[MAXT];
    int ti[MAXT], tj[MAXT];
    bool valid[MAXT];
    double acc[MAXT][2];
#pragma unroll
    for (int q = 0; q < MAXT; ++q) {
        const int t = warp + q * nw;
        valid[q] = t < ntl;
        int a = 0;
        while ((a + 1) * (a + 2) / 2 <= t) ++a;
        ti[q] = a;
        tj[q] = t - a * (a + 1) / 2;
        pa[q] = Gt + (ti[q] * 8 + fr) * ldg + fk;
        pb[q] = Gt + (tj[q] * 8 + fr) * ldg + fk;
        acc[q][0] = acc[q][1] = 0.0;
    }
#pragma unroll 2
    for (int kk = 0; kk < kpad4; kk += 4) {
#pragma unroll
        for (int q = 0; q < MAXT; ++q) {
            if (valid[q]) dmma884(acc[q][0], acc[q][1], pa[q][kk], pb[q][kk]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < MAXT; ++q) {
        if (valid[q]) {
            const int gi = ti[q] * 8 + fr;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gj = tj[q] * 8 + 2 * fk + e;
                if (gi < n && gj <= gi) H[gj * ldh + gi] = acc[q][e];
            }
        }
    }
}

// Cholesky of H (n x n lower, column-major, ld) together with the explicit inverse of the triangular factor:
// on exit Li (same shape) holds L^-1.  H is destroyed (its columns end up holding L's columns times their pivot's
// square root).  Register-tiled: the NP x NP working matrix lives in registers, warp w owning rows
// [w*RPW, (w+1)*RPW) and lane l the columns l, l+32, ...  One barrier per column: before it the owners publish
// column j (into H's column j), row j of the partial inverse (into Li's row j) and the reciprocal pivot; after it
// every warp that still has rows below j applies the rank-1 update  wk[i][c] -= m_i * v_c.  Unscaled elimination
// (see DESIGN.md): entries left of column j hold the inverse being built, entries right of it the Schur
// complement.  *fail is set on a pivot that is not > 0 (cholesky!'s PosDefException, src/densesolver.jl:47,51).
template <int NP, int NWARPS>
__device__ __forceinline__ int cta_chol_inv(double* H, double* Li, int n, int ld, double* ipiv, int* fail,
                                            const int tid, const int lane, const int warp) {
    constexpr int RPW = NP / NWARPS;
    constexpr int CPL = (NP + 31) / 32;
    const int r0 = warp * RPW;
    const int nrow = min(RPW, n - r0);             // rows of this warp that exist (may be <= 0)
    double wk[RPW][CPL];
#pragma unroll
    for (int a = 0; a < RPW; ++a)
#pragma unroll
        for (int b = 0; b < CPL; ++b) {
            const int r = r0 + a, c = lane + 32 * b;
            wk[a][b] = (r < n && c <= r) ? H[c * ld + r] : 0.0;
        }
    __syncthreads();
    int ok = 1;
#pragma unroll
    for (int bj = 0; bj < CPL; ++bj) {             // column block: the register column index is static
        const int jend = ok ? min(32, n - 32 * bj) : 0;
        for (int jl = 0; jl < jend; ++jl) {
            const int j = 32 * bj + jl;
            double* colj = H + j * ld;             // column j of H doubles as the broadcast buffer
            // ---- publish column j (every warp: its rows), row j and 1/pivot (the warp that owns row j)
            if (lane == jl) {
#pragma unroll
                for (int a = 0; a < RPW; ++a)
                    if (a < nrow) colj[r0 + a] = wk[a][bj];
            }
            const int aj = j - r0;
            if (aj >= 0 && aj < RPW) {             // warp-uniform
#pragma unroll
                for (int b = 0; b < CPL; ++b) {
                    double v = wk[0][b];
#pragma unroll
                    for (int a = 1; a < RPW; ++a) v = (a == aj) ? wk[a][b] : v;
                    const int c = lane + 32 * b;
                    if (c < j) Li[c * ld + j] = v;
                    else if (c == j) { Li[c * ld + j] = 1.0; ipiv[j] = fast_rcp(v); }
                }
            }
            __syncthreads();
            const double piv = colj[j];
            if (!(piv > 0.0)) { ok = 0; break; }   // uniform: every thread reads the same word
            if (r0 + RPW - 1 > j && nrow > 0) {    // this warp still has rows below j
                const double ip = ipiv[j];
                double m[RPW];
#pragma unroll
                for (int a = 0; a < RPW; ++a) {
                    const double hv = (a < nrow) ? colj[r0 + a] : 0.0;
                    m[a] = (r0 + a > j) ? hv * ip : 0.0;
                }
#pragma unroll
                for (int b = 0; b < CPL; ++b) {
                    const int c = lane + 32 * b;
                    double v;
                    if (b > bj) v = (c < n) ? colj[c] : 0.0;            // c > j: Schur-complement part
                    else if (b < bj) v = Li[c * ld + j];                // c < j: inverse part
                    else v = (c > j) ? ((c < n) ? colj[c] : 0.0) : ((c < j) ? Li[c * ld + j] : 0.0);
#pragma unroll
                    for (int a = 0; a < RPW; ++a) wk[a][b] = fma(-m[a], v, wk[a][b]);
                }
                if (lane == jl) {                  // column j itself turns into the inverse's column: -m
#pragma unroll
                    for (int a = 0; a < RPW; ++a) wk[a][bj] = (r0 + a > j) ? -m[a] : wk[a][bj];
                }
            }
        }
    }
    if (!ok) {
        if (tid == 0) *fail = 1;
        __syncthreads();
        return 0;
    }
    __syncthreads();
    // deferred scaling of the inverse: Li[i][c] = y[i][c] / sqrt(pivot_i)
    for (int i = tid; i < n; i += NWARPS * 32) ipiv[i] = fast_rsqrt(H[i * ld + i]);
    __syncthreads();
    for (int c = warp; c < n; c += NWARPS)
        for (int i = c + lane; i < n; i += 32) Li[c * ld + i] *= ipiv[i];
    __syncthreads();
    return 1;
}

// x <- (L L')^-1 x through the explicit triangular inverse: tmp = Li x ; x = Li' tmp.
__device__ __forceinline__ void cta_potrs_inv(const double* Li, int n, int ld, double* x, double* tmp) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {         // tmp[i] = sum_{c<=i} Li[i][c] x[c]
        double a0 = 0.0, a1 = 0.0;
        int c = 0;
        for (; c + 1 <= i; c += 2) {
            a0 = fma(Li[c * ld + i], x[c], a0);
            a1 = fma(Li[(c + 1) * ld + i], x[c + 1], a1);
        }
        if (c <= i) a0 = fma(Li[c * ld + i], x[c], a0);
        tmp[i] = a0 + a1;
    }
    __syncthreads();
    {                                                           // x[c] = sum_{i>=c} Li[i][c] tmp[i], quarter-warp per column
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        const int cq = lane >> 2, rl = lane & 3;
        for (int c0 = warp * 8; c0 < n; c0 += nw * 8) {
            const int c = c0 + cq;
            double a = 0.0;
            if (c < n)
                for (int i = c + rl; i < n; i += 4) a = fma(Li[c * ld + i], tmp[i], a);
            a += __shfl_xor_sync(FULL_MASK, a, 1);
            a += __shfl_xor_sync(FULL_MASK, a, 2);
            if (rl == 0 && c < n) x[c] = a;
        }
    }
    __syncthreads();
}

// Optional per-phase cycle counters (profiling build only: -DSOCP_PHASE_TIMING).  CTA 0 / thread 0 adds the
// clock64() deltas of every phase into phase_clk[]; read back by tools/phase_timing.py.
#ifdef SOCP_PHASE_TIMING
__device__ unsigned long long g_phase_clk[16];
#define PT_INIT() long long pt_t0 = clock64()
#define PT_MARK(idx)                                                              \
    do {                                                                          \
        if (tid == 0 && blockIdx.x == 0) {                                        \
            const long long t_ = clock64();                                       \
            atomicAdd(&g_phase_clk[idx], (unsigned long long)(t_ - pt_t0));       \
            pt_t0 = t_;                                                           \
        }                                                                         \
    } while (0)
#else
#define PT_INIT()
#define PT_MARK(idx)
#endif
enum { PT_LOAD = 0, PT_SYRK, PT_CHOL, PT_EQ, PT_SOLVE, PT_INITSHIFT, PT_MID, PT_POST, PT_SCAL_RESID, PT_PRE, PT_BUILD, PT_OUT };

struct FusedArgs {
    Ws g;                 // global arrays of the shard (problem data, outputs, per-problem words)
    FusedOffsets off;
    LoopParams P;
    int batch;
    int* counter;
};

constexpr int FUSED_MAX_CONES = 64;

template <int MAXT, int NP, int NWARPS, int MINB>
__global__ void __launch_bounds__(NWARPS * 32, MINB) k_fused_solve(FusedArgs a) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double scratch[128];
    __shared__ int iscratch[32];
    __shared__ int s_prob;
    __shared__ int s_kind[FUSED_MAX_CONES], s_offs[FUSED_MAX_CONES], s_dim[FUSED_MAX_CONES];
    const FusedOffsets& o = a.off;
    int tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));     // opaque: keeps ptxas from re-reading SR_TID in hot loops
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int nw = NWARPS;
    double* G = sm + o.G;
    double* Gt = sm + o.GtH;
    double* H = sm + o.GtH + o.Hoff;
    double* Li = sm + o.GtH + o.Lioff;
    double* A = sm + o.A;
    double* HiAt = sm + o.HiAt;
    double* M = sm + o.M;
    double* Mi = sm + o.Mi;
    double* t2 = sm + o.t2;
    double* dinv = sm + o.dinv;

    // shared-memory view of the work set: the loop bodies of tiled_kernels.cuh run on it with b = 0
    Ws w = a.g;
    for (int c = tid; c < a.g.L.ncones; c += blockDim.x) {
        s_kind[c] = a.g.L.kind[c];
        s_offs[c] = a.g.L.offs[c];
        s_dim[c] = a.g.L.dim[c];
    }
    w.L.kind = s_kind; w.L.offs = s_offs; w.L.dim = s_dim;
    const ConeLayout& L = w.L;
    const int n = L.n, p = L.p, k = L.k;
    w.x = sm + o.x; w.y = sm + o.y; w.z = sm + o.z; w.s = sm + o.s;
    w.lam = sm + o.lam; w.wb = sm + o.wb; w.iwb = sm + o.iwb; w.eta = sm + o.eta;
    w.dx = sm + o.dx; w.dy = sm + o.dy; w.dz = sm + o.dz; w.ds = sm + o.ds;
    w.rx = sm + o.rx; w.ry = sm + o.ry; w.rz = sm + o.rz; w.rs = sm + o.rs;
    w.k0 = sm + o.k0; w.k2 = sm + o.k2; w.u = sm + o.u; w.kt2 = sm + o.kt2; w.kt3 = sm + o.kt3;
    w.nactive = nullptr;
    const double* cs = sm + o.c;
    const double* bs = sm + o.b;
    const double* hs = sm + o.h;

    // zero the whole work set once (pad rows/columns of G and Gt must be zero)
    for (int q = tid; q < o.total; q += blockDim.x) sm[q] = 0.0;
    __syncthreads();

    for (;;) {
        if (tid == 0) s_prob = atomicAdd(a.counter, 1);
        __syncthreads();
        const int b = s_prob;
        __syncthreads();
        if (b >= a.batch) break;
        PT_INIT();

        // ---- load the problem (global -> shared)
        {
            const double* Gg = a.g.G + (int64_t)b * a.g.sG;
            for (int col = warp; col < n; col += nw)
                for (int r = lane; r < k; r += 32) G[col * o.ldg + r] = Gg[(int64_t)col * k + r];
            const double* Ag = a.g.A + (int64_t)b * a.g.sA;
            for (int q = tid; q < p * n; q += blockDim.x) A[q] = Ag[q];
            for (int i = tid; i < n; i += blockDim.x) sm[o.c + i] = a.g.c[(int64_t)b * n + i];
            for (int i = tid; i < p; i += blockDim.x) sm[o.b + i] = a.g.b[(int64_t)b * p + i];
            for (int i = tid; i < k; i += blockDim.x) sm[o.h + i] = a.g.h[(int64_t)b * k + i];
        }
        w.status = a.g.status + b; w.iters = a.g.iters + b; w.active = a.g.active + b; w.fail = a.g.fail + b;
        w.sc = a.g.sc + b;
        if (tid == 0) { *w.status = ST_RUNNING; *w.iters = 0; *w.active = 1; *w.fail = 0; }
        __syncthreads();

        // The initial point (src/solver.jl:68-104, W = I) is the same factor + solve as a loop iteration with
        // u = h, dx = -c, dy = b, k2 = h: then rx = x, ry = y and u = G x - h = z0.
        for (int i = tid; i < k; i += blockDim.x) { w.u[i] = hs[i]; w.k2[i] = hs[i]; }
        for (int i = tid; i < n; i += blockDim.x) w.dx[i] = -cs[i];
        for (int i = tid; i < p; i += blockDim.x) w.dy[i] = bs[i];
        __syncthreads();

        PT_MARK(PT_LOAD);
        const double* src = G;      // operand of the SYRK: G for the initial point, Gt = W^-1 G afterwards
        int phase = 0;              // 0: initial point, 1: affine direction (solve #1), 2: combined (solve #2)
        int it = 0;
        bool need_factor = true;
        for (;;) {
            if (need_factor) {
                // ---- KKT factor, src/densesolver.jl:41-52
                cta_syrk_dmma<MAXT>(src, o.ldg, o.kpad4, n, o.npad, H, o.ldh);          // :42-43
                __syncthreads();
                PT_MARK(PT_SYRK);
                int ok = cta_chol_inv<NP, NWARPS>(H, Li, n, o.ldh, dinv, w.fail, tid, lane, warp);        // :47-48
                PT_MARK(PT_CHOL);
                if (ok && p > 0) {
                    for (int q = 0; q < p; ++q) {                                        // HiAt[:,q] = H^-1 A[q,:]'  :49
                        for (int i = tid; i < n; i += blockDim.x) HiAt[q * n + i] = A[i * p + q];
                        __syncthreads();
                        cta_potrs_inv(Li, n, o.ldh, HiAt + q * n, t2);
                    }
                    for (int q = tid; q < p * p; q += blockDim.x) {                      // M = A HiAt  :50
                        const int i = q % p, j = q / p;
                        double acc = 0.0;
                        for (int c = 0; c < n; ++c) acc = fma(A[c * p + i], HiAt[j * n + c], acc);
                        M[j * o.ldm + i] = acc;
                    }
                    __syncthreads();
                    cta_chol_inv<NP, NWARPS>(M, Mi, p, o.ldm, dinv + n, w.fail, tid, lane, warp);         // :51
                }
                PT_MARK(PT_EQ);
            }
            if (!*w.fail) {
                // ---- middle of solve_kkt, src/densesolver.jl:66-85: in u = W^-2 k2, out rx, ry, u = G cx - k2
                cta_gemv_t(G, o.ldg, k, n, w.u, w.rx, 1.0, w.dx, 1.0, nullptr, 0.0);    // n0  :66-67
                __syncthreads();
                cta_potrs_inv(Li, n, o.ldh, w.rx, t2);                                  // t = H^-1 n0
                if (p > 0) {
                    cta_gemv_n(A, p, p, n, w.rx, w.ry, 1.0, w.dy, -1.0, nullptr, 0.0);  // m0 = A t - dy  :73-74
                    __syncthreads();
                    cta_potrs_inv(Mi, p, o.ldm, w.ry, t2);                              // cy  :75
                    cta_gemv_n(HiAt, n, n, p, w.ry, w.rx, -1.0, w.rx, 1.0, nullptr, 0.0);   // cx = t - HiAt cy  :76-83
                    __syncthreads();
                }
                cta_gemv_n(G, o.ldg, k, n, w.rx, w.u, 1.0, w.k2, -1.0, nullptr, 0.0);   // k1 = G cx - k2  :84-85
            }
            __syncthreads();
            PT_MARK(PT_SOLVE);

            if (phase == 0) {
                if (*w.fail) {
                    if (tid == 0) { *w.status = ST_NUMERICAL; *w.active = 0; }
                    for (int i = tid; i < n; i += blockDim.x) w.x[i] = 0.0;
                    for (int i = tid; i < p; i += blockDim.x) w.y[i] = 0.0;
                    for (int i = tid; i < k; i += blockDim.x) { w.z[i] = 0.0; w.s[i] = 0.0; }
                    break;
                }
                for (int i = tid; i < n; i += blockDim.x) w.x[i] = w.rx[i];
                for (int i = tid; i < p; i += blockDim.x) w.y[i] = w.ry[i];
                for (int i = tid; i < k; i += blockDim.x) w.z[i] = w.u[i];
                __syncthreads();
                dev_init_shift(w, 0, a.P, scratch);                                      // src/solver.jl:86-104
                __syncthreads();
                PT_MARK(PT_INITSHIFT);
            } else if (phase == 1) {
                if (!dev_mid(w, 0, a.P, scratch, iscratch)) break;                       // :128-140 + head of solve #2
                __syncthreads();
                PT_MARK(PT_MID);
                phase = 2;
                need_factor = false;
                continue;
            } else {
                if (!dev_post(w, 0, a.P, scratch, iscratch)) break;                      // :143-150
                __syncthreads();
                PT_MARK(PT_POST);
                ++it;
            }
            // ---- top of a Mehrotra iteration, src/solver.jl:105-126
            if (it >= a.P.max_iter) break;
            dev_scaling(L, 0, w.s, w.z, w.lam, w.wb, w.iwb, w.eta, w.fail);                     // :106
            cta_gemv_t(G, o.ldg, k, n, w.z, w.dx, -1.0, cs, -1.0, nullptr, 0.0);         // negated residuals :110-118,:125
            cta_gemv_n(G, o.ldg, k, n, w.x, w.dz, -1.0, w.s, -1.0, hs, 1.0);
            __syncthreads();
            if (p > 0) {
                cta_gemv_t(A, p, p, n, w.y, w.dx, -1.0, w.dx, 1.0, nullptr, 0.0);
                cta_gemv_n(A, p, p, n, w.x, w.dy, -1.0, bs, 1.0, nullptr, 0.0);
                __syncthreads();
            }
            PT_MARK(PT_SCAL_RESID);
            if (!dev_pre(w, 0, a.P, it, scratch)) break;                                 // :120-125 + head of solve #1
            __syncthreads();
            PT_MARK(PT_PRE);
            // Gt = W^-1 G, column by column (setup_iter, :41-43).  H/Li of the previous factorisation alias Gt:
            // its pad rows / pad columns must be zero again.
            {
                const int padr = o.kpad4 - k;
                for (int q = tid; q < padr * o.npad; q += blockDim.x) Gt[(q / padr) * o.ldg + k + (q % padr)] = 0.0;
                for (int q = tid; q < (o.npad - n) * o.kpad4; q += blockDim.x)
                    Gt[(n + q / o.kpad4) * o.ldg + (q % o.kpad4)] = 0.0;
                for (int c = 0; c < L.ncones; ++c) {
                    const int kind = s_kind[c], offs = s_offs[c], dim = s_dim[c];
                    for (int col = warp; col < n; col += nw) {
                        const double* sp = G + col * o.ldg + offs;
                        double* dp = Gt + col * o.ldg + offs;
                        if (kind == KIND_POC) warp_poc_apply<APPLY_WINV>(w.wb + offs, w.iwb + offs, sp, dp, dim, lane);
                        else warp_soc_apply<APPLY_WINV>(w.wb + offs, w.eta + c, L.ncones, sp, dp, dim, lane);
                    }
                }
            }
            __syncthreads();
            PT_MARK(PT_BUILD);
            src = Gt;
            phase = 1;
            need_factor = true;
        }
        __syncthreads();
        if (tid == 0 && *w.active) { *w.status = ST_MAXITER; *w.active = 0; }

        // ---- results: iterate and objectives (pobj = c'x, dobj = -b'y - h'z)
        {
            double po = 0.0, d = 0.0;
            for (int i = tid; i < n; i += blockDim.x) {
                const double xi = w.x[i];
                a.g.x[(int64_t)b * n + i] = xi;
                po = fma(cs[i], xi, po);
            }
            for (int i = tid; i < p; i += blockDim.x) {
                const double yi = w.y[i];
                a.g.y[(int64_t)b * p + i] = yi;
                d = fma(-bs[i], yi, d);
            }
            for (int i = tid; i < k; i += blockDim.x) {
                const double zi = w.z[i];
                a.g.z[(int64_t)b * k + i] = zi;
                a.g.s[(int64_t)b * k + i] = w.s[i];
                d = fma(-hs[i], zi, d);
            }
            po = block_sum(po, scratch);
            d = block_sum(d, scratch);
            if (tid == 0) { w.sc->pobj = po; w.sc->dobj = d; }
        }
        __syncthreads();
        PT_MARK(PT_OUT);
    }
}

template <int MAXT, int NP, int NWARPS, int MINB>
inline void fused_launch(const FusedPlan& plan, const FusedArgs& args, int grid, cudaStream_t stream) {
    cudaFuncSetAttribute(k_fused_solve<MAXT, NP, NWARPS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
    k_fused_solve<MAXT, NP, NWARPS, MINB><<<grid, NWARPS * 32, plan.smem, stream>>>(args);
}

inline void solve_fused(FusedPlan& plan, const Ws& g, int batch, int max_iter, double tol, double step_damp,
                        double init_eps, cudaStream_t stream) {
    cudaMemsetAsync(plan.d_counter, 0, sizeof(int), stream);
    FusedArgs args;
    args.g = g;
    args.off = plan.off;
    args.P = LoopParams{max_iter, tol, step_damp, init_eps};
    args.batch = batch;
    args.counter = plan.d_counter;
    const int grid = std::min(batch, plan.num_sms * plan.ctas_per_sm);
    switch (plan.variant) {
        // last parameter: resident CTAs per SM the register allocation must allow
        case 0: fused_launch<2, 16, 2, 12>(plan, args, grid, stream); break;     // n <= 16  (64 threads, <= 80 regs)
        case 1: fused_launch<4, 32, 4, 4>(plan, args, grid, stream); break;      // n <= 32  (128 threads, <= 128 regs)
        case 2: fused_launch<4, 64, 8, 2>(plan, args, grid, stream); break;      // n <= 64  (256 threads, <= 128 regs)
        default: fused_launch<12, 128, 8, 1>(plan, args, grid, stream); break;   // n <= 104 (256 threads)
    }
}

}  // namespace socp
