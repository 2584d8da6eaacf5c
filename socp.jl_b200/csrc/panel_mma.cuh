// panel_mma.cuh -- DMMA panel kernels of the tiled path's blocked Cholesky (n > what fits the fused kernel).
// Reuses the tile algorithm of fused_v2.cuh (f2_chol_inv).  Replaces LAPACK dpotrf's panel work behind
// cholesky! (reference src/densesolver.jl:47,51).
#pragma once
#include "fused_v2.cuh"

namespace socp {

// ---------------------------------------------------------------------------
// DMMA panel of the blocked Cholesky (replaces k_potrf_diag + k_trsm_panel when the
// panel is a full 64 columns wide):
//   k_potrf_diag_mma : one CTA (8 warps) per problem factors the 64 x 64 diagonal block
//                      with the tile algorithm of the fused kernel (f2_chol_inv: 8x8
//                      tiles, mma.sync f64, the inverse carried along), writes L11 back
//                      and X11 = L11^-1 (dense 64 x 64, zero above the diagonal) to Xd.
//   k_trsm_mma       : L21 = A21 X11' as a DMMA product; CTA x owns 128 rows of the
//                      sub-panel.                grid (ceil((n-j-64)/128), batch)
// ---------------------------------------------------------------------------
constexpr int PANEL_LDH = 68;                                   // 64 + 4: conflict-free fragment loads
constexpr size_t POTRF_MMA_SMEM = sizeof(double) * (64 * PANEL_LDH + f2_xsize(8) + 192) + sizeof(uint2) * f2_trail_base(8, 8);

// Xinv: [batch][nblk][64*64], block j/64 of problem b receives the inverse of its diagonal block (dense column-major,
// zero above the diagonal; a partial last block is padded with the identity).
__global__ void __launch_bounds__(256)
k_potrf_diag_mma(double* __restrict__ H, int64_t strideH, int ldh, int n, int j, double* __restrict__ Xinv, int nblk,
                 int* __restrict__ fail, const int* __restrict__ active) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    extern __shared__ __align__(16) double psm[];
    __shared__ int sfail;
    double* Hs = psm;
    double* Xs = Hs + 64 * PANEL_LDH;
    double* Dinv = Xs + f2_xsize(8);
    uint2* desc = reinterpret_cast<uint2*>(Dinv + 192);
    double* Hb = H + (int64_t)b * strideH + (int64_t)j * ldh + j;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int jb = min(64, n - j);
    if (tid == 0) sfail = 0;
    for (int q = tid; q < 64 * 64; q += 256) {
        const int c = q >> 6, r = q & 63;
        double v = 0.0;
        if (r < jb && c < jb) v = (r >= c) ? Hb[(int64_t)c * ldh + r] : 0.0;
        else if (r == c) v = 1.0;
        Hs[c * PANEL_LDH + r] = v;
    }
    for (int q = tid; q < f2_xsize(8) + 192; q += 256) Xs[q] = 0.0;
    f2_build_trail(desc, 8, PANEL_LDH, 64 * PANEL_LDH, tid, 256);
    __syncthreads();
    const int ok = f2_chol_inv<8, true>(Hs, Xs, Dinv, desc, 8, PANEL_LDH, &sfail, lane, warp);
    if (!ok) {
        if (tid == 0) fail[b] = 1;
        return;
    }
    double* Xb = Xinv + ((int64_t)b * nblk + j / 64) * 64 * 64;
    for (int q = tid; q < 64 * 64; q += 256) {
        const int c = q >> 6, r = q & 63;
        if (r >= c && r < jb) Hb[(int64_t)c * ldh + r] = Hs[c * PANEL_LDH + r];
        const int cb = c >> 3;
        Xb[q] = (r >= cb * 8) ? Xs[f2_xbase(8, cb) + (c & 7) * f2_xld(8, cb) + r - cb * 8] : 0.0;
    }
}

constexpr int TRSM_LDA = 132;                                   // 128 + 4
constexpr size_t TRSM_MMA_SMEM = sizeof(double) * (64 * PANEL_LDH + 64 * TRSM_LDA);

__global__ void __launch_bounds__(256)
k_trsm_mma(double* __restrict__ H, int64_t strideH, int ldh, int n, int j, const double* __restrict__ Xinv, int nblk,
           const int* __restrict__ fail, const int* __restrict__ active, int nbatch) {
    const int b = batch_index();
    if (b >= nbatch) return;
    if (active && !active[b]) return;
    if (fail[b]) return;
    extern __shared__ __align__(16) double psm[];
    double* Xs = psm;                       // X11, 64 x 64, ld PANEL_LDH
    double* As = psm + 64 * PANEL_LDH;      // 128 rows x 64 columns of A21, ld TRSM_LDA
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = j + 64 + blockIdx.x * 128;
    const int nrow = min(128, n - row0);
    double* Hb = H + (int64_t)b * strideH;
    const double* Xb = Xinv + ((int64_t)b * nblk + j / 64) * 64 * 64;
    for (int q = tid; q < 64 * 64; q += 256) Xs[(q >> 6) * PANEL_LDH + (q & 63)] = Xb[q];
    for (int q = tid; q < 64 * 128; q += 256) {
        const int c = q >> 7, r = q & 127;
        As[c * TRSM_LDA + r] = (r < nrow) ? Hb[(int64_t)(j + c) * ldh + row0 + r] : 0.0;
    }
    __syncthreads();
    // warp w: rows [16w, 16w+16), all 64 columns: C[r][c] = sum_{m <= c} A[r][m] X[c][m]
    const int fr = lane >> 2, fk = lane & 3;
    double acc[2][8][2];
#pragma unroll
    for (int rt = 0; rt < 2; ++rt)
#pragma unroll
        for (int ct = 0; ct < 8; ++ct) acc[rt][ct][0] = acc[rt][ct][1] = 0.0;
#pragma unroll
    for (int kk = 0; kk < 64; kk += 4) {
        const double a0 = As[(kk + fk) * TRSM_LDA + warp * 16 + fr];
        const double a1 = As[(kk + fk) * TRSM_LDA + warp * 16 + 8 + fr];
#pragma unroll
        for (int ct = 0; ct < 8; ++ct) {
            if (kk < (ct + 1) * 8) {        // X is lower triangular: X[c][m] = 0 for m > c
                const double xb = Xs[(kk + fk) * PANEL_LDH + ct * 8 + fr];
                dmma884(acc[0][ct][0], acc[0][ct][1], a0, xb);
                dmma884(acc[1][ct][0], acc[1][ct][1], a1, xb);
            }
        }
    }
#pragma unroll
    for (int rt = 0; rt < 2; ++rt) {
        const int r = warp * 16 + rt * 8 + fr;
        if (r < nrow) {
#pragma unroll
            for (int ct = 0; ct < 8; ++ct)
#pragma unroll
                for (int e = 0; e < 2; ++e) Hb[(int64_t)(j + ct * 8 + 2 * fk + e) * ldh + row0 + r] = acc[rt][ct][e];
        }
    }
}

// ---------------------------------------------------------------------------
// One panel step of the blocked Cholesky as ONE launch, for a handful of large problems (C5: one n = 4096 problem,
// where the factorisation is a serial chain of 64 panel steps and every kernel boundary on it is pure latency):
//   CTA 0 of a problem     : applies the pending update of the panel group to its 64 x 64 diagonal block
//                            (T -= P_D P_D', P_D = L[jm:jm+64, j0:jm]), factors it (as k_potrf_diag_mma), writes L11 and
//                            X11 = L11^-1 and releases the problem's flag;
//   CTAs 1.. (128 rows each): load their rows of the panel, apply the same pending update
//                            (A21 -= P_R P_D', P_R = L[rows, j0:jm]) WHILE CTA 0 factors, wait for the flag, then
//                            L21 = A21 X11' (as k_trsm_mma).
// This replaces thin update -> diagonal block -> TRSM (three launches, 61 us per link on C5) by one launch whose
// critical path is the diagonal block alone.  The waiting CTAs spin on a flag in global memory: every CTA of the
// launch must be resident, so the host uses this kernel only when the whole grid fits the machine at one CTA per SM.
// flags[b] receives `gen` (a value never used before for this handle).
// ---------------------------------------------------------------------------
constexpr int PANEL_UPD_KT = 16;
constexpr size_t PANEL_FUSED_SMEM = TRSM_MMA_SMEM + sizeof(double) * PANEL_UPD_KT * (TRSM_LDA + PANEL_LDH);
constexpr int PANEL_DIAG_KT = 128;      // the diagonal CTA is the critical path: its pending update in at most two shots
static_assert(PANEL_FUSED_SMEM >= POTRF_MMA_SMEM + sizeof(double) * PANEL_DIAG_KT * PANEL_LDH, "diagonal CTA fits too");

// The pending update of the 64 x 64 diagonal block, T -= P_D P_D' (lower triangle), P_D = L[j:j+64, j0:j]: all the
// loads of up to 128 columns of P_D go out at once (this CTA is the serial link of the whole factorisation), the
// eight warps split the block 4 x 2.
__device__ __forceinline__ void panel_diag_pending_update(double* Hs, const double* __restrict__ Hb, int ldh, int n, int j0,
                                                          int j, double* PD, int tid) {
    const int lane = tid & 31, warp = tid >> 5, fr = lane >> 2, fk = lane & 3;
    const int K = j - j0;
    if (K <= 0) return;
    const int rbase = (warp & 3) * 16, cbase = (warp >> 2) * 32;
    const bool needed = rbase + 15 >= cbase;          // some entry of the warp's 16 x 32 piece is on or below the diagonal
    double acc[2][4][2];
#pragma unroll
    for (int rt = 0; rt < 2; ++rt)
#pragma unroll
        for (int ct = 0; ct < 4; ++ct) acc[rt][ct][0] = acc[rt][ct][1] = 0.0;
    for (int k0 = 0; k0 < K; k0 += PANEL_DIAG_KT) {
        const int kc = min(PANEL_DIAG_KT, K - k0);
        __syncthreads();
        for (int q = tid; q < kc * 64; q += 256) {
            const int kk = q >> 6, c = q & 63;
            PD[kk * PANEL_LDH + c] = (j + c < n) ? Hb[(int64_t)(j0 + k0 + kk) * ldh + j + c] : 0.0;
        }
        __syncthreads();
        if (needed) {
#pragma unroll 4
            for (int kk = 0; kk < kc; kk += 4) {
                const double a0 = PD[(kk + fk) * PANEL_LDH + rbase + fr];
                const double a1 = PD[(kk + fk) * PANEL_LDH + rbase + 8 + fr];
#pragma unroll
                for (int ct = 0; ct < 4; ++ct) {
                    const double xb = PD[(kk + fk) * PANEL_LDH + cbase + ct * 8 + fr];
                    dmma884(acc[0][ct][0], acc[0][ct][1], a0, xb);
                    dmma884(acc[1][ct][0], acc[1][ct][1], a1, xb);
                }
            }
        }
    }
#pragma unroll
    for (int rt = 0; rt < 2; ++rt) {
        const int r = rbase + rt * 8 + fr;
#pragma unroll
        for (int ct = 0; ct < 4; ++ct)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = cbase + ct * 8 + 2 * fk + e;
                if (r >= c) Hs[c * PANEL_LDH + r] -= acc[rt][ct][e];
            }
    }
    __syncthreads();
}

// T (column-major in shared memory, ld LDT, rows [0, R)) -= P_R P_D' over the K = jm - j0 pending columns; warp w
// owns rows [16w, 16w + 16).  LOWER: only entries on or below the diagonal of T are touched.
template <int LDT, bool LOWER>
__device__ __forceinline__ void panel_pending_update(double* T, int R, const double* __restrict__ Hb, int ldh, int n,
                                                     int row0, int j0, int jm, double* PA, double* PD, int tid) {
    const int lane = tid & 31, warp = tid >> 5, fr = lane >> 2, fk = lane & 3;
    const int K = jm - j0;
    if (K <= 0) return;
    const int nrow = min(R, n - row0);
    double acc[2][8][2];
#pragma unroll
    for (int rt = 0; rt < 2; ++rt)
#pragma unroll
        for (int ct = 0; ct < 8; ++ct) acc[rt][ct][0] = acc[rt][ct][1] = 0.0;
    for (int k0 = 0; k0 < K; k0 += PANEL_UPD_KT) {
        __syncthreads();
        for (int q = tid; q < PANEL_UPD_KT * 128; q += 256) {
            const int kk = q >> 7, r = q & 127;
            PA[kk * TRSM_LDA + r] = (r < nrow) ? Hb[(int64_t)(j0 + k0 + kk) * ldh + row0 + r] : 0.0;
        }
        for (int q = tid; q < PANEL_UPD_KT * 64; q += 256) {
            const int kk = q >> 6, c = q & 63;
            PD[kk * PANEL_LDH + c] = (jm + c < n) ? Hb[(int64_t)(j0 + k0 + kk) * ldh + jm + c] : 0.0;
        }
        __syncthreads();
        if (warp * 16 < R) {
#pragma unroll
            for (int kk = 0; kk < PANEL_UPD_KT; kk += 4) {
                const double a0 = PA[(kk + fk) * TRSM_LDA + warp * 16 + fr];
                const double a1 = PA[(kk + fk) * TRSM_LDA + warp * 16 + 8 + fr];
#pragma unroll
                for (int ct = 0; ct < 8; ++ct) {
                    const double xb = PD[(kk + fk) * PANEL_LDH + ct * 8 + fr];
                    dmma884(acc[0][ct][0], acc[0][ct][1], a0, xb);
                    dmma884(acc[1][ct][0], acc[1][ct][1], a1, xb);
                }
            }
        }
    }
#pragma unroll
    for (int rt = 0; rt < 2; ++rt) {
        const int r = warp * 16 + rt * 8 + fr;
        if (r < R) {
#pragma unroll
            for (int ct = 0; ct < 8; ++ct)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = ct * 8 + 2 * fk + e;
                    if (!LOWER || r >= c) T[c * LDT + r] -= acc[rt][ct][e];
                }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
k_panel_fused(double* __restrict__ H, int64_t strideH, int ldh, int n, int j0, int j, double* __restrict__ Xinv, int nblk,
              int* __restrict__ fail, const int* __restrict__ active, int* __restrict__ flags, int gen, int nbatch) {
    const int b = batch_index();
    if (b >= nbatch) return;
    if (active && !active[b]) return;
    extern __shared__ __align__(16) double psm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* Hb = H + (int64_t)b * strideH;
    if (blockIdx.x == 0) {
        // ---- diagonal block (layout of k_potrf_diag_mma, the update staging behind it)
        __shared__ int sfail;
        double* Hs = psm;
        double* Xs = Hs + 64 * PANEL_LDH;
        double* Dinv = Xs + f2_xsize(8);
        uint2* desc = reinterpret_cast<uint2*>(Dinv + 192);
        double* PD = reinterpret_cast<double*>(desc + f2_trail_base(8, 8));
        double* Hd = Hb + (int64_t)j * ldh + j;
        const int jb = min(64, n - j);
        if (tid == 0) sfail = 0;
        for (int q = tid; q < 64 * 64; q += 256) {
            const int c = q >> 6, r = q & 63;
            double v = 0.0;
            if (r < jb && c < jb) v = (r >= c) ? Hd[(int64_t)c * ldh + r] : 0.0;
            else if (r == c) v = 1.0;
            Hs[c * PANEL_LDH + r] = v;
        }
        for (int q = tid; q < f2_xsize(8) + 192; q += 256) Xs[q] = 0.0;
        f2_build_trail(desc, 8, PANEL_LDH, 64 * PANEL_LDH, tid, 256);
        __syncthreads();
        panel_diag_pending_update(Hs, Hb, ldh, n, j0, j, PD, tid);
        const int ok = f2_chol_inv<8, true>(Hs, Xs, Dinv, desc, 8, PANEL_LDH, &sfail, lane, warp);
        if (!ok) {
            if (tid == 0) fail[b] = 1;
        } else {
            double* Xb = Xinv + ((int64_t)b * nblk + j / 64) * 64 * 64;
            for (int q = tid; q < 64 * 64; q += 256) {
                const int c = q >> 6, r = q & 63;
                if (r >= c && r < jb) Hd[(int64_t)c * ldh + r] = Hs[c * PANEL_LDH + r];
                const int cb = c >> 3;
                Xb[q] = (r >= cb * 8) ? Xs[f2_xbase(8, cb) + (c & 7) * f2_xld(8, cb) + r - cb * 8] : 0.0;
            }
        }
        __threadfence();                       // L11, X11 (and fail) before the flag, for every CTA of the launch
        __syncthreads();
        if (tid == 0) atomicExch(flags + b, gen);
        return;
    }
    // ---- 128 rows of the sub-panel
    if (fail[b]) return;                       // an earlier panel of this problem failed
    double* Xs = psm;                          // X11, 64 x 64, ld PANEL_LDH
    double* As = psm + 64 * PANEL_LDH;         // 128 rows x 64 columns of A21, ld TRSM_LDA
    double* PA = As + 64 * TRSM_LDA;
    double* PD = PA + PANEL_UPD_KT * TRSM_LDA;
    const int row0 = j + 64 + ((int)blockIdx.x - 1) * 128;
    const int nrow = min(128, n - row0);
    for (int q = tid; q < 64 * 128; q += 256) {
        const int c = q >> 7, r = q & 127;
        As[c * TRSM_LDA + r] = (r < nrow) ? Hb[(int64_t)(j + c) * ldh + row0 + r] : 0.0;
    }
    __syncthreads();
    panel_pending_update<TRSM_LDA, false>(As, 128, Hb, ldh, n, row0, j0, j, PA, PD, tid);
    if (tid == 0) {
        while (atomicAdd(flags + b, 0) != gen) __nanosleep(200);
        __threadfence();
    }
    __syncthreads();
    if (__ldcg(fail + b)) return;              // this panel's diagonal block is not positive definite
    const double* Xb = Xinv + ((int64_t)b * nblk + j / 64) * 64 * 64;
    for (int q = tid; q < 64 * 64; q += 256) Xs[(q >> 6) * PANEL_LDH + (q & 63)] = __ldcg(Xb + q);
    __syncthreads();
    const int fr = lane >> 2, fk = lane & 3;
    double acc[2][8][2];
#pragma unroll
    for (int rt = 0; rt < 2; ++rt)
#pragma unroll
        for (int ct = 0; ct < 8; ++ct) acc[rt][ct][0] = acc[rt][ct][1] = 0.0;
#pragma unroll
    for (int kk = 0; kk < 64; kk += 4) {
        const double a0 = As[(kk + fk) * TRSM_LDA + warp * 16 + fr];
        const double a1 = As[(kk + fk) * TRSM_LDA + warp * 16 + 8 + fr];
#pragma unroll
        for (int ct = 0; ct < 8; ++ct) {
            if (kk < (ct + 1) * 8) {            // X is lower triangular: X[c][m] = 0 for m > c
                const double xb = Xs[(kk + fk) * PANEL_LDH + ct * 8 + fr];
                dmma884(acc[0][ct][0], acc[0][ct][1], a0, xb);
                dmma884(acc[1][ct][0], acc[1][ct][1], a1, xb);
            }
        }
    }
#pragma unroll
    for (int rt = 0; rt < 2; ++rt) {
        const int r = warp * 16 + rt * 8 + fr;
        if (r < nrow) {
#pragma unroll
            for (int ct = 0; ct < 8; ++ct)
#pragma unroll
                for (int e = 0; e < 2; ++e) Hb[(int64_t)(j + ct * 8 + 2 * fk + e) * ldh + row0 + r] = acc[rt][ct][e];
        }
    }
}

// ---------------------------------------------------------------------------
// Triangular solves with the factor and the inverted 64 x 64 diagonal blocks kept by k_potrf_diag_mma:
//   k_trsv_blk_fwd:  X <- L^-1 X        k_trsv_blk_bwd:  X <- L^-T X        (in place, n x nrhs, ld = ldx)
// One CTA (NT = 256 threads, 1024 when the batch is too small to fill the machine) per (rhs, problem), the right-
// hand side in shared memory.  Per 64-block: a 64 x 64 gemv
// with the inverted block (no serial substitution chain) and a gemv with the panel below it.
// grid (nrhs, batch), dynamic smem = (n + 64) * 8 (+ (NT / 256) * n * 8 for the forward kernel with NT > 256).
// ---------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT)
k_trsv_blk_fwd(const double* __restrict__ L, int64_t strideL, int ldl, int n, const double* __restrict__ Xinv,
               int nblk, int blk0, double* __restrict__ X, int64_t strideX, int ldx, const int* __restrict__ active, int nbatch) {
    const int b = batch_index();
    if (b >= nbatch) return;
    if (active && !active[b]) return;
    extern __shared__ double tsm[];
    double* xs = tsm;
    double* ys = tsm + n;
    const double* Lb = L + (int64_t)b * strideL;
    const double* Xi = Xinv + ((int64_t)b * nblk + blk0) * 4096;
    double* xg = X + (int64_t)b * strideX + (int64_t)blockIdx.x * ldx;
    const int tid = threadIdx.x;
    for (int i = tid; i < n; i += NT) xs[i] = xg[i];
    __syncthreads();
    for (int jb = 0; jb < n; jb += 64) {
        const int w = min(64, n - jb);
        const double* Xb = Xi + (int64_t)(jb >> 6) * 4096;
        {   // ys = Xbb xs[jb .. jb+w): NT/64 threads per row, 64/(NT/64) columns each
            constexpr int PARTS = NT / 64, CPP = 64 / PARTS;
            const int r = tid / PARTS, q = tid % PARTS;
            double acc = 0.0;
#pragma unroll 4
            for (int c = q * CPP; c < q * CPP + CPP; ++c)
                if (c <= r && c < w) acc = fma(Xb[c * 64 + r], xs[jb + c], acc);
#pragma unroll
            for (int o = 1; o < PARTS; o <<= 1) acc += __shfl_xor_sync(FULL_MASK, acc, o);
            if (q == 0) ys[r] = acc;
        }
        __syncthreads();
        if constexpr (NT == 256) {
            for (int r = jb + 64 + tid; r < n; r += NT) {    // rows below: xs[r] -= L[r, jb..jb+64) ys
                const double* lp = Lb + (int64_t)jb * ldl + r;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
                for (int c = 0; c < 64; c += 4) {
                    a0 = fma(lp[(int64_t)c * ldl], ys[c], a0);
                    a1 = fma(lp[(int64_t)(c + 1) * ldl], ys[c + 1], a1);
                    a2 = fma(lp[(int64_t)(c + 2) * ldl], ys[c + 2], a2);
                    a3 = fma(lp[(int64_t)(c + 3) * ldl], ys[c + 3], a3);
                }
                xs[r] -= (a0 + a1) + (a2 + a3);
            }
        } else {
            // few problems: the 64 columns are split over NT/256 thread groups as well (16 loads in flight per thread,
            // one round for 256 rows); partial sums meet in shared memory
            constexpr int CS = NT / 256, CW = 64 / CS;
            const int rl = tid & 255, cp = tid >> 8;
            double* part = tsm + n + 64;                     // [CS][n]
            for (int r = jb + 64 + rl; r < n; r += 256) {
                const double* lp = Lb + (int64_t)(jb + cp * CW) * ldl + r;
                const double* yp = ys + cp * CW;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                for (int c = 0; c < CW; c += 4) {
                    a0 = fma(lp[(int64_t)c * ldl], yp[c], a0);
                    a1 = fma(lp[(int64_t)(c + 1) * ldl], yp[c + 1], a1);
                    a2 = fma(lp[(int64_t)(c + 2) * ldl], yp[c + 2], a2);
                    a3 = fma(lp[(int64_t)(c + 3) * ldl], yp[c + 3], a3);
                }
                part[cp * n + r] = (a0 + a1) + (a2 + a3);
            }
            __syncthreads();
            for (int r = jb + 64 + tid; r < n; r += NT) {
                double a = 0.0;
#pragma unroll
                for (int q = 0; q < CS; ++q) a += part[q * n + r];
                xs[r] -= a;
            }
        }
        if (tid < w) xs[jb + tid] = ys[tid];
        __syncthreads();
    }
    for (int i = tid; i < n; i += NT) xg[i] = xs[i];
}

template <int NT>
__global__ void __launch_bounds__(NT)
k_trsv_blk_bwd(const double* __restrict__ L, int64_t strideL, int ldl, int n, const double* __restrict__ Xinv,
               int nblk, int blk0, double* __restrict__ X, int64_t strideX, int ldx, const int* __restrict__ active, int nbatch) {
    const int b = batch_index();
    if (b >= nbatch) return;
    if (active && !active[b]) return;
    extern __shared__ double tsm[];
    double* xs = tsm;
    double* ts = tsm + n;
    const double* Lb = L + (int64_t)b * strideL;
    const double* Xi = Xinv + ((int64_t)b * nblk + blk0) * 4096;
    double* xg = X + (int64_t)b * strideX + (int64_t)blockIdx.x * ldx;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n; i += NT) xs[i] = xg[i];
    __syncthreads();
    for (int jb = ((n - 1) >> 6) << 6; jb >= 0; jb -= 64) {
        const int w = min(64, n - jb);
        const double* Xb = Xi + (int64_t)(jb >> 6) * 4096;
        // ts[c] = xs[jb+c] - sum_{r >= jb+64} L[r, jb+c] xs[r]: warp per column, lanes over the rows below
        for (int c = warp; c < 64; c += NT / 32) {
            double a0 = 0.0, a1 = 0.0;
            if (c < w) {
                const double* col = Lb + (int64_t)(jb + c) * ldl;
                int r = jb + 64 + lane;
                for (; r + 32 < n; r += 64) {
                    a0 = fma(col[r], xs[r], a0);
                    a1 = fma(col[r + 32], xs[r + 32], a1);
                }
                if (r < n) a0 = fma(col[r], xs[r], a0);
            }
            const double acc = warp_sum(a0 + a1);
            if (lane == 0) ts[c] = (c < w) ? xs[jb + c] - acc : 0.0;
        }
        __syncthreads();
        {   // xs[jb+c] = sum_{r >= c} Xbb[r, c] ts[r]: NT/64 threads per column
            constexpr int PARTS = NT / 64, RPP = 64 / PARTS;
            const int c = tid / PARTS, q = tid % PARTS;
            double acc = 0.0;
#pragma unroll 4
            for (int r = q * RPP; r < q * RPP + RPP; ++r)
                if (r >= c && r < w) acc = fma(Xb[c * 64 + r], ts[r], acc);
#pragma unroll
            for (int o = 1; o < PARTS; o <<= 1) acc += __shfl_xor_sync(FULL_MASK, acc, o);
            if (q == 0 && c < w) xs[jb + c] = acc;
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += NT) xg[i] = xs[i];
}

}  // namespace socp
