// linalg.cuh -- batched dense FP64 kernels of the tiled (global-memory) path:
// gemv and SYRK on the DMMA tensor pipe (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4); the
// SYRK is also the trailing update of the blocked right-looking Cholesky whose panel
// kernels and triangular solves live in panel_mma.cuh.  Column-major everywhere,
// batch = slowest index.
//
// They replace the LAPACK/BLAS calls of the reference's dense back end:
// mul! (src/densesolver.jl:42-43,49-50,66,73,83-86), cholesky! (:47,:51) and
// ldiv! (:48,:75).
#pragma once
#include "common.cuh"

namespace socp {

// ------------------------------------------------------------------ cp.async
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = pred ? 16 : 0;   // src-size 0 => zero fill, no global read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// dmma884 (mma.sync m8n8k4 f64) lives in common.cuh

// ---------------------------------------------------------------------------
// gemv, transposed:  out[c] = sum_r M[r,c] x[r]  (+ c1*v1[c] + c2*v2[c] [+ out[c]])
// grid (ceil(cols/8), batch), block 256: one warp per column, lanes over rows.
// ---------------------------------------------------------------------------
struct GemvEpi {
    const double* v1; double c1; int64_t s1;
    const double* v2; double c2; int64_t s2;
    int accumulate;
};

__global__ void __launch_bounds__(256)
k_gemv_t(const double* __restrict__ M, int64_t strideM, int ld, int rows, int cols,
         const double* __restrict__ x, int64_t strideX, double* __restrict__ out, int64_t strideOut,
         double alpha, GemvEpi epi, const int* __restrict__ active, const uint8_t* __restrict__ flag, int nbatch) {
    const int b = batch_index();
    if (b >= nbatch) return;
    if (active && !active[b]) return;
    if (flag && !flag[b]) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 8 + warp;
    if (c >= cols) return;
    const double* col = M + (int64_t)b * strideM + (int64_t)c * ld;
    const double* xb = x + (int64_t)b * strideX;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int r = lane;
    for (; r + 96 < rows; r += 128) {
        a0 = fma(col[r], xb[r], a0);
        a1 = fma(col[r + 32], xb[r + 32], a1);
        a2 = fma(col[r + 64], xb[r + 64], a2);
        a3 = fma(col[r + 96], xb[r + 96], a3);
    }
    for (; r < rows; r += 32) a0 = fma(col[r], xb[r], a0);
    double acc = alpha * warp_sum((a0 + a1) + (a2 + a3));
    if (lane == 0) {
        if (epi.v1) acc += epi.c1 * epi.v1[(int64_t)b * epi.s1 + c];
        if (epi.v2) acc += epi.c2 * epi.v2[(int64_t)b * epi.s2 + c];
        double* o = out + (int64_t)b * strideOut + c;
        if (epi.accumulate) acc += *o;
        *o = acc;
    }
}

// ---------------------------------------------------------------------------
// gemv, not transposed:  out[r] = sum_c M[r,c] x[c]  (+ epilogue as above)
// grid (ceil(rows/32), batch), block (32, 8): lane = row, y = column phase.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_gemv_n(const double* __restrict__ M, int64_t strideM, int ld, int rows, int cols,
         const double* __restrict__ x, int64_t strideX, double* __restrict__ out, int64_t strideOut,
         double alpha, GemvEpi epi, const int* __restrict__ active, const uint8_t* __restrict__ flag, int nbatch) {
    const int b = batch_index();
    if (b >= nbatch) return;
    if (active && !active[b]) return;
    if (flag && !flag[b]) return;
    __shared__ double part[8][33];
    const int lane = threadIdx.x, ph = threadIdx.y;
    const int r = blockIdx.x * 32 + lane;
    const double* Mb = M + (int64_t)b * strideM;
    const double* xb = x + (int64_t)b * strideX;
    double a0 = 0.0, a1 = 0.0;
    if (r < rows) {
        int c = ph;
        for (; c + 8 < cols; c += 16) {
            a0 = fma(Mb[r + (int64_t)c * ld], xb[c], a0);
            a1 = fma(Mb[r + (int64_t)(c + 8) * ld], xb[c + 8], a1);
        }
        for (; c < cols; c += 8) a0 = fma(Mb[r + (int64_t)c * ld], xb[c], a0);
    }
    part[ph][lane] = a0 + a1;
    __syncthreads();
    if (ph == 0 && r < rows) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) acc += part[q][lane];
        acc *= alpha;
        if (epi.v1) acc += epi.c1 * epi.v1[(int64_t)b * epi.s1 + r];
        if (epi.v2) acc += epi.c2 * epi.v2[(int64_t)b * epi.s2 + r];
        double* o = out + (int64_t)b * strideOut + r;
        if (epi.accumulate) acc += *o;
        *o = acc;
    }
}

// ---------------------------------------------------------------------------
// SYRK on the FP64 tensor pipe.
//   KMAJOR = true : A is K x N column-major (ld = lda), C[i,j] = sum_r A[r,i] A[r,j]
//                   (H = Gt' Gt, reference src/densesolver.jl:42-43)
//   KMAJOR = false: A is N x K column-major,            C[i,j] = sum_c A[i,c] A[j,c]
//                   (trailing update of the blocked Cholesky)
//   C = beta*C + alpha*(...) (+ addC where addFlag[b]) on the lower-triangle TILES
//   (of a diagonal tile the warp sub-tiles on or below the diagonal are written).
// Tile BT x BT per CTA, NW x NW warps each owning a (BT/NW)^2 sub-tile made of
// m8n8k4 DMMAs; operands staged through shared memory with a 3-stage cp.async
// ring.  Requirements: A 16-byte aligned, lda even; KMAJOR: K % KT == 0 with the
// pad rows of A zero.
// grid (ntiles*(ntiles+1)/2, batch); with first_col_only grid (ntiles, batch): only the first block column of C
// (the next panel of a blocked Cholesky whose trailing update is delayed).
// ---------------------------------------------------------------------------
template <int BT, int NW, int KT, bool KMAJOR>
struct SyrkCfg {
    static constexpr int THREADS = NW * NW * 32;
    static constexpr int WT = BT / NW;          // warp tile edge
    static constexpr int MT = WT / 8;           // 8x8 mma tiles per warp-tile edge
    static constexpr int STAGES = 3;
    // KMAJOR: tile stored [col][r], r contiguous, row stride KT+4 (== 4 mod 16 for KT=16/32)
    // else  : tile stored [c][i],  i contiguous, row stride BT+4
    static constexpr int TS_ROWS = KMAJOR ? BT : KT;
    static constexpr int TS_LD = KMAJOR ? (KT + 4) : (BT + 4);
    static constexpr int TILE_DOUBLES = TS_ROWS * TS_LD;
    static constexpr size_t SMEM = (size_t)STAGES * 2 * TILE_DOUBLES * sizeof(double);
};

template <int BT, int NW, int KT, bool KMAJOR>
__global__ void __launch_bounds__(NW * NW * 32)
k_syrk(const double* __restrict__ A, int64_t strideA, int lda, int N, int K,
       double* __restrict__ C, int64_t strideC, int ldc, double alpha, double beta,
       const double* __restrict__ addC, int64_t strideAdd, int ldadd, const uint8_t* __restrict__ addFlag,
       const int* __restrict__ active, int first_col_only, int nbatch) {
    using Cfg = SyrkCfg<BT, NW, KT, KMAJOR>;
    const int b = batch_index();
    if (b >= nbatch) return;
    if (active && !active[b]) return;
    // lower-triangle tile index -> (ti, tj), ti >= tj; first_col_only: the tiles (t, 0) of the first block column
    int t = blockIdx.x;
    int ti, tj;
    if (first_col_only) { ti = t; tj = 0; }
    else {
        ti = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
        while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
        while (ti * (ti + 1) / 2 > t) --ti;
        tj = t - ti * (ti + 1) / 2;
    }
    const int i0 = ti * BT, j0 = tj * BT;
    const bool diag = (ti == tj);

    extern __shared__ __align__(16) double smem[];
    const double* Ab = A + (int64_t)b * strideA;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // Warp (wr, wc) owns the (BT/NW)^2 sub-tile at (wr, wc).  On a diagonal tile only sub-tiles with wr >= wc are
    // needed (the factorisation and the solves read the lower triangle): those NW(NW+1)/2 sub-tiles go to the first
    // warps -- warp ids map to the four SM sub-partitions round-robin, so the DMMA work stays balanced -- and the
    // remaining warps only help with the loads.
    int wr = warp / NW, wc = warp % NW;
    bool compute = true;
    if (diag) {
        constexpr int NACT = NW * (NW + 1) / 2;
        if (warp < NACT) {
            int a = 0;
            while ((a + 1) * (a + 2) / 2 <= warp) ++a;
            wr = a;
            wc = warp - a * (a + 1) / 2;
        } else compute = false;
    }
    const int nk = (K + KT - 1) / KT;

    auto tileA = [&](int s) { return smem + (size_t)s * 2 * Cfg::TILE_DOUBLES; };
    auto tileB = [&](int s) { return smem + (size_t)s * 2 * Cfg::TILE_DOUBLES + Cfg::TILE_DOUBLES; };

    auto load_stage = [&](int s, int kb) {
        const int r0 = kb * KT;
        if (KMAJOR) {
            // BT columns x KT rows, 16B chunks along r
            constexpr int CH = KT / 2;
            for (int q = tid; q < BT * CH; q += Cfg::THREADS) {
                const int col = q / CH, ch = q % CH;
                const int r = r0 + ch * 2;
                {
                    const int gc = i0 + col;
                    const bool ok = (gc < N) && (r < K);
                    cp_async16(tileA(s) + col * Cfg::TS_LD + ch * 2, Ab + (int64_t)gc * lda + r, ok);
                }
                if (!diag) {
                    const int gc = j0 + col;
                    const bool ok = (gc < N) && (r < K);
                    cp_async16(tileB(s) + col * Cfg::TS_LD + ch * 2, Ab + (int64_t)gc * lda + r, ok);
                }
            }
        } else {
            // KT columns (c) x BT rows (i), 16B chunks along i
            constexpr int CH = BT / 2;
            for (int q = tid; q < KT * CH; q += Cfg::THREADS) {
                const int c = q / CH, ch = q % CH;
                const int gc = r0 + c;
                {
                    const int gi = i0 + ch * 2;
                    const bool ok = (gc < K) && (gi < N);
                    cp_async16(tileA(s) + c * Cfg::TS_LD + ch * 2, Ab + (int64_t)gc * lda + gi, ok);
                }
                if (!diag) {
                    const int gi = j0 + ch * 2;
                    const bool ok = (gc < K) && (gi < N);
                    cp_async16(tileB(s) + c * Cfg::TS_LD + ch * 2, Ab + (int64_t)gc * lda + gi, ok);
                }
            }
        }
    };

    double acc[Cfg::MT][Cfg::MT][2];
#pragma unroll
    for (int a = 0; a < Cfg::MT; ++a)
#pragma unroll
        for (int c = 0; c < Cfg::MT; ++c) acc[a][c][0] = acc[a][c][1] = 0.0;

    // prologue
#pragma unroll
    for (int s = 0; s < Cfg::STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    for (int kb = 0; kb < nk; ++kb) {
        cp_async_wait<Cfg::STAGES - 2>();
        __syncthreads();
        {   // prefetch stage kb + STAGES-1 into the slot consumed at iteration kb-1
            const int nxt = kb + Cfg::STAGES - 1;
            if (nxt < nk) load_stage(nxt % Cfg::STAGES, nxt);
            cp_async_commit();
        }
        const int s = kb % Cfg::STAGES;
        const double* ta = tileA(s);
        const double* tb = diag ? tileA(s) : tileB(s);
        const int fr = lane >> 2, fk = lane & 3;
        if (compute)
#pragma unroll
        for (int kk = 0; kk < KT; kk += 4) {
            double af[Cfg::MT], bf[Cfg::MT];
#pragma unroll
            for (int a = 0; a < Cfg::MT; ++a) {
                const int row = wr * Cfg::WT + a * 8 + fr;
                af[a] = KMAJOR ? ta[row * Cfg::TS_LD + kk + fk] : ta[(kk + fk) * Cfg::TS_LD + row];
            }
#pragma unroll
            for (int c = 0; c < Cfg::MT; ++c) {
                const int col = wc * Cfg::WT + c * 8 + fr;
                bf[c] = KMAJOR ? tb[col * Cfg::TS_LD + kk + fk] : tb[(kk + fk) * Cfg::TS_LD + col];
            }
#pragma unroll
            for (int a = 0; a < Cfg::MT; ++a)
#pragma unroll
                for (int c = 0; c < Cfg::MT; ++c) dmma884(acc[a][c][0], acc[a][c][1], af[a], bf[c]);
        }
    }
    cp_async_wait<0>();

    // epilogue: C fragment layout of m8n8k4: row = lane/4, cols = 2*(lane%4) + {0,1}
    if (!compute) return;
    double* Cb = C + (int64_t)b * strideC;
    const bool add = addC && (!addFlag || addFlag[b]);
    const double* Ad = add ? addC + (int64_t)b * strideAdd : nullptr;
#pragma unroll
    for (int a = 0; a < Cfg::MT; ++a)
#pragma unroll
        for (int c = 0; c < Cfg::MT; ++c) {
            const int gi = i0 + wr * Cfg::WT + a * 8 + (lane >> 2);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gj = j0 + wc * Cfg::WT + c * 8 + 2 * (lane & 3) + e;
                if (gi < N && gj < N) {
                    double v = alpha * acc[a][c][e];
                    double* dst = Cb + (int64_t)gj * ldc + gi;
                    if (beta != 0.0) v += beta * (*dst);
                    if (add) v += Ad[(int64_t)gj * ldadd + gi];
                    *dst = v;
                }
            }
        }
}

constexpr int CHOL_NB = 64;      // panel width of the blocked Cholesky (panel kernels: panel_mma.cuh)

}  // namespace socp
