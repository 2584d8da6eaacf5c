// syrk_tma.cu -- the Gram product of the tiled path, H = Gt' Gt (reference src/densesolver.jl:42-43: G'W^-2 G with
// Gt = W^-1 G), on the FP64 tensor pipe (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4) with the operands fed by TMA:
// a tensor map over Gt ([batch][n columns][K rows], rows contiguous), one elected thread issuing two
// cp.async.bulk.tensor.3d tile loads per k-step (SASS UTMALDG) into a 4-stage shared-memory ring, completion
// signalled through mbarriers (SASS SYNCS).  The 512 compute threads issue no load instructions for the operands.
//
// Shared-memory tile: 128 columns x 20 rows of Gt, rows contiguous -- a pitch of 20 doubles (== 4 mod 8) makes the
// m8n8k4 fragment loads (8 columns x 4 rows per warp) bank-conflict free without swizzling, so the k-step is 20 rows
// (5 DMMA k-blocks) and the TMA box is (20, 128, 1).  Rows beyond K and columns beyond n are zero-filled by the TMA
// unit (out-of-bounds box elements), so K needs no padding to a multiple of 20.
//
// Same tiling, warp layout and epilogue as k_syrk<128, 4, 16, true> in linalg.cuh (the cp.async version, which
// remains for the small-tile and N x K cases); results are bit-identical only up to the order of the k-blocks inside
// a tile, which is the same (ascending), so they are bit-identical.
#include "syrk_tma.cuh"
#include "common.cuh"
#include <cuda.h>
#include <cstdio>

namespace socp {

namespace {

constexpr int TM_BT = 128, TM_NW = 4, TM_KT = 20, TM_STAGES = 4;
constexpr int TM_THREADS = TM_NW * TM_NW * 32;
constexpr int TM_WT = TM_BT / TM_NW, TM_MT = TM_WT / 8;
constexpr int TM_TILE = TM_BT * TM_KT;                     // doubles per operand tile
constexpr size_t TM_SMEM = (size_t)TM_STAGES * 2 * TM_TILE * sizeof(double) + 128;

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2),
          "r"((unsigned)__cvta_generic_to_shared(bar))
        : "memory");
}

__global__ void __launch_bounds__(TM_THREADS, 1)
k_syrk_tma(const __grid_constant__ CUtensorMap mapA, int N, int K, double* __restrict__ C, int64_t strideC, int ldc,
           double alpha, double beta, const double* __restrict__ addC, int64_t strideAdd, int ldadd,
           const uint8_t* __restrict__ addFlag, const int* __restrict__ active, int nbatch) {
    const int b = batch_index();
    if (b >= nbatch) return;
    if (active && !active[b]) return;
    int t = blockIdx.x;
    int ti = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    while (ti * (ti + 1) / 2 > t) --ti;
    const int tj = t - ti * (ti + 1) / 2;
    const int i0 = ti * TM_BT, j0 = tj * TM_BT;
    const bool diag = (ti == tj);

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    __shared__ __align__(8) uint64_t full[TM_STAGES];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // warp (wr, wc) owns a 32 x 32 sub-tile; on a diagonal tile only the 10 sub-tiles on or below the diagonal exist
    // and go to the first 10 warps (spread over the four SM sub-partitions), see k_syrk in linalg.cuh
    int wr = warp / TM_NW, wc = warp % TM_NW;
    bool compute = true;
    if (diag) {
        constexpr int NACT = TM_NW * (TM_NW + 1) / 2;
        if (warp < NACT) {
            int a = 0;
            while ((a + 1) * (a + 2) / 2 <= warp) ++a;
            wr = a;
            wc = warp - a * (a + 1) / 2;
        } else compute = false;
    }
    const int nk = (K + TM_KT - 1) / TM_KT;
    auto tileA = [&](int s) { return tiles + (size_t)s * 2 * TM_TILE; };
    auto tileB = [&](int s) { return tiles + (size_t)s * 2 * TM_TILE + TM_TILE; };
    const unsigned stage_bytes = (unsigned)((diag ? 1 : 2) * TM_TILE * sizeof(double));

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TM_STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int kb) {          // one thread: both operand tiles of k-step kb into slot kb % STAGES
        const int s = kb % TM_STAGES;
        mbar_expect_tx(&full[s], stage_bytes);
        tma_load_3d(tileA(s), &mapA, kb * TM_KT, i0, b, &full[s]);
        if (!diag) tma_load_3d(tileB(s), &mapA, kb * TM_KT, j0, b, &full[s]);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TM_STAGES - 1; ++s)
            if (s < nk) issue(s);
    }

    double acc[TM_MT][TM_MT][2];
#pragma unroll
    for (int a = 0; a < TM_MT; ++a)
#pragma unroll
        for (int c = 0; c < TM_MT; ++c) acc[a][c][0] = acc[a][c][1] = 0.0;

    const int fr = lane >> 2, fk = lane & 3;
    for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % TM_STAGES;
        mbar_wait(&full[s], (unsigned)((kb / TM_STAGES) & 1));
        __syncthreads();                 // every warp is done with the slot that the next load overwrites
        if (tid == 0 && kb + TM_STAGES - 1 < nk) issue(kb + TM_STAGES - 1);
        const double* ta = tileA(s);
        const double* tb = diag ? tileA(s) : tileB(s);
        if (compute) {
#pragma unroll
            for (int kk = 0; kk < TM_KT; kk += 4) {
                double af[TM_MT], bf[TM_MT];
#pragma unroll
                for (int a = 0; a < TM_MT; ++a) af[a] = ta[(wr * TM_WT + a * 8 + fr) * TM_KT + kk + fk];
#pragma unroll
                for (int c = 0; c < TM_MT; ++c) bf[c] = tb[(wc * TM_WT + c * 8 + fr) * TM_KT + kk + fk];
#pragma unroll
                for (int a = 0; a < TM_MT; ++a)
#pragma unroll
                    for (int c = 0; c < TM_MT; ++c) dmma884(acc[a][c][0], acc[a][c][1], af[a], bf[c]);
            }
        }
    }

    if (!compute) return;
    double* Cb = C + (int64_t)b * strideC;
    const bool add = addC && (!addFlag || addFlag[b]);
    const double* Ad = add ? addC + (int64_t)b * strideAdd : nullptr;
#pragma unroll
    for (int a = 0; a < TM_MT; ++a)
#pragma unroll
        for (int c = 0; c < TM_MT; ++c) {
            const int gi = i0 + wr * TM_WT + a * 8 + (lane >> 2);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gj = j0 + wc * TM_WT + c * 8 + 2 * (lane & 3) + e;
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

}  // namespace

bool syrk_tma_supported(const double* A, int64_t strideA, int lda, int N, int K) {
    if (getenv("SOCP_B200_NO_TMA")) return false;
    if (!encode_fn()) return false;
    // tensor-map requirements: 16-byte aligned base and strides
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (lda & 1) || strideA <= 0 || (strideA & 1)) return false;
    return N > 0 && K > 0;
}

cudaError_t syrk_tma_launch(cudaStream_t stream, dim3 grid, const double* A, int64_t strideA, int lda, int N, int K,
                            double* C, int64_t strideC, int ldc, double alpha, double beta, const double* addC,
                            int64_t strideAdd, int ldadd, const uint8_t* addFlag, const int* active, int nbatch) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_syrk_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TM_SMEM);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)N, (cuuint64_t)nbatch};
    const cuuint64_t strides[2] = {(cuuint64_t)lda * sizeof(double), (cuuint64_t)strideA * sizeof(double)};
    const cuuint32_t box[3] = {(cuuint32_t)TM_KT, (cuuint32_t)TM_BT, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = encode_fn()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(A), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    k_syrk_tma<<<grid, TM_THREADS, TM_SMEM, stream>>>(map, N, K, C, strideC, ldc, alpha, beta, addC, strideAdd, ldadd, addFlag,
                                                      active, nbatch);
    return cudaPeekAtLastError();
}

}  // namespace socp
