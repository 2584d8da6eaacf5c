// solver.cu -- host side of libsocp_b200: device shards, the tiled solve driver
// (reference src/solver.jl:40-152 over the batch) and the C ABI of
// include/socp_b200.h.  CUDA runtime only; no torch types, no CPU fallback.
#include "../../include/socp_b200.h"
#include "linalg.cuh"
#include "tiled_kernels.cuh"
#include "fused_v2.cuh"
#include "fused_v3.cuh"
#include "fused_lane.cuh"
#include "cone_batch.cuh"
#include "panel_mma.cuh"
#include "syrk_tma.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using namespace socp;

namespace {

constexpr int POC_CHUNK = 128;

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct CudaErr {
    int code;
    std::string msg;
};

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            char buf_[512];                                                                       \
            snprintf(buf_, sizeof buf_, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,      \
                     cudaGetErrorString(e_));                                                     \
            throw CudaErr{(int)e_, buf_};                                                         \
        }                                                                                         \
    } while (0)

struct UsageErr {
    int code;
    std::string msg;
};

// One device's contiguous shard of the batch.
struct Shard {
    int device = 0;
    int64_t first = 0;     // first problem of the shard in the caller's batch
    int batch = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t up_stream = nullptr, down_stream = nullptr;     // H2D / D2H legs of the pipelined one-shot solve
    cudaStream_t alt_stream = nullptr;                           // second compute stream: consecutive chunks backfill
    std::vector<cudaEvent_t> pipe_ev;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t split_ev[2] = {nullptr, nullptr};                // fork / join of the two-stream factor
    int* d_panel_flags = nullptr;                                // [batch] release flags of k_panel_fused
    int panel_gen = 0;                                           // ... and the value the next launch releases with
    std::vector<void*> allocs;
    Ws w{};
    // layout (device copies)
    int *d_kind = nullptr, *d_offs = nullptr, *d_dim = nullptr;
    BLayout bl{};          // batch-wide lane-group view of the cones (cone_batch.cuh)
    bool bl_ok = false;    // every second-order cone has dimension <= 128
    bool bl_warp = false;  // ... and the cones of one problem fit in one warp (nsoc * lpc <= 32)
    int bl_sw = 32;        // lanes per problem in the per-problem reductions (power of two >= nsoc * lpc)
    // owned copies of the problem data
    double *d_c = nullptr, *d_A = nullptr, *d_b = nullptr, *d_G = nullptr, *d_h = nullptr;
    uint8_t* d_sing = nullptr;
    int* h_nactive = nullptr;   // pinned
    int64_t launches = 0;
    bool have_data = false, have_scaling = false, have_factor = false;
    bool any_sing = false;
    bool sharedA = false, sharedG = false;
    int threads = 256;     // CTA size of the per-problem cone kernels
    socp_timings tim{};
    F2Plan fused2{};
    // second-generation whole-solve kernel (fused_v3.cuh): planned per data set, from the row pattern of G
    F3Plan fused3{};
    // lane-per-problem whole-solve kernel (fused_lane.cuh): tiny problems, large batches
    FLPlan lane{};
    bool f3_planned = false;     // fused3 belongs to the data now resident
    bool f3_dense = false;       // ... and was planned with every row dense (fallback after a pattern violation)
    int* d_f3_tables = nullptr;  // capacity 2k + n + 1 ints
    int* d_rowcol = nullptr;     // k ints: row pattern found on the device
    int* d_npattern = nullptr;   // problems reported ST_PATTERN by the verifying kernel
    int* h_rowcol = nullptr;     // pinned, k + 1 ints
    bool sing_known = false;     // d_sing holds the flags (given by the caller, tested on the device, or found by the kernel)
    bool prepared = false;       // tiled-path prerequisites done (sing known, any_sing scanned, A'A when needed)
    // staging of the CSC one-shot solve (socp_b200_solve_host_csc): stored values [batch][nnz] and the linear indices
    double *d_valG = nullptr, *d_valA = nullptr;
    int *d_linG = nullptr, *d_linA = nullptr;
    size_t cap_valG = 0, cap_valA = 0, cap_linG = 0, cap_linA = 0;

    template <class T>
    T* alloc(size_t count, bool zero = true) {
        void* p = nullptr;
        if (count == 0) count = 1;
        CK(cudaMalloc(&p, count * sizeof(T)));
        allocs.push_back(p);
        if (zero) CK(cudaMemsetAsync(p, 0, count * sizeof(T), stream));
        return (T*)p;
    }
    void release() {
        cudaSetDevice(device);
        for (void* p : allocs) cudaFree(p);
        allocs.clear();
        if (h_nactive) cudaFreeHost(h_nactive);
        if (h_rowcol) cudaFreeHost(h_rowcol);
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        for (auto& e : pipe_ev) cudaEventDestroy(e);
        for (auto& e : split_ev)
            if (e) cudaEventDestroy(e);
        if (up_stream) cudaStreamDestroy(up_stream);
        if (alt_stream) cudaStreamDestroy(alt_stream);
        if (down_stream) cudaStreamDestroy(down_stream);
        if (stream) cudaStreamDestroy(stream);
    }
};

}  // namespace

struct socp_handle {
    int n = 0, p = 0, k = 0;
    std::vector<int> kind, offs, dim;          // caller's cones
    std::vector<int> wkind, woffs, wdim;       // work cones
    std::vector<int> first_work;               // caller cone -> first work cone
    int deg = 0;
    int64_t batch = 0;
    std::vector<Shard> shards;
    std::string err;
    socp_timings tim{};
};

namespace {

// ---------------------------------------------------------------- launch helpers
#define LAUNCH(sh, kern, grid, block, smem, ...)                       \
    do {                                                               \
        kern<<<grid, block, smem, (sh).stream>>>(__VA_ARGS__);         \
        CK(cudaPeekAtLastError());   /* configuration errors must not pass silently */ \
        (sh).launches++;                                               \
    } while (0)

// Grid of a batch-wide tiled kernel: `tiles` CTAs per problem on x, the batch folded over y and z (gridDim.y and
// gridDim.z are capped at 65535; the kernels recover b with batch_index() and guard b < nbatch).
dim3 batch_grid(int tiles, int batch) {
    const int gz = (batch + 65534) / 65535;
    const int gy = (batch + gz - 1) / gz;
    return dim3((unsigned)tiles, (unsigned)gy, (unsigned)gz);
}

GemvEpi epi(const double* v1 = nullptr, double c1 = 0, int64_t s1 = 0, const double* v2 = nullptr, double c2 = 0,
            int64_t s2 = 0, int acc = 0) {
    GemvEpi e;
    e.v1 = v1; e.c1 = c1; e.s1 = s1;
    e.v2 = v2; e.c2 = c2; e.s2 = s2;
    e.accumulate = acc;
    return e;
}

// out[cols] = alpha * M' x + epilogue
void gemv_t(Shard& sh, const double* M, int64_t sM, int ld, int rows, int cols, const double* x, int64_t sx,
            double* out, int64_t so, double alpha, GemvEpi e, const int* active, const uint8_t* flag = nullptr) {
    if (cols == 0) return;
    dim3 grid = batch_grid((cols + 7) / 8, sh.batch);
    LAUNCH(sh, k_gemv_t, grid, 256, 0, M, sM, ld, rows, cols, x, sx, out, so, alpha, e, active, flag, sh.batch);
}
void gemv_n(Shard& sh, const double* M, int64_t sM, int ld, int rows, int cols, const double* x, int64_t sx,
            double* out, int64_t so, double alpha, GemvEpi e, const int* active, const uint8_t* flag = nullptr) {
    if (rows == 0) return;
    dim3 grid = batch_grid((rows + 31) / 32, sh.batch);
    LAUNCH(sh, k_gemv_n, grid, dim3(32, 8), 0, M, sM, ld, rows, cols, x, sx, out, so, alpha, e, active, flag, sh.batch);
}

template <int BT, int NW, int KT, bool KM>
void syrk_launch(Shard& sh, const double* A, int64_t sA, int lda, int N, int K, double* C, int64_t sC, int ldc,
                 double alpha, double beta, const double* addC, int64_t sAdd, int ldadd, const uint8_t* addFlag,
                 const int* active, bool first_col_only = false) {
    using Cfg = SyrkCfg<BT, NW, KT, KM>;
    static bool configured[64] = {};
    if (!configured[sh.device]) {
        CK(cudaFuncSetAttribute(k_syrk<BT, NW, KT, KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        configured[sh.device] = true;
    }
    const int nt = (N + BT - 1) / BT;
    dim3 grid = batch_grid(first_col_only ? nt : nt * (nt + 1) / 2, sh.batch);
    LAUNCH(sh, (k_syrk<BT, NW, KT, KM>), grid, Cfg::THREADS, Cfg::SMEM, A, sA, lda, N, K, C, sC, ldc, alpha, beta,
           addC, sAdd, ldadd, addFlag, active, first_col_only ? 1 : 0, sh.batch);
}
void syrk(Shard& sh, bool kmajor, const double* A, int64_t sA, int lda, int N, int K, double* C, int64_t sC, int ldc,
          double alpha, double beta, const double* addC, int64_t sAdd, int ldadd, const uint8_t* addFlag,
          const int* active) {
    if (N <= 0) return;
    // 128 x 128 tiles when they fill the machine at least twice over, 64 x 64 tiles (3 CTAs per SM) otherwise
    const long long t128 = (long long)((N + 127) / 128) * ((N + 127) / 128 + 1) / 2 * sh.batch;
    const bool big = N > 256 && t128 >= 2 * 148;
    if (kmajor) {
        if (big && syrk_tma_supported(A, sA, lda, N, K)) {
            // operands fed by TMA (tensor map over A, mbarrier-signalled 4-stage ring): syrk_tma.cu
            const int nt = (N + 127) / 128;
            CK(syrk_tma_launch(sh.stream, batch_grid(nt * (nt + 1) / 2, sh.batch), A, sA, lda, N, K, C, sC, ldc, alpha, beta,
                               addC, sAdd, ldadd, addFlag, active, sh.batch));
            sh.launches++;
        } else if (big) syrk_launch<128, 4, 16, true>(sh, A, sA, lda, N, K, C, sC, ldc, alpha, beta, addC, sAdd, ldadd, addFlag, active);
        else syrk_launch<64, 2, 16, true>(sh, A, sA, lda, N, K, C, sC, ldc, alpha, beta, addC, sAdd, ldadd, addFlag, active);
    } else {
        if (big) syrk_launch<128, 4, 16, false>(sh, A, sA, lda, N, K, C, sC, ldc, alpha, beta, addC, sAdd, ldadd, addFlag, active);
        else syrk_launch<64, 2, 16, false>(sh, A, sA, lda, N, K, C, sC, ldc, alpha, beta, addC, sAdd, ldadd, addFlag, active);
    }
}

// blocked Cholesky of `nn x nn` matrices (ld, stride), lower, in place, through the DMMA panel kernels
// (panel_mma.cuh).  Xinv ([batch][ceil(nn/64)][64*64]) receives the inverted diagonal blocks for the solves.
void potrf(Shard& sh, double* H, int64_t sH, int ld, int nn, double* Xinv, int* fail, const int* active) {
    static bool configured[64] = {};
    if (!configured[sh.device]) {
        CK(cudaFuncSetAttribute(k_potrf_diag_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_MMA_SMEM));
        CK(cudaFuncSetAttribute(k_trsm_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRSM_MMA_SMEM));
        CK(cudaFuncSetAttribute(k_panel_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_FUSED_SMEM));
        CK(cudaFuncSetAttribute(k_trsv_blk_fwd<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((2048 + 64) * sizeof(double))));
        CK(cudaFuncSetAttribute(k_trsv_blk_bwd<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((2048 + 64) * sizeof(double))));
        CK(cudaFuncSetAttribute(k_trsv_blk_fwd<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((5 * 1024 + 64) * sizeof(double))));
        CK(cudaFuncSetAttribute(k_trsv_blk_bwd<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((2048 + 64) * sizeof(double))));
        configured[sh.device] = true;
    }
    const int nblk = (nn + CHOL_NB - 1) / CHOL_NB;
    auto panel = [&](int j) {          // factor the 64-wide panel at column j: diagonal block, then L21 = A21 L11^-T
        LAUNCH(sh, k_potrf_diag_mma, sh.batch, 256, POTRF_MMA_SMEM, H, sH, ld, nn, j, Xinv, nblk, fail, active);
        const int below = nn - j - CHOL_NB;
        if (below > 0) {
            dim3 grid = batch_grid((below + 127) / 128, sh.batch);
            LAUNCH(sh, k_trsm_mma, grid, 256, TRSM_MMA_SMEM, H, sH, ld, nn, j, (const double*)Xinv, nblk, (const int*)fail, active, sh.batch);
        }
    };
    // Panels are taken in groups of CHOL_GROUP with the trailing update delayed (left-looking inside the group, right-
    // looking between groups): before panel m of a group its 64 columns receive the m earlier panels of the group (a
    // thin product, K = 64 m); after the group the rest of the matrix receives all of them at once (K = 64 CHOL_GROUP) --
    // 1/CHOL_GROUP of the read-modify-write passes over the trailing matrix at CHOL_GROUP times the arithmetic intensity.
    constexpr int CHOL_GROUP = 4;
    // A handful of large problems (C5): the factorisation is a serial chain of panel steps, and each step is ONE launch
    // (k_panel_fused: pending update + diagonal block + panel solve, the row CTAs spinning on a flag until the
    // diagonal CTA is done) -- possible only when every CTA of the launch is resident, one per SM.
    static const bool no_fused_panels = getenv("SOCP_B200_NO_FUSED_PANELS") != nullptr;
    const long long panel_ctas = (long long)(1 + std::max(0, (nn - CHOL_NB + 127) / 128)) * sh.batch;
    // (for a large batch the spinning row CTAs only take SMs away from other problems' diagonal blocks: C4 at batch
    // 1000 factors in 6.7 ms this way against 3.7 ms with the three batch-wide kernels)
    const bool fused_panels = !no_fused_panels && nn >= 4 * CHOL_NB && panel_ctas <= sh.fused2.num_sms;
    if (fused_panels && !sh.d_panel_flags) sh.d_panel_flags = sh.alloc<int>(sh.batch);
    for (int j = 0; j < nn; j += CHOL_GROUP * CHOL_NB) {
        for (int m = 0; m < CHOL_GROUP; ++m) {
            const int jm = j + m * CHOL_NB;
            if (jm >= nn) break;
            if (fused_panels) {
                const int below = nn - jm - CHOL_NB;
                dim3 grid = batch_grid(1 + std::max(0, (below + 127) / 128), sh.batch);
                LAUNCH(sh, k_panel_fused, grid, 256, PANEL_FUSED_SMEM, H, sH, ld, nn, j, jm, Xinv, nblk, fail, active,
                       sh.d_panel_flags, ++sh.panel_gen, sh.batch);
                continue;
            }
            if (m > 0) {   // rows >= jm, columns [jm, jm + 64): first block column of what is left
                const double* P = H + (int64_t)j * ld + jm;
                double* T = H + (int64_t)jm * ld + jm;
                // few tiles (a handful of large problems: C5 has at most 64 of them per update): sixteen warps per tile
                // instead of four -- the update is a serial link of the panel chain and one CTA's DMMA rate sets its length
                const long long tiles = (long long)((nn - jm + 63) / 64) * sh.batch;
                if (tiles <= 148) syrk_launch<64, 4, 16, false>(sh, P, sH, ld, nn - jm, m * CHOL_NB, T, sH, ld, -1.0, 1.0, nullptr, 0, 0, nullptr, active, true);
                else syrk_launch<64, 2, 16, false>(sh, P, sH, ld, nn - jm, m * CHOL_NB, T, sH, ld, -1.0, 1.0, nullptr, 0, 0, nullptr, active, true);
            }
            panel(jm);
        }
        const int jn = j + CHOL_GROUP * CHOL_NB;
        if (jn < nn) {
            const double* P = H + (int64_t)j * ld + jn;
            double* T = H + (int64_t)jn * ld + jn;
            syrk(sh, false, P, sH, ld, nn - jn, CHOL_GROUP * CHOL_NB, T, sH, ld, -1.0, 1.0, nullptr, 0, 0, nullptr, active);
        }
    }
}
// X <- (L L')^-1 X, X is nn x nrhs (ld ldx), with the inverted diagonal blocks of potrf.  Up to 1024 rows one CTA
// per (rhs, problem) walks the whole factor; beyond that the factor is cut into diagonal blocks of 512 solved the
// same way, with the off-diagonal panels applied by the batch-wide gemv kernels (many CTAs) in between -- a single
// CTA streaming a 4096^2 factor would be far slower than the gemvs.
void potrs(Shard& sh, const double* L, int64_t sL, int ld, int nn, const double* Xinv, double* X, int64_t sX, int ldx,
           int nrhs, const int* active) {
    if (nn == 0 || nrhs == 0) return;
    constexpr int BS = 512;
    const int nblk = (nn + CHOL_NB - 1) / CHOL_NB;
    dim3 grid = batch_grid(nrhs, sh.batch);
    // few (rhs, problem) pairs (at most two 1024-thread CTAs per SM, one wave): a CTA of 1024 threads walks its
    // factor four times faster than one of 256
    const bool wide = (long long)nrhs * sh.batch <= 2 * 148;
    auto fwd = [&](size_t smem, const double* Lp, int rows, int blk0, double* Xp) {
        if (wide) LAUNCH(sh, k_trsv_blk_fwd<1024>, grid, 1024, smem + (size_t)4 * rows * sizeof(double), Lp, sL, ld, rows, Xinv, nblk, blk0, Xp, sX, ldx, active, sh.batch);
        else LAUNCH(sh, k_trsv_blk_fwd<256>, grid, 256, smem, Lp, sL, ld, rows, Xinv, nblk, blk0, Xp, sX, ldx, active, sh.batch);
    };
    auto bwd = [&](size_t smem, const double* Lp, int rows, int blk0, double* Xp) {
        if (wide) LAUNCH(sh, k_trsv_blk_bwd<1024>, grid, 1024, smem, Lp, sL, ld, rows, Xinv, nblk, blk0, Xp, sX, ldx, active, sh.batch);
        else LAUNCH(sh, k_trsv_blk_bwd<256>, grid, 256, smem, Lp, sL, ld, rows, Xinv, nblk, blk0, Xp, sX, ldx, active, sh.batch);
    };
    if (nn <= 2 * BS) {
        const size_t smem = (size_t)(nn + 64) * sizeof(double);
        fwd(smem, L, nn, 0, X);
        bwd(smem, L, nn, 0, X);
        return;
    }
    for (int jb = 0; jb < nn; jb += BS) {                       // forward: L y = x
        const int bs = std::min(BS, nn - jb), below = nn - jb - bs;
        fwd((size_t)(bs + 64) * sizeof(double), L + (int64_t)jb * ld + jb, bs, jb / CHOL_NB, X + jb);
        for (int q = 0; q < nrhs && below > 0; ++q)
            gemv_n(sh, L + (int64_t)jb * ld + jb + bs, sL, ld, below, bs, X + (int64_t)q * ldx + jb, sX,
                   X + (int64_t)q * ldx + jb + bs, sX, -1.0, epi(nullptr, 0, 0, nullptr, 0, 0, 1), active);
    }
    for (int jb = (nn - 1) / BS * BS; jb >= 0; jb -= BS) {       // backward: L' x = y
        const int bs = std::min(BS, nn - jb), below = nn - jb - bs;
        for (int q = 0; q < nrhs && below > 0; ++q)
            gemv_t(sh, L + (int64_t)jb * ld + jb + bs, sL, ld, below, bs, X + (int64_t)q * ldx + jb + bs, sX,
                   X + (int64_t)q * ldx + jb, sX, -1.0, epi(nullptr, 0, 0, nullptr, 0, 0, 1), active);
        bwd((size_t)(bs + 64) * sizeof(double), L + (int64_t)jb * ld + jb, bs, jb / CHOL_NB, X + jb);
    }
}

// grid of the batch-wide cone kernels: enough 256-thread CTAs for one lane group per cone / one thread per
// positive-orthant row, capped at 16 CTAs per SM (grid-stride beyond that)
int cone_grid(const Shard& sh) {
    const long long thr = std::max((long long)sh.batch * sh.bl.nsoc * sh.bl.lpc, (long long)sh.batch * sh.bl.kpoc);
    return (int)std::max(1LL, std::min((thr + 255) / 256, 148LL * 16));
}
// Gt = W^-1 G (identity: Gt = G) for the whole shard
void launch_build_gt(Shard& sh, bool identity, const int* active) {
    Ws& w = sh.w;
    const int n = w.L.n;
    const long long pairs = (long long)sh.batch * n * std::max({sh.bl.nsoc, sh.bl.kpoc, 1});
    if (sh.bl_ok && !identity && pairs < (1LL << 31)) {
        // one lane group per (problem, cone, chunk of 8 columns); one thread per positive-orthant entry
        const long long thr = std::max((long long)sh.batch * ((n + 7) / 8) * sh.bl.nsoc * sh.bl.lpc, (long long)sh.batch * n * sh.bl.kpoc);
        const int grid = (int)std::max(1LL, std::min((thr + 255) / 256, 148LL * 16));
        LAUNCH(sh, bk_build_gt, grid, 256, 0, sh.bl, sh.batch, n, w.G, w.sG, w.wb, w.iwb, w.eta, w.Gt, w.ldgt, active);
        return;
    }
    const int cols_per_cta = std::max(1, std::min(n, 2048 / std::max(1, w.L.ncones * 8)));
    dim3 g = batch_grid((n + cols_per_cta - 1) / cols_per_cta, sh.batch);
    LAUNCH(sh, k_build_gt, g, 256, 0, w.L, w.G, w.sG, w.wb, w.iwb, w.eta, w.Gt, w.ldgt, identity ? 1 : 0, cols_per_cta, active, sh.batch);
}
void launch_scaling(Shard& sh, const int* active) {
    Ws& w = sh.w;
    if (sh.bl_ok) LAUNCH(sh, bk_scaling, cone_grid(sh), 256, 0, sh.bl, sh.batch, w.s, w.z, w.lam, w.wb, w.iwb, w.eta, w.fail, active);
    else LAUNCH(sh, k_scaling, sh.batch, sh.threads, 0, w.L, w.s, w.z, w.lam, w.wb, w.iwb, w.eta, w.fail, active);
}
void launch_apply(Shard& sh, int mode, const double* v, double* out) {
    Ws& w = sh.w;
    const int B = sh.batch;
    if (sh.bl_ok) {
        const int g = cone_grid(sh);
        if (mode == 0) LAUNCH(sh, bk_apply<APPLY_W>, g, 256, 0, sh.bl, B, w.wb, w.iwb, w.eta, v, out);
        else if (mode == 1) LAUNCH(sh, bk_apply<APPLY_WINV>, g, 256, 0, sh.bl, B, w.wb, w.iwb, w.eta, v, out);
        else LAUNCH(sh, bk_apply<APPLY_WINV2>, g, 256, 0, sh.bl, B, w.wb, w.iwb, w.eta, v, out);
    } else {
        if (mode == 0) LAUNCH(sh, k_apply<APPLY_W>, B, sh.threads, 0, w.L, w.wb, w.iwb, w.eta, v, out);
        else if (mode == 1) LAUNCH(sh, k_apply<APPLY_WINV>, B, sh.threads, 0, w.L, w.wb, w.iwb, w.eta, v, out);
        else LAUNCH(sh, k_apply<APPLY_WINV2>, B, sh.threads, 0, w.L, w.wb, w.iwb, w.eta, v, out);
    }
}
void launch_vprod(Shard& sh, const double* u, const double* v, double* t) {
    if (sh.bl_ok) LAUNCH(sh, bk_vprod, cone_grid(sh), 256, 0, sh.bl, sh.batch, u, v, t);
    else LAUNCH(sh, k_vprod, sh.batch, sh.threads, 0, sh.w.L, u, v, t);
}
void launch_iprod(Shard& sh, const double* lam, const double* v, double* t) {
    if (sh.bl_ok) LAUNCH(sh, bk_iprod, cone_grid(sh), 256, 0, sh.bl, sh.batch, lam, v, t);
    else LAUNCH(sh, k_iprod, sh.batch, sh.threads, 0, sh.w.L, lam, v, t);
}
void launch_max_step(Shard& sh, const double* x, double* out) {
    if (sh.bl_warp) LAUNCH(sh, bk_max_step, (sh.batch + 8 * (32 / sh.bl_sw) - 1) / (8 * (32 / sh.bl_sw)), 256, 0, sh.bl, sh.batch, sh.bl_sw, x, out);
    else LAUNCH(sh, k_max_step, sh.batch, sh.threads, 0, sh.w.L, x, out);
}
void launch_compute_step(Shard& sh, const double* lam, const double* ds, const double* dz, double* out) {
    if (sh.bl_warp) LAUNCH(sh, bk_compute_step, (sh.batch + 8 * (32 / sh.bl_sw) - 1) / (8 * (32 / sh.bl_sw)), 256, 0, sh.bl, sh.batch, sh.bl_sw, lam, ds, dz, out);
    else LAUNCH(sh, k_compute_step, sh.batch, sh.threads, 0, sh.w.L, lam, ds, dz, out);
}

void factor_one(Shard& sh, bool identity, bool add_aa, const int* active);

// The factor of a large batch without equality rows, as two half batches on two streams: the Gram product of one
// half (DMMA bound) runs under the Cholesky panels of the other (short latency-bound kernels: diagonal blocks,
// triangular panel solves), which on their own leave most of the machine idle.  Same kernels, same per-problem
// arithmetic: results are bit-identical to the single-stream order.
bool factor_split_ok(const Shard& sh) {
    static const bool off = getenv("SOCP_B200_NO_SPLIT_FACTOR") != nullptr;
    return !off && sh.w.L.p == 0 && sh.batch >= 256 && sh.w.L.n >= 128;
}
void factor_split(Shard& sh, bool identity, const int* active) {
    Ws& w = sh.w;
    const int n = w.L.n, k = w.L.k, nc = w.L.ncones;
    if (!sh.alt_stream) CK(cudaStreamCreateWithFlags(&sh.alt_stream, cudaStreamNonBlocking));
    if (!sh.split_ev[0]) {
        CK(cudaEventCreateWithFlags(&sh.split_ev[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&sh.split_ev[1], cudaEventDisableTiming));
    }
    const int B = sh.batch, h0 = B / 2;
    const Ws saved = w;
    cudaStream_t main_stream = sh.stream;
    CK(cudaEventRecord(sh.split_ev[0], main_stream));
    CK(cudaStreamWaitEvent(sh.alt_stream, sh.split_ev[0], 0));
    const int nblk = (n + CHOL_NB - 1) / CHOL_NB;
    for (int part = 0; part < 2; ++part) {
        const int64_t off = part ? h0 : 0;
        sh.batch = part ? B - h0 : h0;
        sh.stream = part ? sh.alt_stream : main_stream;
        w = saved;
        w.G = saved.G + off * saved.sG;
        w.wb = saved.wb + off * k;
        w.iwb = saved.iwb + off * k;
        w.eta = saved.eta + off * 4 * nc;
        w.Gt = saved.Gt + off * (int64_t)saved.ldgt * n;
        w.H = saved.H + off * (int64_t)saved.ldh * n;
        w.XH = saved.XH + off * (int64_t)nblk * 4096;
        w.fail = saved.fail + off;
        w.sing = saved.sing ? saved.sing + off : nullptr;
        factor_one(sh, identity, false, active ? active + off : nullptr);
    }
    w = saved;
    sh.batch = B;
    sh.stream = main_stream;
    CK(cudaEventRecord(sh.split_ev[1], sh.alt_stream));
    CK(cudaStreamWaitEvent(main_stream, sh.split_ev[1], 0));
}

// KKT factor, reference src/densesolver.jl:41-52.  identity: W = I (initial point, sing test).
void factor(Shard& sh, bool identity, bool add_aa, const int* active) {
    if (factor_split_ok(sh)) factor_split(sh, identity, active);
    else factor_one(sh, identity, add_aa, active);
}
void factor_one(Shard& sh, bool identity, bool add_aa, const int* active) {
    Ws& w = sh.w;
    const int n = w.L.n, p = w.L.p;
    launch_build_gt(sh, identity, active);
    const bool aa = add_aa && p > 0 && sh.any_sing;
    syrk(sh, true, w.Gt, (int64_t)w.ldgt * n, w.ldgt, n, w.kpad, w.H, (int64_t)w.ldh * n, w.ldh, 1.0, 0.0,
         aa ? w.AA : nullptr, (int64_t)w.ldh * n, w.ldh, w.sing, active);
    potrf(sh, w.H, (int64_t)w.ldh * n, w.ldh, n, w.XH, w.fail, active);
    if (p > 0) {
        dim3 gt = batch_grid(std::max(1, std::min(64, (p * n + 255) / 256)), sh.batch);
        LAUNCH(sh, k_transpose_A, gt, 256, 0, w.A, w.sA, p, n, w.HiAt, active, sh.batch);
        potrs(sh, w.H, (int64_t)w.ldh * n, w.ldh, n, w.XH, w.HiAt, (int64_t)n * p, n, p, active);
        dim3 gm = batch_grid(std::max(1, std::min(64, (p * p + 255) / 256)), sh.batch);
        LAUNCH(sh, k_small_gemm, gm, 256, 0, w.A, w.sA, p, n, w.HiAt, w.M, w.ldm, active, sh.batch);
        potrf(sh, w.M, (int64_t)w.ldm * p, w.ldm, p, w.XM, w.fail, active);
    }
}

// Middle of solve_kkt, reference src/densesolver.jl:66-85, between kkt_head and
// kkt_tail.  On entry u = W^-2 k2; on exit rx = cx, ry = cy, u = G cx - k2.
void kkt_middle(Shard& sh, const int* active) {
    Ws& w = sh.w;
    const int n = w.L.n, p = w.L.p, k = w.L.k;
    // n0 = G'u + dx (+A'dy if sing)                                      :66-71
    gemv_t(sh, w.G, w.sG, k, k, n, w.u, k, w.rx, n, 1.0, epi(w.dx, 1.0, n), active);
    if (p > 0 && sh.any_sing)
        gemv_t(sh, w.A, w.sA, p, p, n, w.dy, p, w.rx, n, 1.0, epi(nullptr, 0, 0, nullptr, 0, 0, 1), active, w.sing);
    // t = H^-1 n0
    potrs(sh, w.H, (int64_t)w.ldh * n, w.ldh, n, w.XH, w.rx, n, n, 1, active);
    if (p > 0) {
        // m0 = A t - dy; cy = M^-1 m0                                     :73-75
        gemv_n(sh, w.A, w.sA, p, p, n, w.rx, n, w.ry, p, 1.0, epi(w.dy, -1.0, p), active);
        potrs(sh, w.M, (int64_t)w.ldm * p, w.ldm, p, w.XM, w.ry, p, p, 1, active);
        // cx = H^-1 (n0 + A'(sing ? dy - cy : -cy)) = t - HiAt cy (+ HiAt dy if sing)   :76-83
        gemv_n(sh, w.HiAt, (int64_t)n * p, n, n, p, w.ry, p, w.rx, n, -1.0, epi(nullptr, 0, 0, nullptr, 0, 0, 1), active);
        if (sh.any_sing)
            gemv_n(sh, w.HiAt, (int64_t)n * p, n, n, p, w.dy, p, w.rx, n, 1.0, epi(nullptr, 0, 0, nullptr, 0, 0, 1), active, w.sing);
    }
    // k1 = G cx - k2                                                      :84-85
    gemv_n(sh, w.G, w.sG, k, k, n, w.rx, n, w.u, k, 1.0, epi(w.k2, -1.0, k), active);
}

__global__ void k_kkt_head(Ws w) { kkt_head(w, blockIdx.x); }
__global__ void k_kkt_tail(Ws w) { kkt_tail(w, blockIdx.x); }

// Initial point, reference src/solver.jl:68-104, by block elimination with W = I
// (the oracle's initial_point_reduced).
void initial_point(Shard& sh, const LoopParams& P) {
    Ws& w = sh.w;
    const int n = w.L.n, p = w.L.p, k = w.L.k, B = sh.batch;
    factor(sh, true, true, w.active);
    LAUNCH(sh, k_finalize, B, 32, 0, w, 0);
    // n0 = -c + G'h (+A'b if sing)
    gemv_t(sh, w.G, w.sG, k, k, n, w.h, k, w.x, n, 1.0, epi(w.c, -1.0, n), w.active);
    if (p > 0 && sh.any_sing)
        gemv_t(sh, w.A, w.sA, p, p, n, w.b, p, w.x, n, 1.0, epi(nullptr, 0, 0, nullptr, 0, 0, 1), w.active, w.sing);
    potrs(sh, w.H, (int64_t)w.ldh * n, w.ldh, n, w.XH, w.x, n, n, 1, w.active);          // t = H^-1 n0
    if (p > 0) {
        gemv_n(sh, w.A, w.sA, p, p, n, w.x, n, w.y, p, 1.0, epi(w.b, -1.0, p), w.active);   // A t - b
        potrs(sh, w.M, (int64_t)w.ldm * p, w.ldm, p, w.XM, w.y, p, p, 1, w.active);       // y
        gemv_n(sh, w.HiAt, (int64_t)n * p, n, n, p, w.y, p, w.x, n, -1.0, epi(nullptr, 0, 0, nullptr, 0, 0, 1), w.active);
    }
    gemv_n(sh, w.G, w.sG, k, k, n, w.x, n, w.z, k, 1.0, epi(w.h, -1.0, k), w.active);       // z0 = G x - h
    LAUNCH(sh, k_init_shift, B, sh.threads, 0, w, P);
}

// The Mehrotra loop, reference src/solver.jl:105-151, over the shard.
void solve_tiled(Shard& sh, const socp_params& prm) {
    Ws& w = sh.w;
    const int n = w.L.n, p = w.L.p, k = w.L.k, B = sh.batch;
    LoopParams P{prm.max_iter, prm.tol, prm.step_damp, prm.init_eps};
    LAUNCH(sh, k_reset, (B + 255) / 256, 256, 0, w, B);
    CK(cudaMemsetAsync(w.nactive, 0, sizeof(int) * (prm.max_iter + 2), sh.stream));
    initial_point(sh, P);
    int itmax = 0;
    for (int it = 0; it < prm.max_iter; ++it) {
        launch_scaling(sh, w.active);                                                   // :106
        // negated residuals                                                           :110-118,:125
        gemv_t(sh, w.G, w.sG, k, k, n, w.z, k, w.dx, n, -1.0, epi(w.c, -1.0, n), w.active);
        if (p > 0) {
            gemv_t(sh, w.A, w.sA, p, p, n, w.y, p, w.dx, n, -1.0, epi(nullptr, 0, 0, nullptr, 0, 0, 1), w.active);
            gemv_n(sh, w.A, w.sA, p, p, n, w.x, n, w.dy, p, -1.0, epi(w.b, 1.0, p), w.active);
        }
        gemv_n(sh, w.G, w.sG, k, k, n, w.x, n, w.dz, k, -1.0, epi(w.s, -1.0, k, w.h, 1.0, k), w.active);
        LAUNCH(sh, k_pre, B, sh.threads, 0, w, P, it);
        CK(cudaMemcpyAsync(sh.h_nactive, w.nactive + it, sizeof(int), cudaMemcpyDeviceToHost, sh.stream));
        CK(cudaStreamSynchronize(sh.stream));
        if (*sh.h_nactive == 0) break;
        itmax = it + 1;
        factor(sh, false, true, w.active);                                              // :126
        kkt_middle(sh, w.active);                                                       // :127
        LAUNCH(sh, k_mid, B, sh.threads, 0, w, P);                                     // :128-140
        kkt_middle(sh, w.active);                                                       // :141
        LAUNCH(sh, k_post, B, sh.threads, 0, w, P);                                    // :143-150
    }
    LAUNCH(sh, k_finalize, B, sh.threads, 0, w, 1);
    sh.tim.iterations_max = itmax;
    sh.tim.path_used = SOCP_PATH_TILED;
}

void build_shard(socp_handle* h, Shard& sh) {
    CK(cudaSetDevice(sh.device));
    CK(cudaStreamCreateWithFlags(&sh.stream, cudaStreamNonBlocking));
    for (auto& e : sh.ev) CK(cudaEventCreate(&e));
    const int n = h->n, p = h->p, k = h->k, B = sh.batch;
    const int nc = (int)h->wkind.size();
    Ws& w = sh.w;
    sh.d_kind = sh.alloc<int>(nc);
    sh.d_offs = sh.alloc<int>(nc);
    sh.d_dim = sh.alloc<int>(nc);
    CK(cudaMemcpyAsync(sh.d_kind, h->wkind.data(), nc * sizeof(int), cudaMemcpyHostToDevice, sh.stream));
    CK(cudaMemcpyAsync(sh.d_offs, h->woffs.data(), nc * sizeof(int), cudaMemcpyHostToDevice, sh.stream));
    CK(cudaMemcpyAsync(sh.d_dim, h->wdim.data(), nc * sizeof(int), cudaMemcpyHostToDevice, sh.stream));
    w.L = ConeLayout{n, p, k, nc, h->deg, sh.d_kind, sh.d_offs, sh.d_dim};
    w.kpad = round_up(std::max(k, 1), 16);
    w.ldgt = w.kpad;
    w.ldh = round_up(std::max(n, 1), 2);
    w.ppad = round_up(std::max(p, 1), 16);
    w.ldap = w.ppad;
    w.ldm = round_up(std::max(p, 1), 2);
    sh.d_c = sh.alloc<double>((size_t)B * n);
    sh.d_b = sh.alloc<double>((size_t)B * p);
    sh.d_h = sh.alloc<double>((size_t)B * k);
    sh.d_A = sh.alloc<double>((size_t)B * p * n);
    sh.d_G = sh.alloc<double>((size_t)B * k * n);
    sh.d_sing = sh.alloc<uint8_t>(B);
    w.c = sh.d_c; w.A = sh.d_A; w.b = sh.d_b; w.G = sh.d_G; w.h = sh.d_h; w.sing = sh.d_sing;
    w.sA = (int64_t)p * n; w.sG = (int64_t)k * n;
    w.x = sh.alloc<double>((size_t)B * n);  w.y = sh.alloc<double>((size_t)B * p);
    w.z = sh.alloc<double>((size_t)B * k);  w.s = sh.alloc<double>((size_t)B * k);
    w.lam = sh.alloc<double>((size_t)B * k); w.wb = sh.alloc<double>((size_t)B * k);
    w.iwb = sh.alloc<double>((size_t)B * k);
    w.eta = sh.alloc<double>((size_t)B * 4 * nc);
    w.dx = sh.alloc<double>((size_t)B * n);  w.dy = sh.alloc<double>((size_t)B * p);
    w.dz = sh.alloc<double>((size_t)B * k);  w.ds = sh.alloc<double>((size_t)B * k);
    w.rx = sh.alloc<double>((size_t)B * n);  w.ry = sh.alloc<double>((size_t)B * p);
    w.rz = sh.alloc<double>((size_t)B * k);  w.rs = sh.alloc<double>((size_t)B * k);
    w.k0 = sh.alloc<double>((size_t)B * k);  w.k2 = sh.alloc<double>((size_t)B * k);
    w.u = sh.alloc<double>((size_t)B * k);
    w.kt2 = sh.alloc<double>((size_t)B * k); w.kt3 = sh.alloc<double>((size_t)B * k);
    w.sc = sh.alloc<ProbScalars>(B);
    w.pobj = sh.alloc<double>(B); w.dobj = sh.alloc<double>(B);
    w.status = sh.alloc<int>(B); w.iters = sh.alloc<int>(B);
    w.active = sh.alloc<int>(B); w.fail = sh.alloc<int>(B);
    w.nactive = sh.alloc<int>(4096);
    CK(cudaMallocHost((void**)&sh.h_nactive, 64));
    // tiled-path factor workspaces are allocated lazily (ensure_tiled): the fused
    // path does not need them and they dominate the footprint
    sh.threads = std::min(256, std::max(32, 32 * nc));
    {
        std::vector<int> so, sd, sw;
        int kpoc = 0, maxd = 1;
        for (size_t i = 0; i < h->kind.size(); ++i) {
            if (h->kind[i] == SOCP_CONE_POC) kpoc += h->dim[i];
            else { so.push_back(h->offs[i]); sd.push_back(h->dim[i]); sw.push_back(h->first_work[i]); maxd = std::max(maxd, h->dim[i]); }
        }
        const int ns = (int)so.size();
        int* d = sh.alloc<int>(3 * std::max(ns, 1));
        if (ns) {
            CK(cudaMemcpyAsync(d, so.data(), ns * sizeof(int), cudaMemcpyHostToDevice, sh.stream));
            CK(cudaMemcpyAsync(d + ns, sd.data(), ns * sizeof(int), cudaMemcpyHostToDevice, sh.stream));
            CK(cudaMemcpyAsync(d + 2 * ns, sw.data(), ns * sizeof(int), cudaMemcpyHostToDevice, sh.stream));
        }
        sh.bl = BLayout{k, kpoc, ns, f2_lpc(maxd), nc, d, d + ns, d + 2 * ns};
        // the batch-wide cone kernels index (problem, cone) pairs and orthant rows with 32-bit arithmetic
        sh.bl_ok = maxd <= 128 && (long long)B * std::max(ns, 1) < (1LL << 31) && (long long)B * std::max(kpoc, 1) < (1LL << 31);
        sh.bl_warp = sh.bl_ok && ns * sh.bl.lpc <= 32;
        sh.bl_sw = 1;
        while (sh.bl_sw < std::max(1, ns * sh.bl.lpc)) sh.bl_sw <<= 1;
        sh.bl_sw = std::min(sh.bl_sw, 32);
    }
    f2_plan(sh.fused2, n, p, k, h->kind, h->offs, h->dim, sh.device);
    sh.fused2.d_counter = sh.alloc<int>(16);
    sh.fused2.d_clk = sh.alloc<unsigned long long>(32);
    {
        int dev_smem = 0, sms = 148;
        CK(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, sh.device));
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, sh.device));
        fl_plan(sh.lane, n, p, k, h->kind, h->offs, h->dim, dev_smem, sms);
        sh.lane.d_counter = sh.alloc<int>(FL_WS_SETS);
    }
    sh.d_f3_tables = sh.alloc<int>((size_t)2 * k + n + 2);
    sh.d_rowcol = sh.alloc<int>(k);
    sh.d_npattern = sh.alloc<int>(1);
    CK(cudaMallocHost((void**)&sh.h_rowcol, sizeof(int) * (k + 1)));
    CK(cudaStreamSynchronize(sh.stream));
}

void ensure_tiled(Shard& sh) {
    Ws& w = sh.w;
    if (w.Gt) return;
    const int n = w.L.n, p = w.L.p, B = sh.batch;
    w.Gt = sh.alloc<double>((size_t)B * w.ldgt * n);
    w.H = sh.alloc<double>((size_t)B * w.ldh * n);
    w.HiAt = sh.alloc<double>((size_t)B * n * p);
    w.M = sh.alloc<double>((size_t)B * w.ldm * p);
    w.AA = sh.alloc<double>((size_t)B * w.ldh * n * (p > 0 ? 1 : 0));
    w.Ap = sh.alloc<double>((size_t)B * w.ldap * n * (p > 0 ? 1 : 0));
    w.XH = sh.alloc<double>((size_t)B * ((n + 63) / 64) * 4096, false);
    w.XM = sh.alloc<double>((size_t)B * ((p + 63) / 64) * 4096 * (p > 0 ? 1 : 0), false);
}

// Prerequisites of the tiled path (and of get_sing): `sing` known -- when the caller gave none, the test of
// src/Socp.jl:49-56, cholesky(G'G) fails -- and AA = A'A (reference src/densesolver.jl:32) when some problem is sing.
// Lazy: the whole-solve kernel of fused_v3.cuh needs neither (it tests and handles sing problems itself).
void ensure_prepared(Shard& sh) {
    if (sh.prepared) return;
    Ws& w = sh.w;
    const int n = w.L.n, p = w.L.p, B = sh.batch;
    if (!sh.sing_known) {
        ensure_tiled(sh);
        LAUNCH(sh, k_reset, (B + 255) / 256, 256, 0, w, B);
        const int cols_per_cta = std::max(1, std::min(n, 2048 / std::max(1, w.L.ncones * 8)));
        dim3 g = batch_grid((n + cols_per_cta - 1) / cols_per_cta, B);
        LAUNCH(sh, k_build_gt, g, 256, 0, w.L, w.G, w.sG, w.wb, w.iwb, w.eta, w.Gt, w.ldgt, 1, cols_per_cta, (const int*)nullptr, B);
        syrk(sh, true, w.Gt, (int64_t)w.ldgt * n, w.ldgt, n, w.kpad, w.H, (int64_t)w.ldh * n, w.ldh, 1.0, 0.0, nullptr, 0,
             0, nullptr, nullptr);
        potrf(sh, w.H, (int64_t)w.ldh * n, w.ldh, n, w.XH, w.fail, nullptr);
        LAUNCH(sh, k_fail_to_sing, (B + 255) / 256, 256, 0, w.fail, sh.d_sing, B);
        sh.sing_known = true;
    }
    std::vector<uint8_t> hs(B);
    CK(cudaMemcpyAsync(hs.data(), sh.d_sing, B, cudaMemcpyDeviceToHost, sh.stream));
    CK(cudaStreamSynchronize(sh.stream));
    sh.any_sing = false;
    for (uint8_t v : hs) sh.any_sing |= (v != 0);
    if (p > 0 && sh.any_sing) {
        ensure_tiled(sh);
        const int64_t sA = w.sA;
        dim3 g = batch_grid(std::max(1, std::min(64, (p * n + 255) / 256)), sh.sharedA ? 1 : B);
        LAUNCH(sh, k_pad_copy, g, 256, 0, w.A, sA, p, n, w.Ap, w.ldap, sh.sharedA ? 1 : B);
        // shared A: every problem reads slice 0 of Ap
        syrk(sh, true, w.Ap, sh.sharedA ? 0 : (int64_t)w.ldap * n, w.ldap, n, w.ppad, w.AA, (int64_t)w.ldh * n, w.ldh,
             1.0, 0.0, nullptr, 0, 0, nullptr, nullptr);
        CK(cudaStreamSynchronize(sh.stream));
    }
    sh.prepared = true;
}

template <class F>
int guarded(socp_handle* h, F f) {
    try {
        f();
        return 0;
    } catch (const CudaErr& e) {
        if (h) h->err = e.msg;
        cudaGetLastError();
        return e.code > 0 ? e.code : 999;
    } catch (const UsageErr& e) {
        if (h) h->err = e.msg;
        return e.code;
    } catch (const std::exception& e) {
        if (h) h->err = e.what();
        return SOCP_ERR_NOMEM;
    }
}

template <class F>
void for_each_shard(socp_handle* h, F f) {
    if (h->shards.size() == 1) {
        CK(cudaSetDevice(h->shards[0].device));
        f(h->shards[0]);
        return;
    }
    std::vector<std::thread> th;
    std::vector<CudaErr> errs(h->shards.size(), CudaErr{0, ""});
    for (size_t i = 0; i < h->shards.size(); ++i) {
        th.emplace_back([&, i]() {
            try {
                CK(cudaSetDevice(h->shards[i].device));
                f(h->shards[i]);
            } catch (const CudaErr& e) {
                errs[i] = e;
            } catch (const UsageErr& e) {
                errs[i] = CudaErr{e.code, e.msg};
            } catch (const std::exception& e) {
                errs[i] = CudaErr{SOCP_ERR_NOMEM, e.what()};
            } catch (...) {
                errs[i] = CudaErr{SOCP_ERR_NOMEM, "unknown exception in a shard worker"};
            }
        });
    }
    for (auto& t : th) t.join();
    for (auto& e : errs) {
        if (e.code > 0) throw e;
        if (e.code < 0) throw UsageErr{e.code, e.msg};
    }
}

void h2d(Shard& sh, void* dst, const void* src, size_t bytes) {
    if (bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, sh.stream));
}
void d2h(Shard& sh, void* dst, const void* src, size_t bytes) {
    if (bytes && dst) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, sh.stream));
}

void need(bool cond, int code, const char* msg) {
    if (!cond) throw UsageErr{code, msg};
}

// Row pattern of G over problems [first, first + count) of the shard (-1 empty, j single column, -2 dense) ->
// sh.h_rowcol (pinned).  One pass over G; per-CTA state in shared memory, merged into global memory at the end.
__global__ void __launch_bounds__(256)
k_row_pattern(const double* __restrict__ G, int64_t sG, int nbatch, int n, int k, int* __restrict__ rowcol) {
    extern __shared__ int s_col[];
    for (int i = threadIdx.x; i < k; i += blockDim.x) s_col[i] = -1;
    __syncthreads();
    const int64_t per = (int64_t)n * k;
    for (int b = blockIdx.x; b < nbatch; b += gridDim.x) {
        const double* Gb = G + (int64_t)b * sG;
        for (int64_t e = threadIdx.x; e < per; e += blockDim.x) {
            if (Gb[e] != 0.0) {
                const int j = (int)(e / k), i = (int)(e - (int64_t)j * k);
                const int old = atomicCAS(&s_col[i], -1, j);
                if (old != -1 && old != j) s_col[i] = -2;
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const int c = s_col[i];
        if (c == -1) continue;
        if (c == -2) { rowcol[i] = -2; continue; }
        const int old = atomicCAS(&rowcol[i], -1, c);
        if (old != -1 && old != c) rowcol[i] = -2;
    }
}

bool f3_candidate(const socp_handle* h) {
    return h->n > 16 && h->n <= 64 && h->p <= 32 && getenv("SOCP_B200_NO_V3") == nullptr;
}

// plans the whole-solve kernel of fused_v3.cuh for the row pattern `rowcol` (k entries) and uploads its tables
void plan_fused3(const socp_handle* h, Shard& sh, const std::vector<int>& rowcol, cudaStream_t stream) {
    int dev_smem = 0, sms = 148;
    CK(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, sh.device));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, sh.device));
    std::vector<int> tables;
    int* counter = sh.fused2.d_counter;
    unsigned long long* clk = sh.fused2.d_clk + 16;
    f3_plan(sh.fused3, h->n, h->p, h->k, h->kind, h->offs, h->dim, rowcol, dev_smem, sms, tables);
    sh.fused3.d_counter = counter;
    sh.fused3.d_clk = clk;
    sh.fused3.d_tables = sh.d_f3_tables;
    sh.f3_planned = true;
    if (!sh.fused3.fits) return;
    // a layout other than BASELINE.json's C2: the kernel with the layout as compile-time constants, built at run time
    // (once per process, device and layout; lane_jit.cu) -- the runtime-dimension instantiation otherwise
    sh.fused3.jit_fn = nullptr;
    if (!Dims3C2::matches(sh.fused3) && !getenv("SOCP_B200_NO_F3_JIT") && !getenv("SOCP_B200_GENERIC_ONLY")) {
        const char* e4 = getenv("SOCP_B200_F3_TEAMS");
        const int teams = (sh.fused3.teams4 && !(e4 && atoi(e4) == 1)) ? 4 : 1;
        sh.fused3.jit_fn = fused3_jit_get(sh.fused3, teams, sh.device);
        sh.fused3.jit_teams = teams;
    }
    // the tables are read by the launch that follows on the same stream; the host vector dies with this call, so
    // the copy must have left pageable memory before returning (cudaMemcpyAsync from pageable memory stages it)
    CK(cudaMemcpyAsync(sh.d_f3_tables, tables.data(), sizeof(int) * tables.size(), cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
}

// pattern of problems [first, first + count) (or of the one shared G) found on the device, then the plan
void detect_and_plan_fused3(const socp_handle* h, Shard& sh, int first, int count, cudaStream_t stream) {
    const int k = h->k, n = h->n;
    std::vector<int> rowcol(k, -2);
    if (!sh.f3_dense) {
        CK(cudaMemsetAsync(sh.d_rowcol, 0xFF, sizeof(int) * k, stream));
        const int nb = sh.sharedG ? 1 : count;
        const double* G = sh.w.G + (sh.sharedG ? 0 : (int64_t)first * sh.w.sG);
        const int grid = std::max(1, std::min(nb, 148 * 4));
        k_row_pattern<<<grid, 256, sizeof(int) * k, stream>>>(G, sh.w.sG, nb, n, k, sh.d_rowcol);
        CK(cudaPeekAtLastError());
        sh.launches++;
        CK(cudaMemcpyAsync(sh.h_rowcol, sh.d_rowcol, sizeof(int) * k, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        rowcol.assign(sh.h_rowcol, sh.h_rowcol + k);
    }
    plan_fused3(h, sh, rowcol, stream);
}

F3Glob f3_glob(const Shard& sh) {
    const Ws& w = sh.w;
    F3Glob g{};
    g.c = w.c; g.A = w.A; g.b = w.b; g.G = w.G; g.h = w.h;
    g.sA = w.sA; g.sG = w.sG;
    g.sing = sh.sing_known ? sh.d_sing : nullptr;
    g.sing_out = sh.d_sing;
    g.x = w.x; g.y = w.y; g.z = w.z; g.s = w.s; g.pobj = w.pobj; g.dobj = w.dobj;
    g.status = w.status; g.iters = w.iters; g.active = w.active; g.fail = w.fail;
    g.deg = w.L.deg;
    g.npattern = sh.d_npattern;
    g.dbg = nullptr; g.dbg_prob = -1; g.dbg_iter = -1; g.dbg_phase = 1;
    return g;
}

enum { FUSED_NONE = 0, FUSED_V2 = 2, FUSED_V3 = 3, FUSED_LANE = 4 };

// The lane-per-problem kernel keeps 96 problems per SM in flight, but a problem occupies its lane for about 0.1 ms per
// iteration: a launch lasts at least as long as its slowest problem (~1 ms at 19 iterations) however small the batch,
// while fused_v2's one-warp teams (12 per SM, ~0.02 ms per iteration) retire 8.6M problems/s from the first wave on.
// Measured crossover on C3 (profiles/r02_lane_c3_batch_sweep.txt): between 6k and 9k problems, i.e. about half of the
// lane slots.  SOCP_B200_LANE=0 / 1 switches the kernel off / on regardless of the batch size (tests, experiments).
bool lane_wanted(Shard& sh, int batch) {
    if (!sh.lane.fits) return false;
    bool want = 2LL * batch >= (long long)sh.lane.num_sms * sh.lane.pps;
    if (const char* e = getenv("SOCP_B200_LANE")) want = atoi(e) != 0;
    if (want && sh.lane.shape == 100 && !sh.lane.jit_fn) {
        // a layout without a compile-time instantiation: specialise the kernel now (once per process and device; the
        // callers are outside their timed regions); one-warp teams when that is not possible
        if (!sh.lane.jit_tried) {
            sh.lane.jit_tried = true;
            sh.lane.jit_fn = lane_jit_get(sh.lane, sh.device);
        }
        if (!sh.lane.jit_fn) { sh.lane.fits = false; return false; }
    }
    return want;
}
void ensure_lane_ws(Shard& sh) {
    if (!sh.lane.d_ws) sh.lane.d_ws = sh.alloc<double>(sh.lane.ws_doubles * FL_WS_SETS, true);
}

// which whole-solve kernel can take the data now resident (plans fused_v3 on first use)
int fused_kind(const socp_handle* h, Shard& sh) {
    if (f3_candidate(h)) {
        if (!sh.f3_planned) detect_and_plan_fused3(h, sh, 0, sh.batch, sh.stream);
        if (sh.fused3.fits) return FUSED_V3;
    }
    if (sh.fused2.fits) {
        if (!sh.sing_known) ensure_prepared(sh);       // fused_v2 cannot take sing problems: the flags must be known
        if (!sh.prepared) ensure_prepared(sh);
        if (!sh.any_sing) {
            if (lane_wanted(sh, sh.batch)) return FUSED_LANE;
            fused2_prepare(sh.fused2);          // the kernel specialised for this layout, once per plan (NVRTC)
            return FUSED_V2;
        }
    }
    return FUSED_NONE;
}

int choose_path(const socp_handle* h, Shard& sh, const socp_params& prm, int& kind) {
    kind = FUSED_NONE;
    if (prm.path == SOCP_PATH_TILED) return SOCP_PATH_TILED;
    kind = fused_kind(h, sh);
    if (prm.path == SOCP_PATH_FUSED) {
        need(kind != FUSED_NONE, SOCP_ERR_SIZE, "the fused shared-memory kernel needs a layout that fits (and, for n <= 16, no sing problems)");
        return SOCP_PATH_FUSED;
    }
    return kind != FUSED_NONE ? SOCP_PATH_FUSED : SOCP_PATH_TILED;
}

void run_solve(const socp_handle* h, Shard& sh, const socp_params& prm) {
    need(sh.have_data, SOCP_ERR_STATE, "solve called before set_data");
    sh.launches = 0;
    int kind = FUSED_NONE;
    const int path = choose_path(h, sh, prm, kind);      // may run the pattern / sing tests: outside the timed region
    CK(cudaEventRecord(sh.ev[0], sh.stream));
    const bool allow_static = getenv("SOCP_B200_GENERIC_ONLY") == nullptr;
    if (path == SOCP_PATH_FUSED && kind == FUSED_V3) {
        const LoopParams lp{prm.max_iter, prm.tol, prm.step_damp, prm.init_eps};
        const bool detect = !sh.sing_known;
        solve_fused3_ext(sh.fused3, f3_glob(sh), 0, sh.batch, lp, detect ? 1 : 0, 0, sh.stream, allow_static, 0);
        CK(cudaGetLastError());
        if (detect) { sh.sing_known = true; sh.prepared = false; }      // the kernel left the flags in d_sing
        sh.launches += 1;
        sh.tim.path_used = SOCP_PATH_FUSED;
        sh.tim.iterations_max = -1;
    } else if (path == SOCP_PATH_FUSED && kind == FUSED_LANE) {
        ensure_lane_ws(sh);
        CK(cudaEventRecord(sh.ev[0], sh.stream));
        solve_fused_lane_ext(sh.lane, sh.w, 0, sh.batch, LoopParams{prm.max_iter, prm.tol, prm.step_damp, prm.init_eps}, sh.stream, 0);
        CK(cudaGetLastError());
        sh.launches += 1;
        sh.tim.path_used = SOCP_PATH_FUSED;
        sh.tim.iterations_max = -1;
    } else if (path == SOCP_PATH_FUSED) {
        solve_fused2_ext(sh.fused2, sh.w, 0, sh.batch, prm.max_iter, prm.tol, prm.step_damp, prm.init_eps, sh.stream, allow_static, 0);
        CK(cudaGetLastError());
        sh.launches += 1;
        sh.tim.path_used = SOCP_PATH_FUSED;
        sh.tim.iterations_max = -1;
    } else {
        ensure_prepared(sh);
        ensure_tiled(sh);
        CK(cudaEventRecord(sh.ev[0], sh.stream));
        solve_tiled(sh, prm);
    }
    CK(cudaEventRecord(sh.ev[1], sh.stream));
    CK(cudaStreamSynchronize(sh.stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, sh.ev[0], sh.ev[1]));
    sh.tim.solve_ms = ms;
    sh.tim.kernel_launches = sh.launches;
    sh.have_scaling = sh.have_factor = false;
}

void fetch_results(socp_handle* h, Shard& sh, double* x, double* y, double* z, double* s, int32_t* status,
                   int32_t* iters, double* pobj, double* dobj) {
    const int n = h->n, p = h->p, k = h->k, B = sh.batch;
    const int64_t f = sh.first;
    CK(cudaEventRecord(sh.ev[2], sh.stream));
    if (x) d2h(sh, x + f * n, sh.w.x, sizeof(double) * B * n);
    if (y) d2h(sh, y + f * p, sh.w.y, sizeof(double) * B * p);
    if (z) d2h(sh, z + f * k, sh.w.z, sizeof(double) * B * k);
    if (s) d2h(sh, s + f * k, sh.w.s, sizeof(double) * B * k);
    if (status) d2h(sh, status + f, sh.w.status, sizeof(int) * B);
    if (iters) d2h(sh, iters + f, sh.w.iters, sizeof(int) * B);
    if (pobj) d2h(sh, pobj + f, sh.w.pobj, sizeof(double) * B);
    if (dobj) d2h(sh, dobj + f, sh.w.dobj, sizeof(double) * B);
    CK(cudaEventRecord(sh.ev[3], sh.stream));
    CK(cudaStreamSynchronize(sh.stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, sh.ev[2], sh.ev[3]));
    sh.tim.d2h_ms = ms;
}


// Problem(c, A, b, G, h, cones) + solve_socp(prob, ss) in one pass over host data (fused path only): the batch is
// cut into chunks; chunk i+1 uploads (copy engine) while chunk i solves and chunk i-1 downloads.
bool can_pipeline(const socp_handle* h, const Shard& sh, const socp_params& prm, const uint8_t* sing, int64_t first) {
    if (prm.path == SOCP_PATH_TILED) return false;
    if (f3_candidate(h)) return true;          // fused_v3 takes sing problems and finds them itself when sing == NULL
    if (!sh.fused2.fits || !sing) return false;
    for (int64_t q = 0; q < sh.batch; ++q)
        if (sing[first + q]) return false;
    return true;
}

// CSC -> dense column-major on the device: dense[b][lin[j]] = val[b][j], lin = col * rows + row (host-validated, no
// duplicates, so the scatter has no write conflicts).  One thread per (problem, stored entry).
__global__ void __launch_bounds__(256)
k_csc_scatter(const int* __restrict__ lin, const double* __restrict__ val, int64_t stride_val, int nnz,
              double* __restrict__ dense, int64_t stride_dense, int64_t total) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / nnz;
        const int j = (int)(i - b * nnz);
        dense[b * stride_dense + lin[j]] = val[b * stride_val + j];
    }
}

// A and G of the pipelined one-shot solve in the reference's own storage (SparseMatrixCSC): validated patterns as
// linear indices into the dense column-major operands, and the row pattern of G for the fused_v3 plan
struct CscSrc {
    const socp_csc* A;
    const socp_csc* G;
    const std::vector<int>* linA;
    const std::vector<int>* linG;
    const std::vector<int>* rowcolG;
};

// returns false when the pipeline could not be used after all (the first chunk's pattern does not fit the fused
// kernel): the caller falls back to set_data + solve
bool run_pipelined(socp_handle* h, Shard& sh, const socp_params& prm, const double* c, const double* A, const double* b,
                   const double* G, const double* hvec, const uint8_t* sing, int flags, double* x, double* y, double* z,
                   double* s, int32_t* status, int32_t* iters, double* pobj, double* dobj, const CscSrc* csc = nullptr) {
    const int n = h->n, p = h->p, k = h->k, B = sh.batch;
    const int64_t f = sh.first;
    const int64_t nnzG = csc ? csc->G->nnz : 0, nnzA = (csc && p > 0) ? csc->A->nnz : 0;
    if (csc) {
        auto grow_d = [&](double*& ptr, size_t& cap, size_t want) { if (want > cap) { ptr = sh.alloc<double>(want, false); cap = want; } };
        auto grow_i = [&](int*& ptr, size_t& cap, size_t want) { if (want > cap) { ptr = sh.alloc<int>(want, false); cap = want; } };
        grow_d(sh.d_valG, sh.cap_valG, (size_t)std::max<int64_t>(1, nnzG) * B);
        grow_i(sh.d_linG, sh.cap_linG, (size_t)std::max<int64_t>(1, nnzG));
        if (p > 0) {
            grow_d(sh.d_valA, sh.cap_valA, (size_t)std::max<int64_t>(1, nnzA) * B);
            grow_i(sh.d_linA, sh.cap_linA, (size_t)std::max<int64_t>(1, nnzA));
        }
    }
    if (!sh.up_stream) CK(cudaStreamCreateWithFlags(&sh.up_stream, cudaStreamNonBlocking));
    if (!sh.down_stream) CK(cudaStreamCreateWithFlags(&sh.down_stream, cudaStreamNonBlocking));
    if (!sh.alt_stream) CK(cudaStreamCreateWithFlags(&sh.alt_stream, cudaStreamNonBlocking));
    const bool v3 = f3_candidate(h);
    const bool lane = !v3 && lane_wanted(sh, B);
    if (lane) ensure_lane_ws(sh);
    else if (!v3) fused2_prepare(sh.fused2);
    sh.sharedA = (flags & SOCP_FLAG_SHARED_A) != 0;
    sh.sharedG = (flags & SOCP_FLAG_SHARED_G) != 0;
    sh.w.sA = sh.sharedA ? 0 : (int64_t)p * n;
    sh.w.sG = sh.sharedG ? 0 : (int64_t)k * n;
    sh.any_sing = false;
    sh.prepared = false;
    sh.sing_known = sing != nullptr;
    sh.f3_planned = false;
    sh.f3_dense = false;
    sh.have_data = false;
    // Chunk boundaries: two short chunks first (one wave of resident CTAs, then two: PCIe delivers problems about
    // twice as fast as the kernel retires them, so the upload of each next chunk ends before the previous one is
    // solved) so that the solve starts as soon as possible, then up to 8 equal chunks of at least 4 waves each.
    const int slots = v3 ? sh.fused2.num_sms * 4
                         : (lane ? sh.lane.num_sms * sh.lane.pps : sh.fused2.num_sms * sh.fused2.ctas_per_sm);
    // (the lane-per-problem kernel holds 96 problems per SM and fills partial waves evenly: one short chunk, then
    // chunks of two waves)
    std::vector<int> bounds{0};
    if (lane) { if (B > 3 * slots) bounds.push_back(slots); }
    else if (B > 8 * slots) { bounds.push_back(slots); bounds.push_back(3 * slots); }
    {
        const int rest = B - bounds.back();
        const int nrest = std::max(1, std::min(8, rest / std::max(1, (lane ? 2 : 4) * slots)));
        const int per = (rest + nrest - 1) / nrest;
        for (int lo = bounds.back(); lo < B; lo += per) bounds.push_back(std::min(B, lo + per));
    }
    const int nchunk = (int)bounds.size() - 1;
    while ((int)sh.pipe_ev.size() < 2 * nchunk) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        sh.pipe_ev.push_back(e);
    }
    const bool allow_static = getenv("SOCP_B200_GENERIC_ONLY") == nullptr;
    const LoopParams lp{prm.max_iter, prm.tol, prm.step_damp, prm.init_eps};
    auto up = [&](void* dst, const void* src, size_t bytes) {
        if (bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, sh.up_stream));
    };
    auto down = [&](void* dst, const void* src, size_t bytes) {
        if (bytes && dst) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, sh.down_stream));
    };
    auto upload_chunk = [&](int ci) {
        const int lo = bounds[ci], cb = bounds[ci + 1] - lo;
        const int64_t g0 = f + lo;
        up(sh.d_c + (size_t)lo * n, c + g0 * n, sizeof(double) * cb * n);
        up(sh.d_h + (size_t)lo * k, hvec + g0 * k, sizeof(double) * cb * k);
        if (p > 0) up(sh.d_b + (size_t)lo * p, b + g0 * p, sizeof(double) * cb * p);
        if (csc) {
            // only the stored values cross PCIe; the dense operands are assembled on the device (scatter_chunk)
            if (ci == 0) {
                up(sh.d_linG, csc->linG->data(), sizeof(int) * nnzG);
                if (p > 0) up(sh.d_linA, csc->linA->data(), sizeof(int) * nnzA);
            }
            if (p > 0) {
                if (sh.sharedA) { if (ci == 0) up(sh.d_valA, csc->A->nzval, sizeof(double) * nnzA); }
                else up(sh.d_valA + (size_t)lo * nnzA, csc->A->nzval + g0 * nnzA, sizeof(double) * cb * nnzA);
            }
            if (sh.sharedG) { if (ci == 0) up(sh.d_valG, csc->G->nzval, sizeof(double) * nnzG); }
            else up(sh.d_valG + (size_t)lo * nnzG, csc->G->nzval + g0 * nnzG, sizeof(double) * cb * nnzG);
        } else {
            if (p > 0) {
                if (sh.sharedA) { if (ci == 0) up(sh.d_A, A, sizeof(double) * p * n); }
                else up(sh.d_A + (size_t)lo * p * n, A + g0 * p * n, sizeof(double) * cb * p * n);
            }
            if (sh.sharedG) { if (ci == 0) up(sh.d_G, G, sizeof(double) * k * n); }
            else up(sh.d_G + (size_t)lo * k * n, G + g0 * (int64_t)k * n, sizeof(double) * cb * k * n);
        }
        if (sing) up(sh.d_sing + lo, sing + g0, cb);
        CK(cudaEventRecord(sh.pipe_ev[2 * ci], sh.up_stream));
    };
    CK(cudaEventRecord(sh.ev[0], sh.stream));
    CK(cudaStreamWaitEvent(sh.up_stream, sh.ev[0], 0));
    CK(cudaStreamWaitEvent(sh.alt_stream, sh.ev[0], 0));
    sh.launches = 0;
    if (!sing) CK(cudaMemsetAsync(sh.d_sing, 0, B, sh.up_stream));
    if (v3) CK(cudaMemsetAsync(sh.d_npattern, 0, sizeof(int), sh.up_stream));
    upload_chunk(0);
    // Consecutive chunks alternate between two compute streams: the persistent CTAs of chunk i+1 move in as those
    // of chunk i run out of work, so there is no drain bubble between launches.
    for (int ci = 0; ci < nchunk; ++ci) {
        cudaStream_t cs = (ci & 1) ? sh.alt_stream : sh.stream;
        const int lo = bounds[ci], cb = bounds[ci + 1] - lo;
        const int64_t g0 = f + lo;
        if (ci + 1 < nchunk) upload_chunk(ci + 1);        // the copy engine stays one chunk ahead
        CK(cudaStreamWaitEvent(cs, sh.pipe_ev[2 * ci], 0));
        if (csc) {
            // dense column-major operands of this chunk from the stored values (zero fill + scatter, on the chunk's
            // compute stream: ordered after the upload, before the solve)
            auto scatter_chunk = [&](double* dense, const double* val, const int* lin, int64_t nnz, int rows, bool shared) {
                if (shared && ci > 0) return;
                const int64_t nb = shared ? 1 : cb, off = shared ? 0 : lo;
                CK(cudaMemsetAsync(dense + (size_t)off * rows * n, 0, sizeof(double) * nb * rows * n, cs));
                const int64_t total = nb * nnz;
                if (total == 0) return;
                const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
                k_csc_scatter<<<grid, 256, 0, cs>>>(lin, val + (size_t)off * nnz, nnz, (int)nnz, dense + (size_t)off * rows * n,
                                                    (int64_t)rows * n, total);
                CK(cudaPeekAtLastError());
                sh.launches += 1;
            };
            if (p > 0) scatter_chunk(sh.d_A, sh.d_valA, sh.d_linA, nnzA, p, sh.sharedA);
            scatter_chunk(sh.d_G, sh.d_valG, sh.d_linG, nnzG, k, sh.sharedG);
            if (nchunk > 1 && (sh.sharedA || sh.sharedG) && ci == 0) {
                // later chunks run on the other stream as well: they must see the shared operand
                CK(cudaEventRecord(sh.ev[2], cs));
                CK(cudaStreamWaitEvent(sh.alt_stream, sh.ev[2], 0));
            }
        }
        if (v3) {
            if (ci == 0 && csc) {
                plan_fused3(h, sh, *csc->rowcolG, cs);          // the pattern is the CSC pattern: nothing to detect or verify
                if (!sh.fused3.fits) {
                    CK(cudaStreamSynchronize(sh.up_stream));
                    CK(cudaStreamSynchronize(cs));
                    return false;
                }
            } else if (ci == 0) {
                // the row pattern of G is taken from the first chunk (one small kernel and a host round trip while the
                // second chunk uploads); the later chunks are verified against it by the solve kernel itself
                detect_and_plan_fused3(h, sh, 0, cb, cs);
                if (!sh.fused3.fits) {
                    CK(cudaStreamSynchronize(sh.up_stream));
                    return false;
                }
            }
            solve_fused3_ext(sh.fused3, f3_glob(sh), lo, cb, lp, sing ? 0 : 1, ci > 0 && !sh.sharedG && !csc ? 1 : 0, cs, allow_static, ci & 15);
        } else if (lane) {
            solve_fused_lane_ext(sh.lane, sh.w, lo, cb, lp, cs, ci);      // adjacent chunks may overlap: alternate workspace sets
        } else {
            solve_fused2_ext(sh.fused2, sh.w, lo, cb, prm.max_iter, prm.tol, prm.step_damp, prm.init_eps, cs, allow_static, ci & 15);
        }
        CK(cudaGetLastError());
        sh.launches += 1;
        CK(cudaEventRecord(sh.pipe_ev[2 * ci + 1], cs));
        CK(cudaStreamWaitEvent(sh.down_stream, sh.pipe_ev[2 * ci + 1], 0));
        if (x) down(x + g0 * n, sh.w.x + (size_t)lo * n, sizeof(double) * cb * n);
        if (y) down(y + g0 * p, sh.w.y + (size_t)lo * p, sizeof(double) * cb * p);
        if (z) down(z + g0 * k, sh.w.z + (size_t)lo * k, sizeof(double) * cb * k);
        if (s) down(s + g0 * k, sh.w.s + (size_t)lo * k, sizeof(double) * cb * k);
        if (status) down(status + g0, sh.w.status + lo, sizeof(int) * cb);
        if (iters) down(iters + g0, sh.w.iters + lo, sizeof(int) * cb);
        if (pobj) down(pobj + g0, sh.w.pobj + lo, sizeof(double) * cb);
        if (dobj) down(dobj + g0, sh.w.dobj + lo, sizeof(double) * cb);
    }
    if (v3) CK(cudaMemcpyAsync(sh.h_rowcol + k, sh.d_npattern, sizeof(int), cudaMemcpyDeviceToHost, sh.down_stream));
    if (nchunk > 1) CK(cudaStreamWaitEvent(sh.stream, sh.pipe_ev[2 * (nchunk - 1) + 1], 0));   // join the alternate stream
    CK(cudaEventRecord(sh.ev[1], sh.stream));
    CK(cudaStreamSynchronize(sh.up_stream));
    CK(cudaStreamSynchronize(sh.alt_stream));
    CK(cudaStreamSynchronize(sh.stream));
    CK(cudaStreamSynchronize(sh.down_stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, sh.ev[0], sh.ev[1]));
    sh.tim.h2d_ms = 0;
    sh.tim.solve_ms = ms;          // upload + solve, overlapped
    sh.tim.d2h_ms = 0;
    sh.tim.kernel_launches = sh.launches;
    sh.tim.path_used = SOCP_PATH_FUSED;
    sh.tim.iterations_max = -1;
    sh.have_data = true;
    sh.have_scaling = sh.have_factor = false;
    if (v3 && !sing) sh.sing_known = true;          // found by the kernel, left in d_sing
    if (v3 && sh.h_rowcol[k] > 0) {
        // some problem of a later chunk has nonzeros outside the first chunk's pattern: the data is resident, so solve
        // the shard again on the all-dense plan and fetch everything (rare; correctness over speed)
        sh.f3_dense = true;
        sh.f3_planned = false;
        if (!sing) {           // the reported problems were not tested for `sing`: start over
            sh.sing_known = false;
            CK(cudaMemsetAsync(sh.d_sing, 0, B, sh.stream));
        }
        run_solve(h, sh, prm);
        fetch_results(h, sh, x, y, z, s, status, iters, pobj, dobj);
    }
    return true;
}

// what follows the upload in both constructors: bookkeeping; the `sing` test and A'A are left to ensure_prepared
// (only the tiled path and get_sing need them).  rowcol: the row pattern of G when the caller knows it (CSC upload).
void finish_set_data(const socp_handle* h, Shard& sh, const uint8_t* sing, const std::vector<int>* rowcol) {
    const int B = sh.batch;
    sh.sing_known = sing != nullptr;
    sh.prepared = false;
    sh.any_sing = false;
    sh.f3_planned = false;
    sh.f3_dense = false;
    if (!sing) CK(cudaMemsetAsync(sh.d_sing, 0, B, sh.stream));
    CK(cudaStreamSynchronize(sh.stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, sh.ev[0], sh.ev[1]));
    sh.tim.h2d_ms = ms;
    sh.have_data = true;
    sh.have_scaling = sh.have_factor = false;
    if (f3_candidate(h)) {
        if (rowcol) plan_fused3(h, sh, *rowcol, sh.stream);
        else detect_and_plan_fused3(h, sh, 0, B, sh.stream);
    }
}

// host side of the pattern: SparseMatrixCSC invariants -> linear indices into a rows x cols column-major matrix
std::vector<int> csc_linear_index(const socp_csc& m, int rows, int cols, const char* name) {
    need(m.nnz >= 0 && m.nnz <= (int64_t)rows * cols, SOCP_ERR_SIZE, "csc: nnz out of range");
    need(m.colptr && (m.nnz == 0 || (m.rowval && m.nzval)), SOCP_ERR_NULL, "csc: colptr / rowval / nzval must not be null");
    need(m.index_base == 0 || m.index_base == 1, SOCP_ERR_LAYOUT, "csc: index_base must be 0 or 1");
    const int64_t base = m.index_base;
    need(m.colptr[0] == base && m.colptr[cols] == base + m.nnz, SOCP_ERR_LAYOUT, "csc: colptr does not span nnz");
    std::vector<int> lin((size_t)m.nnz);
    for (int cidx = 0; cidx < cols; ++cidx) {
        const int64_t lo = m.colptr[cidx] - base, hi = m.colptr[cidx + 1] - base;
        need(lo <= hi && hi <= m.nnz, SOCP_ERR_LAYOUT, "csc: colptr not monotone");
        int64_t prev = -1;
        for (int64_t j = lo; j < hi; ++j) {
            const int64_t r = m.rowval[j] - base;
            need(r > prev && r < rows, SOCP_ERR_LAYOUT, "csc: row indices must increase strictly inside a column");
            prev = r;
            lin[(size_t)j] = (int)((int64_t)cidx * rows + r);
        }
    }
    (void)name;
    return lin;
}

// upload one matrix of a shard in CSC form and assemble it dense in `dense`
void upload_csc(Shard& sh, const socp_csc& m, const std::vector<int>& lin, bool shared, int rows, int cols,
                double* dense) {
    const int64_t nb = shared ? 1 : sh.batch;
    CK(cudaMemsetAsync(dense, 0, sizeof(double) * nb * rows * cols, sh.stream));
    if (m.nnz == 0) return;
    struct Staging {       // freed on every exit path (a CK throw included)
        int* lin = nullptr;
        double* val = nullptr;
        ~Staging() { if (lin) cudaFree(lin); if (val) cudaFree(val); }
    } st;
    CK(cudaMalloc(&st.lin, sizeof(int) * m.nnz));
    if (cudaMalloc(&st.val, sizeof(double) * nb * m.nnz) != cudaSuccess) {
        cudaGetLastError();
        throw UsageErr{SOCP_ERR_NOMEM, "csc: out of device memory for the value staging buffer"};
    }
    int* d_lin = st.lin;
    double* d_val = st.val;
    h2d(sh, d_lin, lin.data(), sizeof(int) * m.nnz);
    h2d(sh, d_val, m.nzval + (shared ? 0 : sh.first * m.nnz), sizeof(double) * nb * m.nnz);
    const int64_t total = nb * m.nnz;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
    LAUNCH(sh, k_csc_scatter, grid, 256, 0, d_lin, d_val, m.nnz, (int)m.nnz, dense, (int64_t)rows * cols, total);
    CK(cudaStreamSynchronize(sh.stream));
}

}  // namespace

// =========================================================================== C ABI
extern "C" {

int socp_b200_version(void) { return SOCP_B200_VERSION; }

int socp_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void socp_b200_default_params(socp_params* out) {
    if (!out) return;
    out->max_iter = 40;      // reference src/solver.jl:105
    out->path = SOCP_PATH_AUTO;
    out->tol = 1e-5;         // :122
    out->step_damp = 0.99;   // :146
    out->init_eps = 1e-10;   // :91,:97
}

static thread_local std::string g_create_err;

int socp_b200_create(socp_handle** out, const socp_layout* L, int64_t batch, const int32_t* devices, int32_t ndev) {
    if (!out || !L) { g_create_err = "null argument"; return SOCP_ERR_NULL; }
    *out = nullptr;
    if (L->n <= 0 || L->p < 0 || L->k <= 0 || L->ncones <= 0 || batch <= 0) { g_create_err = "bad dimensions"; return SOCP_ERR_LAYOUT; }
    if (!L->cone_kind || !L->cone_offs || !L->cone_dim) { g_create_err = "null cone arrays"; return SOCP_ERR_NULL; }
    int off = 0;
    bool seen_soc = false;
    for (int i = 0; i < L->ncones; ++i) {
        const int kd = L->cone_kind[i];
        if (kd != SOCP_CONE_POC && kd != SOCP_CONE_SOC) { g_create_err = "unknown cone kind"; return SOCP_ERR_LAYOUT; }
        if (L->cone_dim[i] <= 0 || L->cone_offs[i] != off) { g_create_err = "cones must tile 0..k-1 contiguously"; return SOCP_ERR_LAYOUT; }
        if (kd == SOCP_CONE_SOC) seen_soc = true;
        else if (seen_soc) { g_create_err = "POC blocks must precede SOC blocks"; return SOCP_ERR_LAYOUT; }
        off += L->cone_dim[i];
    }
    if (off != L->k) { g_create_err = "cone dims do not sum to k"; return SOCP_ERR_LAYOUT; }
    if (L->n > 8192 || L->p > 8192) { g_create_err = "n, p limited to 8192 (triangular-solve staging)"; return SOCP_ERR_SIZE; }
    socp_handle* h = new socp_handle();
    h->n = L->n; h->p = L->p; h->k = L->k; h->batch = batch;
    for (int i = 0; i < L->ncones; ++i) {
        h->kind.push_back(L->cone_kind[i]);
        h->offs.push_back(L->cone_offs[i]);
        h->dim.push_back(L->cone_dim[i]);
        h->first_work.push_back((int)h->wkind.size());
        if (L->cone_kind[i] == SOCP_CONE_POC) {
            h->deg += L->cone_dim[i];
            for (int o = 0; o < L->cone_dim[i]; o += POC_CHUNK) {
                h->wkind.push_back(KIND_POC);
                h->woffs.push_back(L->cone_offs[i] + o);
                h->wdim.push_back(std::min(POC_CHUNK, L->cone_dim[i] - o));
            }
        } else {
            h->deg += 1;
            h->wkind.push_back(KIND_SOC);
            h->woffs.push_back(L->cone_offs[i]);
            h->wdim.push_back(L->cone_dim[i]);
        }
    }
    std::vector<int> devs;
    if (devices && ndev > 0) devs.assign(devices, devices + ndev);
    else {
        int cur = 0;
        if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); cur = 0; }
        devs.push_back(cur);
    }
    {
        int ndevices = 0;
        if (cudaGetDeviceCount(&ndevices) != cudaSuccess) { cudaGetLastError(); ndevices = 0; }
        for (int d : devs)
            if (d < 0 || d >= ndevices || d >= 64) {
                g_create_err = "device id out of range (0 <= id < cudaGetDeviceCount(), at most 64 devices)";
                delete h;
                return ndevices == 0 ? (int)cudaErrorNoDevice : SOCP_ERR_LAYOUT;
            }
    }
    if ((int64_t)devs.size() > batch) devs.resize((size_t)batch);
    const int64_t per = (batch + (int64_t)devs.size() - 1) / (int64_t)devs.size();
    int64_t first = 0;
    for (size_t d = 0; d < devs.size() && first < batch; ++d) {
        Shard sh;
        sh.device = devs[d];
        sh.first = first;
        sh.batch = (int)std::min<int64_t>(per, batch - first);
        first += sh.batch;
        h->shards.push_back(sh);
    }
    int rc = guarded(h, [&]() {
        for (auto& sh : h->shards) build_shard(h, sh);
    });
    if (rc != 0) {
        g_create_err = h->err;
        for (auto& sh : h->shards) sh.release();
        delete h;
        return rc;
    }
    *out = h;
    return 0;
}

int socp_b200_destroy(socp_handle* h) {
    if (!h) return SOCP_ERR_NULL;
    for (auto& sh : h->shards) sh.release();
    delete h;
    return 0;
}

const char* socp_b200_last_error(const socp_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int socp_b200_set_data(socp_handle* h, const double* c, const double* A, const double* b, const double* G,
                       const double* hvec, const uint8_t* sing, int32_t flags) {
    if (!h) return SOCP_ERR_NULL;
    return guarded(h, [&]() {
        need(c && G && hvec, SOCP_ERR_NULL, "c, G, h must not be null");
        need(h->p == 0 || (A && b), SOCP_ERR_NULL, "A, b must not be null when p > 0");
        const int n = h->n, p = h->p, k = h->k;
        for_each_shard(h, [&](Shard& sh) {
            const int64_t f = sh.first;
            const int B = sh.batch;
            sh.sharedA = (flags & SOCP_FLAG_SHARED_A) != 0;
            sh.sharedG = (flags & SOCP_FLAG_SHARED_G) != 0;
            CK(cudaEventRecord(sh.ev[0], sh.stream));
            h2d(sh, sh.d_c, c + f * n, sizeof(double) * B * n);
            h2d(sh, sh.d_h, hvec + f * k, sizeof(double) * B * k);
            if (p > 0) {
                h2d(sh, sh.d_b, b + f * p, sizeof(double) * B * p);
                if (sh.sharedA) h2d(sh, sh.d_A, A, sizeof(double) * p * n);
                else h2d(sh, sh.d_A, A + f * p * n, sizeof(double) * B * p * n);
            }
            if (sh.sharedG) h2d(sh, sh.d_G, G, sizeof(double) * k * n);
            else h2d(sh, sh.d_G, G + f * (int64_t)k * n, sizeof(double) * B * k * n);
            sh.w.sA = sh.sharedA ? 0 : (int64_t)p * n;
            sh.w.sG = sh.sharedG ? 0 : (int64_t)k * n;
            if (sing) h2d(sh, sh.d_sing, sing + f, B);
            CK(cudaEventRecord(sh.ev[1], sh.stream));
            finish_set_data(h, sh, sing, nullptr);
        });
    });
}

int socp_b200_set_data_csc(socp_handle* h, const double* c, const socp_csc* A, const double* b, const socp_csc* G,
                           const double* hvec, const uint8_t* sing, int32_t flags) {
    if (!h) return SOCP_ERR_NULL;
    return guarded(h, [&]() {
        need(c && G && hvec, SOCP_ERR_NULL, "c, G, h must not be null");
        need(h->p == 0 || (A && b), SOCP_ERR_NULL, "A, b must not be null when p > 0");
        const int n = h->n, p = h->p, k = h->k;
        const std::vector<int> linG = csc_linear_index(*G, k, n, "G");
        const std::vector<int> linA = p > 0 ? csc_linear_index(*A, p, n, "A") : std::vector<int>();
        // row pattern of G from the stored entries: one entry in a row = singleton, none = empty (fused_v3.cuh)
        std::vector<int> rowcolG(k, -1);
        for (int li : linG) {
            const int i = li % k, j = li / k;
            rowcolG[i] = rowcolG[i] == -1 ? j : -2;
        }
        for_each_shard(h, [&](Shard& sh) {
            const int64_t f = sh.first;
            const int B = sh.batch;
            sh.sharedA = (flags & SOCP_FLAG_SHARED_A) != 0;
            sh.sharedG = (flags & SOCP_FLAG_SHARED_G) != 0;
            CK(cudaEventRecord(sh.ev[0], sh.stream));
            h2d(sh, sh.d_c, c + f * n, sizeof(double) * B * n);
            h2d(sh, sh.d_h, hvec + f * k, sizeof(double) * B * k);
            if (p > 0) {
                h2d(sh, sh.d_b, b + f * p, sizeof(double) * B * p);
                upload_csc(sh, *A, linA, sh.sharedA, p, n, sh.d_A);
            }
            upload_csc(sh, *G, linG, sh.sharedG, k, n, sh.d_G);
            sh.w.sA = sh.sharedA ? 0 : (int64_t)p * n;
            sh.w.sG = sh.sharedG ? 0 : (int64_t)k * n;
            if (sing) h2d(sh, sh.d_sing, sing + f, B);
            CK(cudaEventRecord(sh.ev[1], sh.stream));
            finish_set_data(h, sh, sing, &rowcolG);
        });
    });
}

int socp_b200_solve_dev(socp_handle* h, const socp_params* params) {
    if (!h) return SOCP_ERR_NULL;
    return guarded(h, [&]() {
        socp_params prm;
        socp_b200_default_params(&prm);
        if (params) prm = *params;
        need(prm.max_iter >= 0 && prm.max_iter <= 4000, SOCP_ERR_SIZE, "max_iter out of range");
        for_each_shard(h, [&](Shard& sh) { run_solve(h, sh, prm); });
    });
}

int socp_b200_get_results(socp_handle* h, double* x, double* y, double* z, double* s, int32_t* status,
                          int32_t* iters, double* pobj, double* dobj) {
    if (!h) return SOCP_ERR_NULL;
    return guarded(h, [&]() {
        for_each_shard(h, [&](Shard& sh) { fetch_results(h, sh, x, y, z, s, status, iters, pobj, dobj); });
    });
}

int socp_b200_solve(socp_handle* h, const socp_params* params, double* x, double* y, double* z, double* s,
                    int32_t* status, int32_t* iters, double* pobj, double* dobj) {
    if (!h) return SOCP_ERR_NULL;
    return guarded(h, [&]() {
        socp_params prm;
        socp_b200_default_params(&prm);
        if (params) prm = *params;
        need(prm.max_iter >= 0 && prm.max_iter <= 4000, SOCP_ERR_SIZE, "max_iter out of range");
        for_each_shard(h, [&](Shard& sh) {
            run_solve(h, sh, prm);
            fetch_results(h, sh, x, y, z, s, status, iters, pobj, dobj);
        });
    });
}

int socp_b200_solve_host(socp_handle* h, const socp_params* params, const double* c, const double* A, const double* b,
                         const double* G, const double* hvec, const uint8_t* sing, int32_t flags, double* x, double* y,
                         double* z, double* s, int32_t* status, int32_t* iters, double* pobj, double* dobj) {
    if (!h) return SOCP_ERR_NULL;
    socp_params prm;
    socp_b200_default_params(&prm);
    if (params) prm = *params;
    bool pipelined = true;
    int rc = guarded(h, [&]() {
        need(c && G && hvec, SOCP_ERR_NULL, "c, G, h must not be null");
        need(h->p == 0 || (A && b), SOCP_ERR_NULL, "A, b must not be null when p > 0");
        need(prm.max_iter >= 0 && prm.max_iter <= 4000, SOCP_ERR_SIZE, "max_iter out of range");
        for (auto& sh : h->shards) pipelined &= can_pipeline(h, sh, prm, sing, sh.first);
        if (!pipelined) return;
        std::vector<char> okv(h->shards.size(), 1);
        for_each_shard(h, [&](Shard& sh) {
            okv[&sh - h->shards.data()] =
                run_pipelined(h, sh, prm, c, A, b, G, hvec, sing, flags, x, y, z, s, status, iters, pobj, dobj) ? 1 : 0;
        });
        for (char v : okv) pipelined &= (v != 0);
    });
    if (rc != 0 || pipelined) return rc;
    rc = socp_b200_set_data(h, c, A, b, G, hvec, sing, flags);
    if (rc != 0) return rc;
    return socp_b200_solve(h, &prm, x, y, z, s, status, iters, pobj, dobj);
}

int socp_b200_solve_host_csc(socp_handle* h, const socp_params* params, const double* c, const socp_csc* A,
                             const double* b, const socp_csc* G, const double* hvec, const uint8_t* sing, int32_t flags,
                             double* x, double* y, double* z, double* s, int32_t* status, int32_t* iters, double* pobj,
                             double* dobj) {
    if (!h) return SOCP_ERR_NULL;
    socp_params prm;
    socp_b200_default_params(&prm);
    if (params) prm = *params;
    bool pipelined = true;
    int rc = guarded(h, [&]() {
        need(c && G && hvec, SOCP_ERR_NULL, "c, G, h must not be null");
        need(h->p == 0 || (A && b), SOCP_ERR_NULL, "A, b must not be null when p > 0");
        need(prm.max_iter >= 0 && prm.max_iter <= 4000, SOCP_ERR_SIZE, "max_iter out of range");
        for (auto& sh : h->shards) pipelined &= can_pipeline(h, sh, prm, sing, sh.first);
        if (!pipelined) return;
        const int n = h->n, p = h->p, k = h->k;
        const std::vector<int> linG = csc_linear_index(*G, k, n, "G");
        const std::vector<int> linA = p > 0 ? csc_linear_index(*A, p, n, "A") : std::vector<int>();
        std::vector<int> rowcolG(k, -1);
        for (int li : linG) {
            const int i = li % k, j = li / k;
            rowcolG[i] = rowcolG[i] == -1 ? j : -2;
        }
        const CscSrc src{A, G, &linA, &linG, &rowcolG};
        std::vector<char> okv(h->shards.size(), 1);
        for_each_shard(h, [&](Shard& sh) {
            okv[&sh - h->shards.data()] = run_pipelined(h, sh, prm, c, nullptr, b, nullptr, hvec, sing, flags, x, y, z, s,
                                                        status, iters, pobj, dobj, &src) ? 1 : 0;
        });
        for (char v : okv) pipelined &= (v != 0);
    });
    if (rc != 0 || pipelined) return rc;
    rc = socp_b200_set_data_csc(h, c, A, b, G, hvec, sing, flags);
    if (rc != 0) return rc;
    return socp_b200_solve(h, &prm, x, y, z, s, status, iters, pobj, dobj);
}

int socp_b200_get_sing(socp_handle* h, uint8_t* sing) {
    if (!h || !sing) return SOCP_ERR_NULL;
    return guarded(h, [&]() {
        for_each_shard(h, [&](Shard& sh) {
            need(sh.have_data, SOCP_ERR_STATE, "get_sing called before set_data");
            ensure_prepared(sh);
            d2h(sh, sing + sh.first, sh.d_sing, sh.batch);
            CK(cudaStreamSynchronize(sh.stream));
        });
    });
}

int socp_b200_timings(const socp_handle* h, socp_timings* out) {
    if (!h || !out) return SOCP_ERR_NULL;
    socp_timings t{};
    for (const auto& sh : h->shards) {
        t.h2d_ms = std::max(t.h2d_ms, sh.tim.h2d_ms);
        t.solve_ms = std::max(t.solve_ms, sh.tim.solve_ms);
        t.d2h_ms = std::max(t.d2h_ms, sh.tim.d2h_ms);
        t.kernel_launches += sh.tim.kernel_launches;
        t.iterations_max = std::max(t.iterations_max, sh.tim.iterations_max);
        t.path_used = sh.tim.path_used;
    }
    *out = t;
    return 0;
}

// ------------------------------------------------------------------ step level
#define STEP_PROLOGUE(cond_state, msg)                                        \
    if (!h) return SOCP_ERR_NULL;                                             \
    return guarded(h, [&]() {                                                 \
        const int n = h->n, p = h->p, k = h->k;                               \
        (void)n; (void)p; (void)k;                                            \
        for_each_shard(h, [&](Shard& sh) {                                    \
            Ws& w = sh.w;                                                     \
            const int64_t f = sh.first;                                       \
            const int B = sh.batch;                                           \
            (void)w; (void)f; (void)B;                                        \
            need(cond_state, SOCP_ERR_STATE, msg);

#define STEP_EPILOGUE                                                         \
            CK(cudaStreamSynchronize(sh.stream));                             \
        });                                                                   \
    });

int socp_b200_compute_scaling(socp_handle* h, const double* s, const double* z, double* lambda, double* wbs,
                              double* mu, int32_t* fail) {
    if (h && (!s || !z)) return SOCP_ERR_NULL;
    STEP_PROLOGUE(true, "")
        h2d(sh, w.s, s + f * k, sizeof(double) * B * k);
        h2d(sh, w.z, z + f * k, sizeof(double) * B * k);
        CK(cudaMemsetAsync(w.fail, 0, sizeof(int) * B, sh.stream));
        launch_scaling(sh, nullptr);
        if (lambda) d2h(sh, lambda + f * k, w.lam, sizeof(double) * B * k);
        if (wbs) d2h(sh, wbs + f * k, w.wb, sizeof(double) * B * k);
        if (fail) d2h(sh, fail + f, w.fail, sizeof(int) * B);
        std::vector<double> eta;
        const int nc = w.L.ncones;
        if (mu) { eta.resize((size_t)B * 4 * nc); d2h(sh, eta.data(), w.eta, sizeof(double) * B * 4 * nc); }
        CK(cudaStreamSynchronize(sh.stream));
        if (mu) {
            const int no = (int)h->kind.size();
            for (int b = 0; b < B; ++b)
                for (int c = 0; c < no; ++c)
                    mu[(f + b) * no + c] = h->kind[c] == SOCP_CONE_SOC ? eta[(size_t)b * 4 * nc + h->first_work[c]] : 0.0;
        }
        sh.have_scaling = true;
        sh.have_factor = false;
    STEP_EPILOGUE
}

int socp_b200_setup_iter(socp_handle* h, int32_t* fail) {
    STEP_PROLOGUE(sh.have_data && sh.have_scaling, "setup_iter needs set_data and compute_scaling first")
        ensure_prepared(sh);
        ensure_tiled(sh);
        CK(cudaMemsetAsync(w.fail, 0, sizeof(int) * B, sh.stream));
        factor(sh, false, true, nullptr);
        if (fail) d2h(sh, fail + f, w.fail, sizeof(int) * B);
        sh.have_factor = true;
    STEP_EPILOGUE
}

int socp_b200_solve_kkt(socp_handle* h, const double* dx, const double* dy, const double* dz, const double* ds,
                        double* cx, double* cy, double* cz, double* cs) {
    if (h && (!dx || !dz || !ds || (h->p > 0 && !dy))) return SOCP_ERR_NULL;
    STEP_PROLOGUE(sh.have_factor, "solve_kkt needs setup_iter first")
        h2d(sh, w.dx, dx + f * n, sizeof(double) * B * n);
        if (p > 0) h2d(sh, w.dy, dy + f * p, sizeof(double) * B * p);
        h2d(sh, w.dz, dz + f * k, sizeof(double) * B * k);
        h2d(sh, w.ds, ds + f * k, sizeof(double) * B * k);
        LAUNCH(sh, k_kkt_head, B, sh.threads, 0, w);
        kkt_middle(sh, nullptr);
        LAUNCH(sh, k_kkt_tail, B, sh.threads, 0, w);
        if (cx) d2h(sh, cx + f * n, w.rx, sizeof(double) * B * n);
        if (cy && p > 0) d2h(sh, cy + f * p, w.ry, sizeof(double) * B * p);
        if (cz) d2h(sh, cz + f * k, w.rz, sizeof(double) * B * k);
        if (cs) d2h(sh, cs + f * k, w.rs, sizeof(double) * B * k);
    STEP_EPILOGUE
}

static int apply_common(socp_handle* h, const double* in, double* out, int mode) {
    if (h && (!in || !out)) return SOCP_ERR_NULL;
    STEP_PROLOGUE(sh.have_scaling, "needs compute_scaling first")
        h2d(sh, w.kt2, in + f * k, sizeof(double) * B * k);
        launch_apply(sh, mode, w.kt2, w.kt3);
        d2h(sh, out + f * k, w.kt3, sizeof(double) * B * k);
    STEP_EPILOGUE
}
int socp_b200_scale(socp_handle* h, const double* in, double* out) { return apply_common(h, in, out, 0); }
int socp_b200_iscale(socp_handle* h, const double* in, double* out) { return apply_common(h, in, out, 1); }
int socp_b200_iwiw(socp_handle* h, const double* in, double* out) { return apply_common(h, in, out, 2); }

// SqrScaling form of the scaling now in the handle (reference src/sqrscalings.jl:50-58 positive orthant, :98-128
// second-order cone): W^-2 restricted to one cone = diag(D) + u u' - v v'.  One thread per (problem, work cone); the
// O(d) passes over a cone run on that thread (this is an interface of the reference's sparse assembly, not a hot
// kernel: the dense path multiplies by W^-2 in closed form and never forms D, u, v).
__global__ void k_sqr_scaling(ConeLayout L, int nbatch, const double* __restrict__ wb, const double* __restrict__ iwb,
                              const double* __restrict__ eta, double* __restrict__ D, double* __restrict__ u,
                              double* __restrict__ v) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int nc = L.ncones;
    if (t >= (int64_t)nbatch * nc) return;
    const int b = (int)(t / nc), c = (int)(t - (int64_t)b * nc);
    const int kind = L.kind[c], offs = L.offs[c], dim = L.dim[c];
    const int64_t o = (int64_t)b * L.k + offs;
    if (kind == KIND_POC) {
        for (int i = 0; i < dim; ++i) {
            const double iw = iwb[o + i];
            D[o + i] = iw * iw;                                       // z / s                       :52
            u[o + i] = 0.0;
            v[o + i] = 0.0;
        }
        return;
    }
    const double* e4 = eta + (int64_t)b * 4 * nc;
    const double inu = e4[nc + c], inusq = e4[2 * nc + c];             // 1/eta, 1/eta^2             :96-97
    const double wb0 = wb[o];
    double wb1sq = 0.0;
    for (int i = 1; i < dim; ++i) wb1sq = fma(wb[o + i], wb[o + i], wb1sq);
    const double cv = -(1.0 + wb0 + wb1sq / (1.0 + wb0));                                          // :101
    const double d = 1.0 + 2.0 / (1.0 + wb0) + wb1sq / ((1.0 + wb0) * (1.0 + wb0));                // :102
    const double a = (wb0 * wb0 + wb1sq - cv * cv * wb1sq / (1.0 + d * wb1sq)) / 2.0;              // :103
    const double u0 = sqrt(wb0 * wb0 + wb1sq - a);                                                 // :104
    const double u1 = cv / u0;                                                                     // :105
    const double v1 = sqrt(cv * cv / (u0 * u0) - d);                                               // :106
    D[o] = a * inusq;                                                                              // :108
    u[o] = inu * u0;                                                                               // :120
    v[o] = 0.0;
    for (int i = 1; i < dim; ++i) {
        const double wbv = inu * wb[o + i];
        D[o + i] = inusq;                                                                          // :111-113
        u[o + i] = u1 * wbv;                                                                       // :122-127
        v[o + i] = v1 * wbv;
    }
}

int socp_b200_sqr_scaling(socp_handle* h, double* D, double* u, double* v) {
    if (h && (!D || !u || !v)) return SOCP_ERR_NULL;
    STEP_PROLOGUE(sh.have_scaling, "sqr_scaling needs compute_scaling first")
        const int64_t total = (int64_t)B * w.L.ncones;
        LAUNCH(sh, k_sqr_scaling, (int)((total + 127) / 128), 128, 0, w.L, B, w.wb, w.iwb, w.eta, w.kt2, w.kt3, w.k0);
        d2h(sh, D + f * k, w.kt2, sizeof(double) * B * k);
        d2h(sh, u + f * k, w.kt3, sizeof(double) * B * k);
        d2h(sh, v + f * k, w.k0, sizeof(double) * B * k);
    STEP_EPILOGUE
}

int socp_b200_make_e(socp_handle* h, double* out) {
    if (h && !out) return SOCP_ERR_NULL;
    STEP_PROLOGUE(true, "")
        LAUNCH(sh, k_make_e, B, sh.threads, 0, w.L, w.kt3);
        d2h(sh, out + f * k, w.kt3, sizeof(double) * B * k);
    STEP_EPILOGUE
}
int socp_b200_vprod(socp_handle* h, const double* u, const double* v, double* out) {
    if (h && (!u || !v || !out)) return SOCP_ERR_NULL;
    STEP_PROLOGUE(true, "")
        h2d(sh, w.kt2, u + f * k, sizeof(double) * B * k);
        h2d(sh, w.kt3, v + f * k, sizeof(double) * B * k);
        launch_vprod(sh, w.kt2, w.kt3, w.k0);
        d2h(sh, out + f * k, w.k0, sizeof(double) * B * k);
    STEP_EPILOGUE
}
int socp_b200_iprod(socp_handle* h, const double* lambda, const double* v, double* out) {
    if (h && (!lambda || !v || !out)) return SOCP_ERR_NULL;
    STEP_PROLOGUE(true, "")
        h2d(sh, w.kt2, lambda + f * k, sizeof(double) * B * k);
        h2d(sh, w.kt3, v + f * k, sizeof(double) * B * k);
        launch_iprod(sh, w.kt2, w.kt3, w.k0);
        d2h(sh, out + f * k, w.k0, sizeof(double) * B * k);
    STEP_EPILOGUE
}
int socp_b200_max_step(socp_handle* h, const double* x, double* out) {
    if (h && (!x || !out)) return SOCP_ERR_NULL;
    STEP_PROLOGUE(true, "")
        h2d(sh, w.kt2, x + f * k, sizeof(double) * B * k);
        launch_max_step(sh, w.kt2, w.k0);
        d2h(sh, out + f, w.k0, sizeof(double) * B);
    STEP_EPILOGUE
}
int socp_b200_compute_step(socp_handle* h, const double* lambda, const double* ds, const double* dz, double* out) {
    if (h && (!lambda || !ds || !dz || !out)) return SOCP_ERR_NULL;
    STEP_PROLOGUE(true, "")
        h2d(sh, w.k2, lambda + f * k, sizeof(double) * B * k);
        h2d(sh, w.kt2, ds + f * k, sizeof(double) * B * k);
        h2d(sh, w.kt3, dz + f * k, sizeof(double) * B * k);
        launch_compute_step(sh, w.k2, w.kt2, w.kt3, w.k0);
        d2h(sh, out + f, w.k0, sizeof(double) * B);
    STEP_EPILOGUE
}

static int get_mat(socp_handle* h, double* out, bool lower_only) {
    if (h && !out) return SOCP_ERR_NULL;
    STEP_PROLOGUE(sh.have_factor, "needs setup_iter first")
        std::vector<double> tmp((size_t)B * w.ldh * n);
        d2h(sh, tmp.data(), w.H, sizeof(double) * tmp.size());
        CK(cudaStreamSynchronize(sh.stream));
        for (int b = 0; b < B; ++b)
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) {
                    const double lo = tmp[((size_t)b * n + std::min(i, j)) * w.ldh + std::max(i, j)];
                    out[((f + b) * n + j) * n + i] = lower_only ? (i >= j ? lo : 0.0) : lo;
                }
    STEP_EPILOGUE
}
// NB: the Cholesky is in place, so after setup_iter the buffer holds L; get_H
// rebuilds H = L L' on the host from it (debug only).
int socp_b200_get_L(socp_handle* h, double* out) { return get_mat(h, out, true); }
int socp_b200_get_H(socp_handle* h, double* out) {
    if (!h || !out) return SOCP_ERR_NULL;
    const int n = h->n;
    std::vector<double> L((size_t)h->batch * n * n);
    int rc = get_mat(h, L.data(), true);
    if (rc) return rc;
    for (int64_t b = 0; b < h->batch; ++b) {
        const double* Lb = L.data() + b * n * n;
        double* Hb = out + b * n * n;
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                double acc = 0.0;
                for (int c = 0; c <= std::min(i, j); ++c) acc += Lb[(size_t)c * n + i] * Lb[(size_t)c * n + j];
                Hb[(size_t)j * n + i] = acc;
            }
    }
    return 0;
}

// Parity aid (not a reference interface): one Mehrotra step of the fused whole-solve kernel, taken out of the middle
// of a real solve.  Problem `index` is solved again on its own; at iteration `iter` the kernel dumps the (s, z) it
// computes the scaling from, H = G'W^-2 G (+A'A for a sing problem) before the factorisation, and the right-hand side
// and result of the affine (`phase` = 1) or combined (2) solve_kkt -- the inputs and outputs of the reference's
// compute_scaling / setup_iter / solve_kkt (src/densesolver.jl:41-90), so that a test can feed the same inputs to the
// oracle.  Any output pointer may be null.  The problem's entries of the result arrays are overwritten.
int socp_b200_debug_fused_step(socp_handle* h, int64_t index, int32_t iter, int32_t phase, double* s, double* z,
                               double* H, double* dx, double* dy, double* dz, double* ds, double* cx, double* cy,
                               double* cz, double* cs) {
    if (!h) return SOCP_ERR_NULL;
    return guarded(h, [&]() {
        need(index >= 0 && index < h->batch, SOCP_ERR_SIZE, "problem index out of range");
        need(iter >= 0 && (phase == 1 || phase == 2), SOCP_ERR_SIZE, "iter >= 0 and phase 1 or 2");
        const int n = h->n, p = h->p, k = h->k;
        for (auto& sh : h->shards) {
            if (index < sh.first || index >= sh.first + sh.batch) continue;
            CK(cudaSetDevice(sh.device));
            need(sh.have_data, SOCP_ERR_STATE, "debug_fused_step called before set_data");
            need(fused_kind(h, sh) == FUSED_V3, SOCP_ERR_SIZE, "debug_fused_step needs a layout the fused_v3 kernel takes");
            const int len = f3_dbg_size(n, p, k);
            double* d_dbg = nullptr;
            CK(cudaMalloc((void**)&d_dbg, sizeof(double) * len));
            std::vector<double> buf((size_t)len, 0.0);
            int rc = 0;
            try {
                CK(cudaMemsetAsync(d_dbg, 0, sizeof(double) * len, sh.stream));
                socp_params prm;
                socp_b200_default_params(&prm);
                F3Glob g = f3_glob(sh);
                g.dbg = d_dbg;
                g.dbg_prob = (int)(index - sh.first);
                g.dbg_iter = iter;
                g.dbg_phase = phase;
                const LoopParams lp{prm.max_iter, prm.tol, prm.step_damp, prm.init_eps};
                solve_fused3_ext(sh.fused3, g, g.dbg_prob, 1, lp, sh.sing_known ? 0 : 1, 0, sh.stream,
                                 getenv("SOCP_B200_GENERIC_ONLY") == nullptr, 0);
                CK(cudaGetLastError());
                CK(cudaMemcpyAsync(buf.data(), d_dbg, sizeof(double) * len, cudaMemcpyDeviceToHost, sh.stream));
                CK(cudaStreamSynchronize(sh.stream));
            } catch (...) {
                cudaFree(d_dbg);
                throw;
            }
            (void)rc;
            cudaFree(d_dbg);
            const double* o = buf.data();
            auto take = [&](double* dst, int cnt) {
                if (dst) memcpy(dst, o, sizeof(double) * cnt);
                o += cnt;
            };
            take(s, k); take(z, k); take(H, n * n);
            take(dx, n); take(dy, p); take(dz, k); take(ds, k);
            take(cx, n); take(cy, p); take(cz, k); take(cs, k);
            return;
        }
    });
}

// Measurement utility (not a reference interface): runs the device kernels of one step-level call `reps` times on
// the data already resident after compute_scaling / setup_iter and returns the mean CUDA-event time per call.
//   0 compute_scaling   1 scale! (W v)   2 iscale! (W^-1 v)   3 vprod!   4 iprod!   5 compute_step
//   6 Gt = W^-1 G       7 SYRK H = Gt'Gt 8 Cholesky of H (H restored from a copy before every repetition)
//   9 L L' solve, one right-hand side    10 G'v gemv          11 G v gemv
int socp_b200_profile_step(socp_handle* h, int32_t which, int32_t reps, double* ms_per_call) {
    if (!h || !ms_per_call) return SOCP_ERR_NULL;
    return guarded(h, [&]() {
        need(h->shards.size() == 1, SOCP_ERR_STATE, "profile_step works on a single-device handle");
        need(reps > 0 && which >= 0 && which <= 11, SOCP_ERR_SIZE, "bad arguments");
        Shard& sh = h->shards[0];
        CK(cudaSetDevice(sh.device));
        need(sh.have_data && sh.have_scaling, SOCP_ERR_STATE, "profile_step needs set_data and compute_scaling first");
        Ws& w = sh.w;
        const int n = w.L.n, k = w.L.k, B = sh.batch;
        ensure_prepared(sh);
        ensure_tiled(sh);
        double* Hcopy = nullptr;
        const size_t hbytes = sizeof(double) * (size_t)B * w.ldh * n;
        if (which == 7 || which == 8 || which == 9) {
            launch_build_gt(sh, false, nullptr);
            syrk(sh, true, w.Gt, (int64_t)w.ldgt * n, w.ldgt, n, w.kpad, w.H, (int64_t)w.ldh * n, w.ldh, 1.0, 0.0, nullptr, 0, 0, nullptr, nullptr);
            if (which == 8) {
                CK(cudaMalloc((void**)&Hcopy, hbytes));
                CK(cudaMemcpyAsync(Hcopy, w.H, hbytes, cudaMemcpyDeviceToDevice, sh.stream));
            }
            if (which == 9) {
                CK(cudaMemsetAsync(w.fail, 0, sizeof(int) * B, sh.stream));
                potrf(sh, w.H, (int64_t)w.ldh * n, w.ldh, n, w.XH, w.fail, nullptr);
            }
        }
        double total = 0.0;
        for (int r = 0; r < reps + 1; ++r) {          // repetition 0 warms up
            if (which == 8) {
                CK(cudaMemcpyAsync(w.H, Hcopy, hbytes, cudaMemcpyDeviceToDevice, sh.stream));
                CK(cudaMemsetAsync(w.fail, 0, sizeof(int) * B, sh.stream));
            }
            CK(cudaEventRecord(sh.ev[0], sh.stream));
            switch (which) {
                case 0: launch_scaling(sh, nullptr); break;
                case 1: launch_apply(sh, 0, w.s, w.kt3); break;
                case 2: launch_apply(sh, 1, w.s, w.kt3); break;
                case 3: launch_vprod(sh, w.s, w.z, w.k0); break;
                case 4: launch_iprod(sh, w.lam, w.z, w.k0); break;
                case 5: launch_compute_step(sh, w.lam, w.s, w.z, w.k0); break;
                case 6: launch_build_gt(sh, false, nullptr); break;
                case 7: syrk(sh, true, w.Gt, (int64_t)w.ldgt * n, w.ldgt, n, w.kpad, w.H, (int64_t)w.ldh * n, w.ldh, 1.0, 0.0, nullptr, 0, 0, nullptr, nullptr); break;
                case 8: potrf(sh, w.H, (int64_t)w.ldh * n, w.ldh, n, w.XH, w.fail, nullptr); break;
                case 9: potrs(sh, w.H, (int64_t)w.ldh * n, w.ldh, n, w.XH, w.rx, n, n, 1, nullptr); break;
                case 10: gemv_t(sh, w.G, w.sG, k, k, n, w.z, k, w.dx, n, 1.0, epi(), nullptr); break;
                default: gemv_n(sh, w.G, w.sG, k, k, n, w.x, n, w.dz, k, 1.0, epi(), nullptr); break;
            }
            CK(cudaEventRecord(sh.ev[1], sh.stream));
            CK(cudaStreamSynchronize(sh.stream));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, sh.ev[0], sh.ev[1]));
            if (r > 0) total += ms;
        }
        if (Hcopy) cudaFree(Hcopy);
        CK(cudaGetLastError());
        *ms_per_call = total / reps;
        sh.have_factor = false;
    });
}


#ifdef SOCP_PHASE_TIMING
// profiling build only (not declared in include/socp_b200.h): the 2 x 16 phase-cycle counters of shard 0
// (fused_v2 | fused_v3), see tools/phase_timing.py
int socp_b200_debug_phase_clocks(socp_handle* h, unsigned long long* out32, int reset) {
    if (!h) return SOCP_ERR_NULL;
    return guarded(h, [&]() {
        Shard& sh = h->shards[0];
        CK(cudaSetDevice(sh.device));
        CK(cudaStreamSynchronize(sh.stream));
        if (out32) CK(cudaMemcpy(out32, sh.fused2.d_clk, sizeof(unsigned long long) * 32, cudaMemcpyDeviceToHost));
        if (reset) CK(cudaMemset(sh.fused2.d_clk, 0, sizeof(unsigned long long) * 32));
    });
}
#endif

}  // extern "C"
