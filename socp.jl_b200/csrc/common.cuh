// common.cuh -- shared definitions for the socp_b200 CUDA sources (sm_100a).
#pragma once
#if defined(__CUDACC_RTC__)
// run-time compilation of a layout-specialised kernel (lane_jit.cu): NVRTC has the CUDA built-ins but no host headers
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned char uint8_t;
typedef unsigned long long uintptr_t;
#ifndef INFINITY
#define INFINITY __longlong_as_double(0x7ff0000000000000LL)
#endif
#else
#ifdef SOCP_SIMT_EMU
// host build of the kernels for CPU-side tests (tests/simt_emu/, test infrastructure -- never part of libsocp_b200)
#include "simt_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <math.h>
#endif

namespace socp {

constexpr unsigned FULL_MASK = 0xffffffffu;
constexpr int KIND_POC = 0;
constexpr int KIND_SOC = 1;

// per-problem status words (mirror include/socp_b200.h)
constexpr int ST_CONVERGED = 0;
constexpr int ST_MAXITER = 1;
constexpr int ST_NUMERICAL = 2;
constexpr int ST_RUNNING = -1;

// Cone layout as the kernels see it ("work cones": POC blocks are split into
// chunks so that a warp never owns more than POC_CHUNK elementwise entries; SOC
// blocks are kept whole).  Arrays live in global memory, shared by the batch.
struct ConeLayout {
    int n, p, k;
    int ncones;            // work cones
    int deg;               // reference deg(cones): d per POC block, 1 per SOC (src/vectors.jl:165-179)
    const int* kind;
    const int* offs;
    const int* dim;
};

// Per-problem scalar block kept in global memory by the tiled path.
struct ProbScalars {
    double resid;      // ||rx|| + ||ry|| + z's        (src/solver.jl:122)
    double gap;        // z's
    double ll;         // lambda'lambda
    double t;          // affine compute_step          (src/solver.jl:130)
    double sigma;
    double mu;
    double step;
    double pobj;
    double dobj;
};

// global arrays of one device shard (batch slowest), as every kernel sees them
struct Ws {
    ConeLayout L;
    int ldgt, kpad;        // Gt: kpad x n, column stride ldgt (= kpad), pad rows zero
    int ldh;               // H: n x n, column stride ldh
    int ldap, ppad;        // Ap: padded copy of A (ppad x n)
    int ldm;               // M: p x p, column stride ldm
    // problem data (strides 0 when shared across the batch)
    const double *c, *A, *b, *G, *h;
    int64_t sA, sG;
    const uint8_t* sing;
    // iterate
    double *x, *y, *z, *s;
    // scaling
    double *lam, *wb, *iwb, *eta;   // eta: 4 scalars per cone (eta, 1/eta, 1/eta^2, 1/(1+wbar0)), [4][ncones]
    // right-hand side / direction
    double *dx, *dy, *dz, *ds;
    double *rx, *ry, *rz, *rs;
    // solve_kkt temporaries (k-vectors)
    double *k0, *k2, *u;
    double *kt2, *kt3;
    // factor workspaces
    double *Gt, *H, *HiAt, *M, *AA, *Ap;
    double *XH, *XM;       // [batch][ceil(n/64)][64*64] inverted diagonal blocks of the factors of H and M
    ProbScalars* sc;
    double *pobj, *dobj;   // [batch] objectives of the returned iterate
    int *status, *iters, *active, *fail;
    int* nactive;          // [max_iter+2] counters
};

// solver constants of one solve (defaults: reference src/solver.jl:105,122,146,91)
struct LoopParams {
    int max_iter;
    double tol, step_damp, init_eps;
};

// ---------------------------------------------------------------------------
// Fast FP64 reciprocal / rsqrt / sqrt: MUFU seed (rcp.approx / rsqrt.approx, ~20
// bits) + two Newton steps, accurate to ~1-2 ulp.  The IEEE div/sqrt sequences
// cost hundreds of cycles of dependent latency and sit on every serial chain of
// the interior-point iteration; the differences are at the 1e-16 level, far
// inside the parity tolerances (1e-10 per step).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double fast_rcp(double a) {
#ifdef SOCP_SIMT_EMU
    return 1.0 / a;
#endif
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    double e = fma(-a, x, 1.0);
    x = fma(x, e, x);
    e = fma(-a, x, 1.0);
    x = fma(x, e, x);
    return x;
}
__device__ __forceinline__ double fast_rsqrt(double a) {
#ifdef SOCP_SIMT_EMU
    return 1.0 / sqrt(a);
#endif
    double x;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    const double h = 0.5 * a;
    double e = fma(-h * x, x, 0.5);
    x = fma(x, e, x);
    e = fma(-h * x, x, 0.5);
    x = fma(x, e, x);
    return x;
}
// sqrt(a) for a >= 0 (0 -> 0); negative / NaN inputs give NaN like the IEEE one.
__device__ __forceinline__ double fast_sqrt(double a) {
    const double r = fast_rsqrt(a);
    double y = a * r;
    const double d = fma(-y, y, a);
    y = fma(0.5 * r, d, y);
    return (a == 0.0) ? 0.0 : y;
}

// Batch index of a CTA of the batch-wide tiled kernels: grids are (tiles, by, bz) with the batch folded over y and z
// (gridDim.y is capped at 65535); the kernel guards b < nbatch.  Host side: batch_grid() in solver.cu.
__device__ __forceinline__ int batch_index() { return (int)(blockIdx.z * gridDim.y + blockIdx.y); }

// compute_step, reference src/mats.jl:30-40: t = max(scmax(lam, ds), scmax(lam, dz), 0); step = t == 0 ? 1 : min(1, 1/t)
__device__ __forceinline__ double step_from_t(double t) {
    t = fmax(t, 0.0);
    return (t == 0.0) ? 1.0 : fmin(1.0, fast_rcp(t));
}

// D(8x8) += A(8x4, row) * B(4x8, col) on the FP64 tensor pipe (SASS DMMA.8x8x4).  Lane l holds A[l>>2][l&3],
// B[l&3][l>>2] and C[l>>2][2(l&3) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
#ifdef SOCP_SIMT_EMU
    simt_emu::dmma884(c0, c1, a, b);
#else
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
#endif
}
// named barrier `ID` (1..15) over NT threads: arrive-only for a warp that need not wait (it must not touch what the
// waiting warps go on to write before it has synchronised with them again)
template <int ID, int NT>
__device__ __forceinline__ void named_bar_sync() {
#ifdef SOCP_SIMT_EMU
    emu_bar_sync(ID, NT);
#else
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(NT) : "memory");
#endif
}
template <int ID, int NT>
__device__ __forceinline__ void named_bar_arrive() {
#ifdef SOCP_SIMT_EMU
    emu_bar_arrive(ID, NT);
#else
    asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(NT) : "memory");
#endif
}

// the same with the barrier number in a register (teams of one CTA each own a pair of barriers)
template <int NT>
__device__ __forceinline__ void bar_sync_id(int id) {
#ifdef SOCP_SIMT_EMU
    emu_bar_sync(id, NT);
#else
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NT) : "memory");
#endif
}
template <int NT>
__device__ __forceinline__ void bar_arrive_id(int id) {
#ifdef SOCP_SIMT_EMU
    emu_bar_arrive(id, NT);
#else
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(NT) : "memory");
#endif
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ int warp_or(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// CTA-wide reductions through a small shared scratch (>= 32 doubles).  Result is
// broadcast to every thread.  Deterministic (fixed tree), no atomics.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = (lane < nw) ? scratch[lane] : 0.0;
    r = warp_sum(r);
    return r;
}
// four sums at once (scratch >= 4*32 doubles): one barrier pair instead of four
__device__ __forceinline__ void block_sum4(double& a, double& b, double& c, double& d, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c); d = warp_sum(d);
    __syncthreads();
    if (lane == 0) { scratch[warp] = a; scratch[32 + warp] = b; scratch[64 + warp] = c; scratch[96 + warp] = d; }
    __syncthreads();
    a = warp_sum((lane < nw) ? scratch[lane] : 0.0);
    b = warp_sum((lane < nw) ? scratch[32 + lane] : 0.0);
    c = warp_sum((lane < nw) ? scratch[64 + lane] : 0.0);
    d = warp_sum((lane < nw) ? scratch[96 + lane] : 0.0);
}
__device__ __forceinline__ double block_max(double v, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = (lane < nw) ? scratch[lane] : -INFINITY;
    r = warp_max(r);
    return r;
}
__device__ __forceinline__ int block_or(int v, int* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_or(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    int r = (lane < nw) ? scratch[lane] : 0;
    r = warp_or(r);
    return r;
}

}  // namespace socp
