// common.cuh -- shared definitions for the socp_b200 CUDA sources (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

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
