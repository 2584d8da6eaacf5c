// microbench.cu -- B200 latency/throughput probes for the FP64 building blocks of the fused kernel.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../socp.jl_b200/csrc/common.cuh"
#include "../socp.jl_b200/csrc/linalg.cuh"
using namespace socp;

__global__ void k_lat(double* out, long long* clk, int iters) {
    const int lane = threadIdx.x & 31;
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double a = 1.0 + 1e-9 * lane, b = 1.0000001, c = 1e-9;
    long long t0, t1;
    // 0: dependent DFMA chain
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
    t1 = clock64();
    if (threadIdx.x == 0) clk[0] = (t1 - t0);
    // 1: 4 independent DFMA chains (throughput per warp)
    double a1 = a, a2 = a + 1, a3 = a + 2, a4 = a + 3;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); a4 = fma(a4, b, c); }
    t1 = clock64();
    if (threadIdx.x == 0) clk[1] = (t1 - t0);
    a = a1 + a2 + a3 + a4;
    // 2: dependent shuffle chain
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = __shfl_xor_sync(FULL_MASK, a, 1); a = __shfl_xor_sync(FULL_MASK, a, 2); a = __shfl_xor_sync(FULL_MASK, a, 4); a = __shfl_xor_sync(FULL_MASK, a, 8); }
    t1 = clock64();
    if (threadIdx.x == 0) clk[2] = (t1 - t0);
    // 3: dependent LDS chain (pointer chasing through doubles)
    int idx = lane;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { idx = (int)sm[idx & 1023] + lane; idx = (int)sm[idx & 1023] + lane; idx = (int)sm[idx & 1023] + lane; idx = (int)sm[idx & 1023] + lane; }
    t1 = clock64();
    if (threadIdx.x == 0) clk[3] = (t1 - t0);
    a += idx;
    // 4: fast_rcp dependent chain
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = fast_rcp(a); a = fast_rcp(a + 1.0); a = fast_rcp(a + 1.0); a = fast_rcp(a + 1.0); }
    t1 = clock64();
    if (threadIdx.x == 0) clk[4] = (t1 - t0);
    // 5: IEEE division dependent chain
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = 1.0 / (a + 1.0); a = 1.0 / (a + 1.0); a = 1.0 / (a + 1.0); a = 1.0 / (a + 1.0); }
    t1 = clock64();
    if (threadIdx.x == 0) clk[5] = (t1 - t0);
    // 6: IEEE sqrt dependent chain
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = sqrt(a + 1.0); a = sqrt(a + 1.0); a = sqrt(a + 1.0); a = sqrt(a + 1.0); }
    t1 = clock64();
    if (threadIdx.x == 0) clk[6] = (t1 - t0);
    // 7: dependent DMMA chain
    double c0 = 0.0, c1 = 0.0;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { dmma884(c0, c1, a, b); dmma884(c0, c1, a, b); dmma884(c0, c1, a, b); dmma884(c0, c1, a, b); }
    t1 = clock64();
    if (threadIdx.x == 0) clk[7] = (t1 - t0);
    // 8: 4 independent DMMA chains
    double d0 = 0, d1 = 0, e0 = 0, e1 = 0, f0 = 0, f1 = 0;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { dmma884(c0, c1, a, b); dmma884(d0, d1, a, b); dmma884(e0, e1, a, b); dmma884(f0, f1, a, b); }
    t1 = clock64();
    if (threadIdx.x == 0) clk[8] = (t1 - t0);
    // 9: __syncthreads chain
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { __syncthreads(); __syncthreads(); __syncthreads(); __syncthreads(); }
    t1 = clock64();
    if (threadIdx.x == 0) clk[9] = (t1 - t0);
    // 10: fast_rsqrt chain
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = fast_rsqrt(a + 1.0); a = fast_rsqrt(a + 1.0); a = fast_rsqrt(a + 1.0); a = fast_rsqrt(a + 1.0); }
    t1 = clock64();
    if (threadIdx.x == 0) clk[10] = (t1 - t0);
    // 11: STS -> barrier -> LDS round trip
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        sm[threadIdx.x] = a; __syncthreads(); a = sm[(threadIdx.x + 32) % blockDim.x] + 1.0;
        sm[threadIdx.x] = a; __syncthreads(); a = sm[(threadIdx.x + 32) % blockDim.x] + 1.0;
        sm[threadIdx.x] = a; __syncthreads(); a = sm[(threadIdx.x + 32) % blockDim.x] + 1.0;
        sm[threadIdx.x] = a; __syncthreads(); a = sm[(threadIdx.x + 32) % blockDim.x] + 1.0;
    }
    t1 = clock64();
    if (threadIdx.x == 0) clk[11] = (t1 - t0);
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + c0 + c1 + d0 + d1 + e0 + e1 + f0 + f1;
}

// throughput: many warps doing independent DFMA / DMMA
__global__ void k_tput(double* out, int iters, int mode) {
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0000001, c = 1e-9;
    double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
    if (mode == 0) {
        for (int i = 0; i < iters; ++i) {
            x0 = fma(x0, b, c); x1 = fma(x1, b, c); x2 = fma(x2, b, c); x3 = fma(x3, b, c);
            x4 = fma(x4, b, c); x5 = fma(x5, b, c); x6 = fma(x6, b, c); x7 = fma(x7, b, c);
        }
    } else {
        for (int i = 0; i < iters; ++i) {
            dmma884(x0, x1, a, b); dmma884(x2, x3, a, b); dmma884(x4, x5, a, b); dmma884(x6, x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main() {
    double* out; long long* clk;
    cudaMalloc(&out, 8 * 148 * 1024 * 16); cudaMalloc(&clk, 16 * 8);
    const int iters = 2000;
    const char* names[] = {"DFMA dependent", "DFMA 4 indep (per 4)", "SHFL dependent", "LDS dependent (+cvt)", "fast_rcp dependent",
                           "IEEE div dependent", "IEEE sqrt dependent", "DMMA dependent", "DMMA 4 indep (per 4)", "__syncthreads",
                           "fast_rsqrt dependent", "STS+barrier+LDS+DADD"};
    for (int threads : {32, 256}) {
        k_lat<<<1, threads>>>(out, clk, iters);
        cudaDeviceSynchronize();
        long long h[16];
        cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
        printf("--- 1 CTA of %d threads: cycles per op (chains of %d)\n", threads, 4 * iters);
        for (int i = 0; i < 12; ++i) printf("  %-26s %8.1f\n", names[i], (double)h[i] / (4.0 * iters));
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        const int it = 20000;
        k_tput<<<148 * 4, 512>>>(out, it, mode);
        cudaEventRecord(e0);
        k_tput<<<148 * 4, 512>>>(out, it, mode);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)148 * 4 * 512 * it * (mode == 0 ? 8.0 * 2 : 4.0 * 512 / 32.0);   // flops
        // DMMA: per warp-instr 8*8*4*2 = 512 flop; per thread-iteration 4 instr -> 4*512/32 flop per thread
        printf("throughput %s: %.2f TFLOP/s (%.3f ms)\n", mode == 0 ? "DFMA" : "DMMA m8n8k4", ops / (ms * 1e-3) / 1e12, ms);
    }
    printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
