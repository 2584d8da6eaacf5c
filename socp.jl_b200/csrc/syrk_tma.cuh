// syrk_tma.cuh -- entry points of the TMA-fed Gram product (syrk_tma.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace socp {

// A is K x N column-major per problem (ld = lda, problem stride strideA): can the tensor-map version take it?
bool syrk_tma_supported(const double* A, int64_t strideA, int lda, int N, int K);
// C = beta C + alpha A'A (+ addC where addFlag) on the lower-triangle 128 x 128 tiles; grid = (tiles, by, bz) as built
// by batch_grid() for ntiles(ntiles+1)/2 tiles.  Returns the launch status.
cudaError_t syrk_tma_launch(cudaStream_t stream, dim3 grid, const double* A, int64_t strideA, int lda, int N, int K,
                            double* C, int64_t strideC, int ldc, double alpha, double beta, const double* addC,
                            int64_t strideAdd, int ldadd, const uint8_t* addFlag, const int* active, int nbatch);

}  // namespace socp
