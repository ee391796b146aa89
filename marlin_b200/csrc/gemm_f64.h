#pragma once
#include <cuda_runtime.h>

namespace mb {

// C (M x N, col-major, ldc) = alpha * op(A) * op(B) + beta * C, fp64.
// op(X) = X or X^T; A is stored (M x K) if !transA else (K x M); B likewise (K x N) / (N x K).
// Dispatches the TMA + DMMA tensor-core kernel when pointers are 16B aligned and leading
// dimensions are even, otherwise (or if force_generic) the CUDA-core kernel.
cudaError_t gemm_f64(bool transA, bool transB, int M, int N, int K, double alpha, const double* A, long long lda,
                     const double* B, long long ldb, double beta, double* C, long long ldc, int num_sms,
                     cudaStream_t stream, bool force_generic, int* launches);

bool gemm_f64_tma_eligible(const double* A, long long lda, const double* B, long long ldb);

}  // namespace mb
