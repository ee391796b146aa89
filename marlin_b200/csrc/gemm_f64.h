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

// One persistent launch for a blocked multiply: C[c] = sum_kk A[i*k+kk] * B[kk*n+j] for every c = i*n+j in my_c
// (all 'N' column-major blocks).  Returns cudaErrorNotSupported when the group does not fit the fixed-size parameter
// block or an operand is not TMA-addressable; the caller then falls back to one launch per block product.
cudaError_t gemm_f64_grouped(int m, int k, int n, const int* my_c, int num_c, const double* const* A, const long long* lda,
                             const double* const* B, const long long* ldb, double* const* C, const long long* ldc,
                             const int* row_len, const int* k_len, const int* col_len, int num_sms, cudaStream_t stream,
                             int* launches);

bool gemm_f64_tma_eligible(const double* A, long long lda, const double* B, long long ldb);

}  // namespace mb
