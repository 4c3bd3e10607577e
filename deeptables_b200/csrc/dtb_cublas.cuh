// cuBLAS context shared by the plain-GEMM call sites (Dense layers, fp32 CIN formulation).
#pragma once
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include "dtb_common.cuh"

namespace dtb {

cublasHandle_t cublas_handle(cudaStream_t stream);

#define DTB_CUBLAS_OK(expr)                                                              \
  do {                                                                                   \
    cublasStatus_t _s = (expr);                                                          \
    if (_s != CUBLAS_STATUS_SUCCESS) {                                                   \
      dtb::set_error("%s: cuBLAS error %d at %s:%d", __func__, (int)_s, __FILE__, __LINE__); \
      return DTB_ERR_CUBLAS;                                                             \
    }                                                                                    \
  } while (0)

// Row-major helpers.  All matrices row-major with the given leading dimensions.
// C[m,n] (+)= A[m,k] @ B[k,n]
static inline cublasStatus_t gemm_nn(cublasHandle_t h, int m, int n, int k, const float* A, int lda,
                                     const float* B, int ldb, float* C, int ldc, float beta) {
  const float one = 1.f;
  return cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, m, k, &one, B, ldb, A, lda, &beta, C, ldc);
}
// C[m,n] (+)= A[m,k] @ B[n,k]^T
static inline cublasStatus_t gemm_nt(cublasHandle_t h, int m, int n, int k, const float* A, int lda,
                                     const float* B, int ldb, float* C, int ldc, float beta) {
  const float one = 1.f;
  return cublasSgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, n, m, k, &one, B, ldb, A, lda, &beta, C, ldc);
}
// C[m,n] (+)= A[k,m]^T @ B[k,n]
static inline cublasStatus_t gemm_tn(cublasHandle_t h, int m, int n, int k, const float* A, int lda,
                                     const float* B, int ldb, float* C, int ldc, float beta) {
  const float one = 1.f;
  return cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, m, k, &one, B, ldb, A, lda, &beta, C, ldc);
}

}  // namespace dtb
