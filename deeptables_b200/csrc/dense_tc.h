// Internal interface of the tensor-core dense GEMMs (dense_tc.cu), shared by the Dense entry points
// (dense_bn_loss.cu) and the any-shape CIN formulation (cin_fp32.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace dtb {

// bytes of packed bf16 hi/lo weight images for a [K, Nout] operand (the caller's workspace holds them)
size_t dense_tc_pack_bytes(int K, int Nout);

// out[M, Nout] = act(A[M, K] . op(W) + bias).  transposed = 0: op(W)[k, n] = W[k*ldw + n];  1: op(W)[k, n] = W[n*ldw + k].
int dense_tc_rows(const float* A, int lda, const float* W, int ldw, int transposed, const float* bias, float* out,
                  int ldo, int M, int K, int Nout, int act, void* workspace, size_t workspace_bytes, cudaStream_t st);

// dW[K, N] += X[M, K]^T dZ[M, N];  dbias[N] += column sums of dZ (when not null)
int dense_tc_wgrad(const float* X, int ldx, const float* dZ, int ldz, float* dW, int ldw, float* dbias, int M, int K,
                   int N, cudaStream_t st);

}  // namespace dtb
