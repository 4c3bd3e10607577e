// Cross network (Cross.call, layers.py:428-436) on the batch-normalised concat vector:
//   x_{l+1} = x0 * (x_l . w_l) + x_l + b_l           kernels/biases [n_layers, W]
// HBM-bound: one warp owns one batch row; x0 and the running x_l live in shared memory (W floats
// each), the W-long dot product is a strided register sum + 5 shuffles.  Per row the kernel moves
// 2*W*4 bytes (read x, write y) for 4*W*n_layers FLOP.
// Backward recomputes x_1..x_{L-1} from the saved per-layer scalars (x_l . w_l) -- no reductions --
// and accumulates d_kernels / d_biases per CTA in shared memory before one atomic per element.
#include "dtb_common.cuh"

namespace dtb {

constexpr int kCrossWarps = 4;

__global__ void __launch_bounds__(kCrossWarps * 32)
cross_fwd_kernel(const float* __restrict__ X, const float* __restrict__ kernels,
                 const float* __restrict__ biases, float* __restrict__ Y, float* __restrict__ xw_saved, int B,
                 int W, int n_layers) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* x0 = smem + (size_t)wib * 2 * W;
  float* xl = x0 + W;
  const int warp = blockIdx.x * kCrossWarps + wib;
  const int n_warps = gridDim.x * kCrossWarps;
  for (int row = warp; row < B; row += n_warps) {
    for (int c = lane; c < W; c += 32) {
      const float v = X[(int64_t)row * W + c];
      x0[c] = v;
      xl[c] = v;
    }
    __syncwarp();
    for (int l = 0; l < n_layers; ++l) {
      const float* w = kernels + (size_t)l * W;
      const float* b = biases + (size_t)l * W;
      float dot = 0.f;
      for (int c = lane; c < W; c += 32) dot += xl[c] * __ldg(w + c);
      dot = warp_sum(dot);
      if (xw_saved && lane == 0) xw_saved[(int64_t)row * n_layers + l] = dot;
      for (int c = lane; c < W; c += 32) xl[c] = x0[c] * dot + xl[c] + __ldg(b + c);
      __syncwarp();
    }
    for (int c = lane; c < W; c += 32) Y[(int64_t)row * W + c] = xl[c];
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kCrossWarps * 32)
cross_bwd_kernel(const float* __restrict__ X, const float* __restrict__ kernels,
                 const float* __restrict__ biases, const float* __restrict__ xw_saved,
                 const float* __restrict__ dY, float* __restrict__ dX, float* __restrict__ d_kernels,
                 float* __restrict__ d_biases, int B, int W, int n_layers) {
  extern __shared__ float smem[];
  // layout: [2*n_layers*W] CTA accumulators (dW, db) | per warp: x_0..x_{L-1} [L*W], g [W], gx0 [W]
  float* acc_w = smem;
  float* acc_b = smem + (size_t)n_layers * W;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* xs = smem + (size_t)2 * n_layers * W + (size_t)wib * (n_layers + 2) * W;
  float* g = xs + (size_t)n_layers * W;
  float* dx0 = g + W;
  for (int i = threadIdx.x; i < 2 * n_layers * W; i += blockDim.x) smem[i] = 0.f;
  __syncthreads();
  const int warp = blockIdx.x * kCrossWarps + wib;
  const int n_warps = gridDim.x * kCrossWarps;
  for (int row = warp; row < B; row += n_warps) {
    const float* s = xw_saved + (int64_t)row * n_layers;
    for (int c = lane; c < W; c += 32) {
      float x0 = X[(int64_t)row * W + c];
      float xl = x0;
      xs[c] = x0;
      for (int l = 0; l + 1 < n_layers; ++l) {
        xl = x0 * __ldg(s + l) + xl + __ldg(biases + (size_t)l * W + c);
        xs[(size_t)(l + 1) * W + c] = xl;
      }
      g[c] = dY[(int64_t)row * W + c];
      dx0[c] = 0.f;
    }
    __syncwarp();
    for (int l = n_layers - 1; l >= 0; --l) {
      const float* xl = xs + (size_t)l * W;
      float gx0 = 0.f;
      for (int c = lane; c < W; c += 32) gx0 += g[c] * xs[c];
      gx0 = warp_sum(gx0);
      const float sl = __ldg(s + l);
      for (int c = lane; c < W; c += 32) {
        const float gc = g[c];
        atomicAdd(&acc_b[(size_t)l * W + c], gc);
        atomicAdd(&acc_w[(size_t)l * W + c], xl[c] * gx0);
        dx0[c] += gc * sl;
        g[c] = gc + __ldg(kernels + (size_t)l * W + c) * gx0;
      }
      __syncwarp();
    }
    for (int c = lane; c < W; c += 32) dX[(int64_t)row * W + c] = g[c] + dx0[c];
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_layers * W; i += blockDim.x) {
    if (acc_w[i] != 0.f) atomicAdd(d_kernels + i, acc_w[i]);
    if (acc_b[i] != 0.f) atomicAdd(d_biases + i, acc_b[i]);
  }
}

}  // namespace dtb

using namespace dtb;

extern "C" {

int dtb_cross_fwd(const float* X, const float* kernels, const float* biases, float* Y, float* xw_saved, int B,
                  int W, int n_layers, void* stream) {
  DTB_CHECK_ARG(X && kernels && biases && Y, "NULL argument");
  DTB_CHECK_ARG(B >= 0 && W > 0 && n_layers >= 0, "bad shape");
  if (B == 0) return DTB_OK;
  const size_t smem = (size_t)kCrossWarps * 2 * W * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("dtb_cross_fwd: input width %d too large for the shared-memory row buffers", W);
    return DTB_ERR_UNSUPPORTED;
  }
  DTB_CUDA_OK(cudaFuncSetAttribute(cross_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int blocks = ceil_div(B, kCrossWarps);
  const int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  cross_fwd_kernel<<<blocks, kCrossWarps * 32, smem, (cudaStream_t)stream>>>(X, kernels, biases, Y, xw_saved, B,
                                                                             W, n_layers);
  DTB_LAUNCH_OK();
  return DTB_OK;
}

int dtb_cross_bwd(const float* X, const float* kernels, const float* biases, const float* xw_saved,
                  const float* dY, float* dX, float* d_kernels, float* d_biases, int B, int W, int n_layers,
                  void* stream) {
  DTB_CHECK_ARG(X && kernels && biases && xw_saved && dY && dX && d_kernels && d_biases, "NULL argument");
  DTB_CHECK_ARG(B >= 0 && W > 0 && n_layers > 0, "bad shape");
  if (B == 0) return DTB_OK;
  const size_t smem = ((size_t)2 * n_layers * W + (size_t)kCrossWarps * (n_layers + 2) * W) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("dtb_cross_bwd: width %d x %d layers exceeds the shared-memory budget", W, n_layers);
    return DTB_ERR_UNSUPPORTED;
  }
  DTB_CUDA_OK(cudaFuncSetAttribute(cross_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int blocks = ceil_div(B, kCrossWarps * 8);
  const int cap = sm_count() * 2;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cross_bwd_kernel<<<blocks, kCrossWarps * 32, smem, (cudaStream_t)stream>>>(X, kernels, biases, xw_saved, dY,
                                                                             dX, d_kernels, d_biases, B, W,
                                                                             n_layers);
  DTB_LAUNCH_OK();
  return DTB_OK;
}

}  // extern "C"
