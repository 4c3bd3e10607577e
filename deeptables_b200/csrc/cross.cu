// Cross network (Cross.call, layers.py:428-436) on the batch-normalised concat vector:
//   x_{l+1} = x0 * (x_l . w_l) + x_l + b_l           kernels/biases [n_layers, W]
// HBM-bound: one warp owns one batch row; x0 and the running x_l live in shared memory (W floats
// each), the W-long dot product is a strided register sum + 5 shuffles.  Per row the kernel moves
// 2*W*4 bytes (read x, write y) for 4*W*n_layers FLOP.
// Backward: see the register-resident variants below (per-row kernel emits dX + two [B, L] scalar tables, the
// weight gradients are column reductions); the shared-memory kernels serve inputs wider than 1024 columns.
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
  const int wpb = blockDim.x >> 5;
  const int warp = blockIdx.x * wpb + wib;
  const int n_warps = gridDim.x * wpb;
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

// Shared-memory twin of cross_bwd_reg_kernel below for inputs wider than the register variants hold: a warp keeps
// x0 / g / dx0 of its row in shared memory (3 W floats) and emits dX plus the two [B, L] scalar tables; the weight
// gradients come from the same column reductions (cross_colreduce / gsum / combine).
__global__ void cross_bwd_smem_kernel(const float* __restrict__ X, const float* __restrict__ kernels,
                                      const float* __restrict__ xw_saved, const float* __restrict__ dY,
                                      float* __restrict__ dX, float* __restrict__ coef, float* __restrict__ gx, int B,
                                      int W, int n_layers) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* x0 = smem + (size_t)wib * 3 * W;
  float* g = x0 + W;
  float* dx0 = g + W;
  const int warp = blockIdx.x * wpb + wib;
  const int n_warps = gridDim.x * wpb;
  for (int row = warp; row < B; row += n_warps) {
    const float* s = xw_saved + (int64_t)row * n_layers;
    for (int c = lane; c < W; c += 32) {
      x0[c] = X[(int64_t)row * W + c];
      g[c] = dY[(int64_t)row * W + c];
      dx0[c] = 0.f;
    }
    __syncwarp();
    float a_l = 1.f;                       // A_l = 1 + sum_{l' < l} s_l'
    for (int l = 0; l + 1 < n_layers; ++l) a_l += __ldg(s + l);
    for (int l = n_layers - 1; l >= 0; --l) {
      float gx0 = 0.f;
      for (int c = lane; c < W; c += 32) gx0 = fmaf(g[c], x0[c], gx0);
      gx0 = warp_sum(gx0);
      const float sl = __ldg(s + l);
      if (lane == 0) {
        coef[(int64_t)row * n_layers + l] = a_l * gx0;
        gx[(int64_t)row * n_layers + l] = gx0;
      }
      const float* w = kernels + (size_t)l * W;
      for (int c = lane; c < W; c += 32) {
        const float gc = g[c];
        dx0[c] = fmaf(gc, sl, dx0[c]);
        g[c] = fmaf(__ldg(w + c), gx0, gc);
      }
      if (l > 0) a_l -= __ldg(s + l - 1);
      __syncwarp();
    }
    for (int c = lane; c < W; c += 32) dX[(int64_t)row * W + c] = g[c] + dx0[c];
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// Register-resident variants (W <= 32*PER): a lane keeps its PER columns of x0 / x_l / g in registers,
// so a row costs one coalesced read and one coalesced write and the layers run out of registers.
//
// Backward without per-row weight-gradient traffic.  With s_l = x_l . w_l (saved by forward),
//   x_l = x0 * A_l + Bsum_l,   A_l = 1 + sum_{l'<l} s_l',  Bsum_l = sum_{l'<l} b_l'
//   gx_l = g_{l+1} . x0        (g_{l+1} = dLoss/dx_{l+1}),   g_l = g_{l+1} + w_l * gx_l
// so   dW_l[c] = sum_rows x_l[c] gx_l = sum_rows X[row,c] * (A_l gx_l)(row)  +  Bsum_l[c] * sum_rows gx_l
//      db_l[c] = sum_rows g_{l+1}[c]  = colsum(dY)[c] + sum_{l'>l} w_l'[c] * sum_rows gx_l'
// i.e. the per-row kernel only emits dX and two [B, L] scalar tables; the weight gradients are one
// column reduction (X^T coef, colsum dY) plus a W x L combine.
// ------------------------------------------------------------------------------------------
constexpr int kCrossMaxL = 8;

template <int PER>
__global__ void __launch_bounds__(256) cross_fwd_reg_kernel(const float* __restrict__ X, const float* __restrict__ kernels,
                                                             const float* __restrict__ biases, float* __restrict__ Y,
                                                             float* __restrict__ xw_saved, int B, int W, int n_layers) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int row = warp; row < B; row += n_warps) {
    float x0[PER], xl[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int c = lane + 32 * k;
      x0[k] = c < W ? X[(int64_t)row * W + c] : 0.f;
      xl[k] = x0[k];
    }
    for (int l = 0; l < n_layers; ++l) {
      const float* w = kernels + (size_t)l * W;
      const float* b = biases + (size_t)l * W;
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int c = lane + 32 * k;
        if (c < W) dot = fmaf(xl[k], __ldg(w + c), dot);
      }
      dot = warp_sum(dot);
      if (xw_saved && lane == 0) xw_saved[(int64_t)row * n_layers + l] = dot;
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int c = lane + 32 * k;
        if (c < W) xl[k] = fmaf(x0[k], dot, xl[k]) + __ldg(b + c);
      }
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int c = lane + 32 * k;
      if (c < W) Y[(int64_t)row * W + c] = xl[k];
    }
  }
}

template <int PER>
__global__ void __launch_bounds__(256) cross_bwd_reg_kernel(const float* __restrict__ X, const float* __restrict__ kernels,
                                                             const float* __restrict__ xw_saved, const float* __restrict__ dY,
                                                             float* __restrict__ dX, float* __restrict__ coef,
                                                             float* __restrict__ gx, int B, int W, int n_layers) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int row = warp; row < B; row += n_warps) {
    float x0[PER], g[PER], dx0[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int c = lane + 32 * k;
      x0[k] = c < W ? X[(int64_t)row * W + c] : 0.f;
      g[k] = c < W ? dY[(int64_t)row * W + c] : 0.f;
      dx0[k] = 0.f;
    }
    float sl[kCrossMaxL], al[kCrossMaxL];
    float run = 1.f;
#pragma unroll
    for (int l = 0; l < kCrossMaxL; ++l) {
      sl[l] = l < n_layers ? __ldg(xw_saved + (int64_t)row * n_layers + l) : 0.f;
      al[l] = run;
      run += sl[l];
    }
#pragma unroll
    for (int l = kCrossMaxL - 1; l >= 0; --l) {
      if (l < n_layers) {
        float gx0 = 0.f;
#pragma unroll
        for (int k = 0; k < PER; ++k) gx0 = fmaf(g[k], x0[k], gx0);
        gx0 = warp_sum(gx0);
        if (lane == 0) {
          coef[(int64_t)row * n_layers + l] = al[l] * gx0;
          gx[(int64_t)row * n_layers + l] = gx0;
        }
        const float* w = kernels + (size_t)l * W;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
          const int c = lane + 32 * k;
          dx0[k] = fmaf(g[k], sl[l], dx0[k]);
          if (c < W) g[k] = fmaf(__ldg(w + c), gx0, g[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int c = lane + 32 * k;
      if (c < W) dX[(int64_t)row * W + c] = g[k] + dx0[k];
    }
  }
}

// m1[c, l] += sum_rows X[row,c] coef[row,l] ; s[c] += sum_rows dY[row,c]      block = 32 columns x 8 row lanes
__global__ void cross_colreduce_kernel(const float* __restrict__ X, const float* __restrict__ dY,
                                       const float* __restrict__ coef, float* __restrict__ m1, float* __restrict__ s,
                                       int B, int W, int n_layers, int rows_per_block, int ld, int want_s) {
  // coef / m1 are [*, ld] with this launch's (<= kCrossMaxL) layers starting at the given pointers
  __shared__ float red[8][32][kCrossMaxL + 1];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int r_begin = blockIdx.y * rows_per_block;
  const int r_end = min(B, r_begin + rows_per_block);
  float acc[kCrossMaxL + 1];
#pragma unroll
  for (int l = 0; l <= kCrossMaxL; ++l) acc[l] = 0.f;
  if (c < W) {
    for (int r = r_begin + threadIdx.y; r < r_end; r += 8) {
      const float x = X[(int64_t)r * W + c];
      if (want_s) acc[kCrossMaxL] += dY[(int64_t)r * W + c];
#pragma unroll
      for (int l = 0; l < kCrossMaxL; ++l)
        if (l < n_layers) acc[l] = fmaf(x, __ldg(coef + (int64_t)r * ld + l), acc[l]);
    }
  }
#pragma unroll
  for (int l = 0; l <= kCrossMaxL; ++l) red[threadIdx.y][threadIdx.x][l] = acc[l];
  __syncthreads();
  if (threadIdx.y == 0 && c < W) {
    for (int l = 0; l <= kCrossMaxL; ++l) {
      float v = 0.f;
      for (int j = 0; j < 8; ++j) v += red[j][threadIdx.x][l];
      if (l < n_layers) atomicAdd(m1 + (size_t)c * ld + l, v);
      else if (l == kCrossMaxL && want_s) atomicAdd(s + c, v);
    }
  }
}

// G[l] = sum_rows gx[row, l]
__global__ void cross_gsum_kernel(const float* __restrict__ gx, float* __restrict__ G, int B, int n_layers, int ld) {
  float acc[kCrossMaxL];
#pragma unroll
  for (int l = 0; l < kCrossMaxL; ++l) acc[l] = 0.f;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < B; r += gridDim.x * blockDim.x)
#pragma unroll
    for (int l = 0; l < kCrossMaxL; ++l)
      if (l < n_layers) acc[l] += gx[(int64_t)r * ld + l];
#pragma unroll
  for (int l = 0; l < kCrossMaxL; ++l) {
    const float v = warp_sum(acc[l]);
    if ((threadIdx.x & 31) == 0 && l < n_layers && v != 0.f) atomicAdd(G + l, v);
  }
}

__global__ void cross_combine_kernel(const float* __restrict__ m1, const float* __restrict__ s, const float* __restrict__ G,
                                     const float* __restrict__ kernels, const float* __restrict__ biases,
                                     float* __restrict__ d_kernels, float* __restrict__ d_biases, int W, int n_layers) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= W) return;
  float bsum = 0.f;           // Bsum_l[c]
  for (int l = 0; l < n_layers; ++l) {
    d_kernels[(size_t)l * W + c] += m1[(size_t)c * n_layers + l] + bsum * G[l];
    bsum += biases[(size_t)l * W + c];
  }
  float tail = 0.f;           // sum_{l'>l} w_l'[c] G_l'
  for (int l = n_layers - 1; l >= 0; --l) {
    d_biases[(size_t)l * W + c] += s[c] + tail;
    tail += kernels[(size_t)l * W + c] * G[l];
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
  if (W <= 1024) {
    int blocks = ceil_div(B, 8);
    const int capr = sm_count() * 8;
    if (blocks > capr) blocks = capr;
    if (W <= 128)
      cross_fwd_reg_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(X, kernels, biases, Y, xw_saved, B, W, n_layers);
    else if (W <= 512)
      cross_fwd_reg_kernel<16><<<blocks, 256, 0, (cudaStream_t)stream>>>(X, kernels, biases, Y, xw_saved, B, W, n_layers);
    else
      cross_fwd_reg_kernel<32><<<blocks, 256, 0, (cudaStream_t)stream>>>(X, kernels, biases, Y, xw_saved, B, W, n_layers);
    DTB_LAUNCH_OK();
    return DTB_OK;
  }
  // wider rows: x0 / x_l of a row in shared memory, as many warps per CTA (<= 4) as the budget holds
  int warps = (int)((200 * 1024) / ((size_t)2 * W * sizeof(float)));
  if (warps > kCrossWarps) warps = kCrossWarps;
  if (warps < 1) {
    set_error("dtb_cross_fwd: input width %d too large for the shared-memory row buffers", W);
    return DTB_ERR_UNSUPPORTED;
  }
  const size_t smem = (size_t)warps * 2 * W * sizeof(float);
  DTB_CUDA_OK(cudaFuncSetAttribute(cross_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int blocks = ceil_div(B, warps);
  const int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  cross_fwd_kernel<<<blocks, warps * 32, smem, (cudaStream_t)stream>>>(X, kernels, biases, Y, xw_saved, B, W, n_layers);
  DTB_LAUNCH_OK();
  return DTB_OK;
}

size_t dtb_cross_bwd_workspace_bytes(int B, int W, int n_layers) {
  if (B <= 0 || W <= 0 || n_layers <= 0) return 0;
  return ((size_t)2 * B * n_layers + (size_t)W * n_layers + W + n_layers) * sizeof(float) + 64;
}

int dtb_cross_bwd(const float* X, const float* kernels, const float* biases, const float* xw_saved,
                  const float* dY, float* dX, float* d_kernels, float* d_biases, void* workspace,
                  size_t workspace_bytes, int B, int W, int n_layers, void* stream) {
  DTB_CHECK_ARG(X && kernels && biases && xw_saved && dY && dX && d_kernels && d_biases, "NULL argument");
  DTB_CHECK_ARG(B >= 0 && W > 0 && n_layers > 0, "bad shape");
  if (B == 0) return DTB_OK;
  if (!workspace || workspace_bytes < dtb_cross_bwd_workspace_bytes(B, W, n_layers)) {
    set_error("dtb_cross_bwd: workspace missing or smaller than dtb_cross_bwd_workspace_bytes()");
    return DTB_ERR_INVALID_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* coef = reinterpret_cast<float*>(workspace);
  float* gx = coef + (size_t)B * n_layers;
  float* m1 = gx + (size_t)B * n_layers;
  float* s = m1 + (size_t)W * n_layers;
  float* G = s + W;
  DTB_CUDA_OK(cudaMemsetAsync(m1, 0, ((size_t)W * n_layers + W + n_layers) * sizeof(float), st));
  if (W <= 1024 && n_layers <= kCrossMaxL) {
    int blocks = ceil_div(B, 8);
    const int capr = sm_count() * 8;
    if (blocks > capr) blocks = capr;
    if (W <= 128)
      cross_bwd_reg_kernel<4><<<blocks, 256, 0, st>>>(X, kernels, xw_saved, dY, dX, coef, gx, B, W, n_layers);
    else if (W <= 512)
      cross_bwd_reg_kernel<16><<<blocks, 256, 0, st>>>(X, kernels, xw_saved, dY, dX, coef, gx, B, W, n_layers);
    else
      cross_bwd_reg_kernel<32><<<blocks, 256, 0, st>>>(X, kernels, xw_saved, dY, dX, coef, gx, B, W, n_layers);
  } else {
    int warps = (int)((200 * 1024) / ((size_t)3 * W * sizeof(float)));
    if (warps > kCrossWarps) warps = kCrossWarps;
    if (warps < 1) {
      set_error("dtb_cross_bwd: input width %d too large for the shared-memory row buffers", W);
      return DTB_ERR_UNSUPPORTED;
    }
    const size_t smem = (size_t)warps * 3 * W * sizeof(float);
    DTB_CUDA_OK(cudaFuncSetAttribute(cross_bwd_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = ceil_div(B, warps);
    const int cap = sm_count() * 8;
    if (blocks > cap) blocks = cap;
    cross_bwd_smem_kernel<<<blocks, warps * 32, smem, st>>>(X, kernels, xw_saved, dY, dX, coef, gx, B, W, n_layers);
  }
  DTB_LAUNCH_OK();
  const int col_blocks = ceil_div(W, 32);
  int row_blocks = ceil_div((int64_t)sm_count() * 8, col_blocks);
  if (row_blocks > ceil_div(B, 64)) row_blocks = ceil_div(B, 64);
  if (row_blocks < 1) row_blocks = 1;
  const int rpb = ceil_div(B, row_blocks);
  int gb = ceil_div(B, 256 * 8);
  if (gb > sm_count()) gb = sm_count();
  for (int l0 = 0; l0 < n_layers; l0 += kCrossMaxL) {       // the reductions hold <= kCrossMaxL layers in registers
    const int nl = n_layers - l0 < kCrossMaxL ? n_layers - l0 : kCrossMaxL;
    cross_colreduce_kernel<<<dim3(col_blocks, ceil_div(B, rpb)), dim3(32, 8), 0, st>>>(X, dY, coef + l0, m1 + l0, s, B, W, nl,
                                                                                       rpb, n_layers, l0 == 0);
    DTB_LAUNCH_OK();
    cross_gsum_kernel<<<gb < 1 ? 1 : gb, 256, 0, st>>>(gx + l0, G + l0, B, nl, n_layers);
    DTB_LAUNCH_OK();
  }
  cross_combine_kernel<<<ceil_div(W, 128), 128, 0, st>>>(m1, s, G, kernels, biases, d_kernels, d_biases, W, n_layers);
  DTB_LAUNCH_OK();
  return DTB_OK;
}

}  // extern "C"
