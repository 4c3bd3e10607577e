// FiBiNet pieces (reference layers.py:245-382; fibi_nets / fibi_dnn_nets deepnets.py:344-386) on a dense [B, F, D] block
// (the concatenated embeddings, or their SENET re-weighting):
//
//   SENET               Z[b,f] = mean_d (or max_d) X[b,f,d];  A = relu(Dense_F(relu(Dense_r(Z))))  (the two tiny Dense layers
//                       are ordinary Dense calls of this library);  V[b,f,d] = X[b,f,d] A[b,f]
//   BilinearInteraction out[b,p,:] = (x_i W_s(p)) * x_j  for the P = F(F-1)/2 pairs (i < j, itertools.combinations order);
//                       W [n_w, D, D]:  field_all s = 0 | field_each s = i | field_interaction s = p
//
// Same structure as the PNN 'mat' kernels (pnn_attention.cu) with a D-vector per pair instead of a scalar: forward = thread per
// (row, first field), data gradient = thread per (row, field) with register accumulators and plain stores (no atomics: the block
// is dense), weight gradient = thread per (pair, a) streaming row chunks out of shared memory.  20.8 KB of output per row at
// F = 26, D = 16: the forward and the data gradient are bound by writing / reading that tensor.
#include "dtb_common.cuh"

namespace dtb {

constexpr int kBiRows = 128;

__device__ __forceinline__ int bi_pair_index(int a, int b, int F) { return a * (F - 1) - a * (a - 1) / 2 + (b - a - 1); }
__device__ __forceinline__ void bi_pair_of(int p, int F, int& i, int& j) {
  int ii = 0, rem = p;
  while (rem >= F - 1 - ii) {
    rem -= F - 1 - ii;
    ++ii;
  }
  i = ii;
  j = ii + 1 + rem;
}
__device__ __forceinline__ int bi_slot(int type, int i, int p) { return type == 0 ? 0 : (type == 1 ? i : p); }

template <int N>
__device__ __forceinline__ void bi_ld(const float* __restrict__ p, float (&o)[N]) {
#pragma unroll
  for (int c = 0; c < N / 4; ++c) {
    const float4 t = *reinterpret_cast<const float4*>(p + 4 * c);
    o[4 * c] = t.x;
    o[4 * c + 1] = t.y;
    o[4 * c + 2] = t.z;
    o[4 * c + 3] = t.w;
  }
}
template <int N>
__device__ __forceinline__ void bi_st(float* __restrict__ p, const float (&v)[N]) {
#pragma unroll
  for (int c = 0; c < N / 4; ++c)
    *reinterpret_cast<float4*>(p + 4 * c) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}

// t[b] = sum_a x[a] W[a][b], W a [DT x DT] slab in shared memory (rows read as warp broadcasts)
template <int DT>
__device__ __forceinline__ void bi_xw(const float (&x)[DT], const float* __restrict__ w, float (&t)[DT]) {
#pragma unroll
  for (int b = 0; b < DT; ++b) t[b] = 0.f;
#pragma unroll
  for (int a = 0; a < DT; ++a) {
    float wr[DT];
    bi_ld<DT>(w + a * DT, wr);
#pragma unroll
    for (int b = 0; b < DT; ++b) t[b] = fmaf(x[a], wr[b], t[b]);
  }
}

// grid (F-1, row-chunk groups); CTA x = first field i; thread = row.  Shared memory: the slabs this field needs.
template <int DT>
__global__ void __launch_bounds__(kBiRows) bilinear_fwd_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                               float* __restrict__ out, int B, int F, int P, int type) {
  extern __shared__ __align__(16) float ws[];
  const int i = blockIdx.x, n = F - 1 - i;
  const int p0 = bi_pair_index(i, i + 1, F);
  const int n_slab = type == 2 ? n : 1;
  for (int e = threadIdx.x; e < n_slab * DT * DT; e += blockDim.x) {
    const int q = e / (DT * DT);
    ws[e] = __ldg(W + (size_t)bi_slot(type, i, p0 + q) * DT * DT + (e - q * DT * DT));
  }
  __syncthreads();
  const int n_chunks = (B + kBiRows - 1) / kBiRows;
  for (int chunk = blockIdx.y; chunk < n_chunks; chunk += gridDim.y) {
    const int row = chunk * kBiRows + threadIdx.x;
    if (row >= B) continue;
    const float* xr = X + (size_t)row * F * DT;
    float xi[DT], xj[DT], t[DT];
    bi_ld<DT>(xr + i * DT, xi);
    if (type != 2) bi_xw<DT>(xi, ws, t);
    for (int q = 0; q < n; ++q) {
      bi_ld<DT>(xr + (i + 1 + q) * DT, xj);
      if (type == 2) bi_xw<DT>(xi, ws + (size_t)q * DT * DT, t);
      float o[DT];
#pragma unroll
      for (int b = 0; b < DT; ++b) o[b] = t[b] * xj[b];
      bi_st<DT>(out + ((size_t)row * P + p0 + q) * DT, o);
    }
  }
}

// grid (F, row-chunk groups); thread = (row, field f): d x_f over the F-1 pairs that contain f.
//   f first  (f, o):  d x_f[a] += sum_b W[a][b] g[b] x_o[b]
//   f second (o, f):  d x_f[b] += g[b] sum_a x_o[a] W[a][b]
template <int DT>
__global__ void __launch_bounds__(kBiRows) bilinear_bwd_dx_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                                  const float* __restrict__ dOut, float* __restrict__ dX, int B,
                                                                  int F, int P, int type) {
  extern __shared__ __align__(16) float ws[];        // [F-1 visits][DT][DT]
  const int f = blockIdx.x;
  for (int e = threadIdx.x; e < (F - 1) * DT * DT; e += blockDim.x) {
    const int q = e / (DT * DT);
    const int o = q < f ? q : q + 1;
    const int a = o < f ? o : f, b = o < f ? f : o;
    ws[e] = __ldg(W + (size_t)bi_slot(type, a, bi_pair_index(a, b, F)) * DT * DT + (e - q * DT * DT));
  }
  __syncthreads();
  const int n_chunks = (B + kBiRows - 1) / kBiRows;
  for (int chunk = blockIdx.y; chunk < n_chunks; chunk += gridDim.y) {
    const int row = chunk * kBiRows + threadIdx.x;
    if (row >= B) continue;
    const float* xr = X + (size_t)row * F * DT;
    float acc[DT];
#pragma unroll
    for (int d = 0; d < DT; ++d) acc[d] = 0.f;
    for (int q = 0; q < F - 1; ++q) {
      const int o = q < f ? q : q + 1;
      const int p = o < f ? bi_pair_index(o, f, F) : bi_pair_index(f, o, F);
      float xo[DT], g[DT];
      bi_ld<DT>(xr + o * DT, xo);
      bi_ld<DT>(dOut + ((size_t)row * P + p) * DT, g);
      const float* w = ws + (size_t)q * DT * DT;
      if (f < o) {
#pragma unroll
        for (int b = 0; b < DT; ++b) g[b] *= xo[b];
#pragma unroll
        for (int a = 0; a < DT; ++a) {
          float wr[DT];
          bi_ld<DT>(w + a * DT, wr);
          float s = 0.f;
#pragma unroll
          for (int b = 0; b < DT; ++b) s = fmaf(wr[b], g[b], s);
          acc[a] += s;
        }
      } else {
        float t[DT];
        bi_xw<DT>(xo, w, t);
#pragma unroll
        for (int b = 0; b < DT; ++b) acc[b] = fmaf(g[b], t[b], acc[b]);
      }
    }
    bi_st<DT>(dX + ((size_t)row * F + f) * DT, acc);
  }
}

// d W_s[a][b] += sum_rows x_i[a] g_p[b] x_j[b].  grid (pair groups, row groups), 256 threads: thread = (pair of the group, a);
// rows arrive in chunks of R: the X rows and the group's slice of dOut in shared memory.
template <int DT>
__global__ void __launch_bounds__(256) bilinear_bwd_dw_kernel(const float* __restrict__ X, const float* __restrict__ dOut,
                                                              float* __restrict__ dW, int B, int F, int P, int type,
                                                              int rows_per_cta, int R) {
  constexpr int kPg = 256 / DT;
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                                  // [R][F][DT]
  float* gs = sm + (size_t)R * F * DT;             // [R][kPg][DT]
  const int pl = threadIdx.x / DT, a = threadIdx.x - pl * DT;
  const int p = blockIdx.x * kPg + pl;
  const bool live = p < P;
  int i = 0, j = 1;
  if (live) bi_pair_of(p, F, i, j);
  float acc[DT];
#pragma unroll
  for (int b = 0; b < DT; ++b) acc[b] = 0.f;
  const int r_begin = blockIdx.y * rows_per_cta;
  const int r_end = min(B, r_begin + rows_per_cta);
  const int pg0 = blockIdx.x * kPg, npg = min(kPg, P - pg0);
  for (int r0 = r_begin; r0 < r_end; r0 += R) {
    const int nr = min(R, r_end - r0);
    __syncthreads();
    const float4* xsrc = reinterpret_cast<const float4*>(X + (size_t)r0 * F * DT);
    for (int e = threadIdx.x; e < nr * F * (DT / 4); e += blockDim.x) reinterpret_cast<float4*>(xs)[e] = __ldg(xsrc + e);
    for (int e = threadIdx.x; e < nr * npg * (DT / 4); e += blockDim.x) {
      const int r = e / (npg * (DT / 4)), rem = e - r * npg * (DT / 4);
      const float4* gsrc = reinterpret_cast<const float4*>(dOut + ((size_t)(r0 + r) * P + pg0) * DT);
      reinterpret_cast<float4*>(gs + (size_t)r * kPg * DT)[rem] = __ldg(gsrc + rem);
    }
    __syncthreads();
    if (live) {
      for (int r = 0; r < nr; ++r) {
        const float* xr = xs + (size_t)r * F * DT;
        const float xa = xr[i * DT + a];
        float xj[DT], g[DT];
        bi_ld<DT>(xr + j * DT, xj);
        bi_ld<DT>(gs + ((size_t)r * kPg + pl) * DT, g);
#pragma unroll
        for (int b = 0; b < DT; ++b) acc[b] = fmaf(xa, g[b] * xj[b], acc[b]);
      }
    }
  }
  if (live) {
    float4* dst = reinterpret_cast<float4*>(dW + ((size_t)bi_slot(type, i, p) * DT + a) * DT);
#pragma unroll
    for (int c = 0; c < DT / 4; ++c) {
      const float4 v = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
      if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) atomicAdd(dst + c, v);
    }
  }
}

// ---- SENET ---------------------------------------------------------------------------------------------------------
// op 0 = mean, 1 = max over the embedding axis; thread = (b, f)
__global__ void senet_pool_fwd_kernel(const float* __restrict__ X, float* __restrict__ Z, int64_t BF, int D, int op) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= BF) return;
  const float* x = X + t * D;
  float s = op ? -INFINITY : 0.f;
  for (int d = 0; d < D; ++d) s = op ? fmaxf(s, __ldg(x + d)) : s + __ldg(x + d);
  Z[t] = op ? s : s / (float)D;
}
// dX[b,f,d] = dZ[b,f] / D (mean) or dZ[b,f] [x == max] / (number of maxima) (max, ties share the gradient as in
// tf.reduce_max's gradient)
__global__ void senet_pool_bwd_kernel(const float* __restrict__ X, const float* __restrict__ Z, const float* __restrict__ dZ,
                                      float* __restrict__ dX, int64_t BF, int D, int op) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= BF) return;
  const float g = __ldg(dZ + t);
  float* o = dX + t * D;
  if (!op) {
    const float v = g / (float)D;
    for (int d = 0; d < D; ++d) o[d] = v;
  } else {
    const float* x = X + t * D;
    const float m = __ldg(Z + t);
    int n = 0;
    for (int d = 0; d < D; ++d) n += __ldg(x + d) == m;
    const float v = n ? g / (float)n : 0.f;
    for (int d = 0; d < D; ++d) o[d] = __ldg(x + d) == m ? v : 0.f;
  }
}
// V = X * A[..., None]; thread = (b, f)
__global__ void senet_scale_fwd_kernel(const float* __restrict__ X, const float* __restrict__ A, float* __restrict__ V, int64_t BF,
                                       int D) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= BF) return;
  const float a = __ldg(A + t);
  for (int d = 0; d < D; ++d) V[t * D + d] = __ldg(X + t * D + d) * a;
}
// dX = dV * A,  dA[b,f] = sum_d dV X
__global__ void senet_scale_bwd_kernel(const float* __restrict__ X, const float* __restrict__ A, const float* __restrict__ dV,
                                       float* __restrict__ dX, float* __restrict__ dA, int64_t BF, int D) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= BF) return;
  const float a = __ldg(A + t);
  float s = 0.f;
  for (int d = 0; d < D; ++d) {
    const float g = __ldg(dV + t * D + d);
    s = fmaf(g, __ldg(X + t * D + d), s);
    dX[t * D + d] = g * a;
  }
  dA[t] = s;
}

}  // namespace dtb

using namespace dtb;

namespace {
// vector accesses touch the [B, F, D] / [B, P, D] blocks; the weights are read one float at a time (a parameter is a view
// into the model's flat buffer and only 4-byte aligned)
bool bi_shape(int D, const void* a, const void* b) {
  if (D != 4 && D != 8 && D != 16 && D != 32) return false;
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
}
constexpr size_t kBiSmemMax = 200 * 1024;
}  // namespace

#define DTB_BI_DISPATCH(D, ...)                       \
  switch (D) {                                        \
    case 4: { constexpr int DT_ = 4; __VA_ARGS__; } break;   \
    case 8: { constexpr int DT_ = 8; __VA_ARGS__; } break;   \
    case 16: { constexpr int DT_ = 16; __VA_ARGS__; } break; \
    case 32: { constexpr int DT_ = 32; __VA_ARGS__; } break; \
    default: break;                                   \
  }

extern "C" {

int dtb_bilinear_fwd(const float* X, const float* W, float* out, int B, int F, int D, int bilinear_type, void* stream) {
  DTB_CHECK_ARG(X && W && out, "NULL argument");
  DTB_CHECK_ARG(F >= 2 && B >= 0 && bilinear_type >= 0 && bilinear_type <= 2, "need F >= 2 and bilinear_type 0 (all), 1 (each) or 2 (interaction)");
  if (!bi_shape(D, X, out) || (size_t)(F - 1) * D * D * sizeof(float) > kBiSmemMax) {
    set_error("dtb_bilinear_fwd: needs D in {4, 8, 16, 32}, 16-byte aligned buffers and (F-1) D^2 floats of shared memory (F = %d, D = %d)", F, D);
    return DTB_ERR_UNSUPPORTED;
  }
  if (B == 0) return DTB_OK;
  const int P = F * (F - 1) / 2;
  const size_t smem = (size_t)(bilinear_type == 2 ? F - 1 : 1) * D * D * sizeof(float);
  int groups = ceil_div(sm_count() * 8, F - 1);
  if (groups > ceil_div(B, kBiRows)) groups = ceil_div(B, kBiRows);
  DTB_BI_DISPATCH(D, {
    auto k = bilinear_fwd_kernel<DT_>;
    DTB_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<dim3(F - 1, groups), kBiRows, smem, (cudaStream_t)stream>>>(X, W, out, B, F, P, bilinear_type);
  })
  DTB_LAUNCH_OK();
  return DTB_OK;
}

int dtb_bilinear_bwd(const float* X, const float* W, const float* d_out, float* dX, float* dW, int B, int F, int D,
                     int bilinear_type, void* stream) {
  DTB_CHECK_ARG(X && W && d_out && dW, "NULL argument");
  DTB_CHECK_ARG(F >= 2 && B >= 0 && bilinear_type >= 0 && bilinear_type <= 2, "need F >= 2 and bilinear_type 0 (all), 1 (each) or 2 (interaction)");
  if (!bi_shape(D, X, d_out) || ((reinterpret_cast<uintptr_t>(dX) | reinterpret_cast<uintptr_t>(dW)) & 15) ||
      (size_t)(F - 1) * D * D * sizeof(float) > kBiSmemMax) {
    set_error("dtb_bilinear_bwd: needs D in {4, 8, 16, 32}, 16-byte aligned buffers and (F-1) D^2 floats of shared memory (F = %d, D = %d)", F, D);
    return DTB_ERR_UNSUPPORTED;
  }
  if (B == 0) return DTB_OK;
  const int P = F * (F - 1) / 2;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem_x = (size_t)(F - 1) * D * D * sizeof(float);
  int groups = ceil_div(sm_count() * 8, F);
  if (groups > ceil_div(B, kBiRows)) groups = ceil_div(B, kBiRows);
  const int pg = 256 / D, pgroups = ceil_div(P, pg);
  int R = 32;
  while (R > 2 && (size_t)R * (F + pg) * D * sizeof(float) > kBiSmemMax / 2) R /= 2;
  const size_t smem_w = (size_t)R * (F + pg) * D * sizeof(float);
  if (smem_w > kBiSmemMax) {
    set_error("dtb_bilinear_bwd: %d fields do not fit the weight-gradient staging", F);
    return DTB_ERR_UNSUPPORTED;
  }
  int row_groups = ceil_div(sm_count() * 2, pgroups);
  if (row_groups > ceil_div(B, R)) row_groups = ceil_div(B, R);
  const int rows_per_cta = ceil_div(ceil_div(B, row_groups), R) * R;
  DTB_BI_DISPATCH(D, {
    if (dX) {
      auto k = bilinear_bwd_dx_kernel<DT_>;
      DTB_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x));
      k<<<dim3(F, groups), kBiRows, smem_x, st>>>(X, W, d_out, dX, B, F, P, bilinear_type);
    }
    auto kw = bilinear_bwd_dw_kernel<DT_>;
    DTB_CUDA_OK(cudaFuncSetAttribute(kw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
    kw<<<dim3(pgroups, ceil_div(B, rows_per_cta)), 256, smem_w, st>>>(X, d_out, dW, B, F, P, bilinear_type, rows_per_cta, R);
  })
  DTB_LAUNCH_OK();
  return DTB_OK;
}

int dtb_senet_pool_fwd(const float* X, float* Z, int B, int F, int D, int pooling_op, void* stream) {
  DTB_CHECK_ARG(X && Z && B >= 0 && F >= 1 && D >= 1 && (pooling_op == 0 || pooling_op == 1), "bad argument (pooling_op 0 mean, 1 max)");
  const int64_t bf = (int64_t)B * F;
  if (bf == 0) return DTB_OK;
  senet_pool_fwd_kernel<<<ceil_div(bf, 256), 256, 0, (cudaStream_t)stream>>>(X, Z, bf, D, pooling_op);
  DTB_LAUNCH_OK();
  return DTB_OK;
}

int dtb_senet_pool_bwd(const float* X, const float* Z, const float* dZ, float* dX, int B, int F, int D, int pooling_op,
                       void* stream) {
  DTB_CHECK_ARG(X && Z && dZ && dX && B >= 0 && F >= 1 && D >= 1 && (pooling_op == 0 || pooling_op == 1), "bad argument");
  const int64_t bf = (int64_t)B * F;
  if (bf == 0) return DTB_OK;
  senet_pool_bwd_kernel<<<ceil_div(bf, 256), 256, 0, (cudaStream_t)stream>>>(X, Z, dZ, dX, bf, D, pooling_op);
  DTB_LAUNCH_OK();
  return DTB_OK;
}

int dtb_senet_scale_fwd(const float* X, const float* A, float* V, int B, int F, int D, void* stream) {
  DTB_CHECK_ARG(X && A && V && B >= 0 && F >= 1 && D >= 1, "bad argument");
  const int64_t bf = (int64_t)B * F;
  if (bf == 0) return DTB_OK;
  senet_scale_fwd_kernel<<<ceil_div(bf, 256), 256, 0, (cudaStream_t)stream>>>(X, A, V, bf, D);
  DTB_LAUNCH_OK();
  return DTB_OK;
}

int dtb_senet_scale_bwd(const float* X, const float* A, const float* dV, float* dX, float* dA, int B, int F, int D,
                        void* stream) {
  DTB_CHECK_ARG(X && A && dV && dX && dA && B >= 0 && F >= 1 && D >= 1, "bad argument");
  const int64_t bf = (int64_t)B * F;
  if (bf == 0) return DTB_OK;
  senet_scale_bwd_kernel<<<ceil_div(bf, 256), 256, 0, (cudaStream_t)stream>>>(X, A, dV, dX, dA, bf, D);
  DTB_LAUNCH_OK();
  return DTB_OK;
}

}  // extern "C"
