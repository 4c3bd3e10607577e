// PNN products (InnerProduct / OuterProduct, layers.py:473-487, 541-581; gather fused) and the
// AutoInt attention core (MultiheadAttention.call, layers.py:129-150: per-head softmax(QK^T/sqrt(dh))V
// + residual + relu; the four relu(Dense) projections and the BatchNormalization around it are Dense /
// BatchNorm calls of this library).  Both are small-F pairwise ops on CUDA cores: per batch row the
// (F x D) block lives in shared memory; there is no large dense contraction to put on tensor cores.
#include "dtb_common.cuh"

namespace dtb {

constexpr int kPnnRows = 8;      // batch rows per CTA (the outer-product kernel is reused across them)
constexpr int kMaxD = 64;

__device__ __forceinline__ void pair_of(int p, int F, int& i, int& j) {
  // pairs (i<j) in row-major order (layers.py:478-483)
  int ii = 0, rem = p;
  while (rem >= F - 1 - ii) {
    rem -= F - 1 - ii;
    ++ii;
  }
  i = ii;
  j = ii + 1 + rem;
}

__device__ __forceinline__ void stage_rows(const int32_t* __restrict__ idx, const float* __restrict__ table,
                                           const int64_t* __restrict__ row_offsets, float* __restrict__ es, int row0,
                                           int n_rows, int B, int F, int D, int* status) {
  for (int e = threadIdx.x; e < n_rows * F * D; e += blockDim.x) {
    const int r = e / (F * D);
    const int rem = e - r * F * D;
    const int f = rem / D, d = rem - f * D;
    float v = 0.f;
    if (row0 + r < B) {
      const int64_t rb = table_row(row_offsets, f, __ldg(idx + (int64_t)(row0 + r) * F + f), D, status);
      if (rb >= 0) v = __ldg(table + rb + d);
    }
    es[e] = v;
  }
}

// thread = pair; loops over the CTA's kPnnRows rows with the kernel slice of the pair held per k-step
__global__ void pnn_fwd_kernel(const int32_t* __restrict__ idx, const float* __restrict__ table,
                               const int64_t* __restrict__ row_offsets, const float* __restrict__ kern,
                               float* __restrict__ ip, float* __restrict__ op, int B, int F, int D, int P, int ktype,
                               int* status) {
  extern __shared__ float es[];    // [rows][F][D]
  const int row0 = blockIdx.x * kPnnRows;
  stage_rows(idx, table, row_offsets, es, row0, kPnnRows, B, F, D, status);
  __syncthreads();
  const int p = threadIdx.x;
  if (p >= P) return;
  int i, j;
  pair_of(p, F, i, j);
  float acc_ip[kPnnRows], acc_op[kPnnRows];
#pragma unroll
  for (int r = 0; r < kPnnRows; ++r) {
    acc_ip[r] = 0.f;
    acc_op[r] = 0.f;
    const float* ei = es + ((size_t)r * F + i) * D;
    const float* ej = es + ((size_t)r * F + j) * D;
    float s = 0.f;
    for (int d = 0; d < D; ++d) s += ei[d] * ej[d];
    acc_ip[r] = s;
  }
  if (op) {
    if (ktype == 0) {
      for (int k = 0; k < D; ++k) {
        const float* kr = kern + ((size_t)k * P + p) * D;
        float t[kPnnRows];
#pragma unroll
        for (int r = 0; r < kPnnRows; ++r) t[r] = 0.f;
        for (int d = 0; d < D; ++d) {
          const float kv = __ldg(kr + d);
#pragma unroll
          for (int r = 0; r < kPnnRows; ++r) t[r] += es[((size_t)r * F + i) * D + d] * kv;
        }
#pragma unroll
        for (int r = 0; r < kPnnRows; ++r) acc_op[r] += t[r] * es[((size_t)r * F + j) * D + k];
      }
    } else if (ktype == 1) {
      for (int d = 0; d < D; ++d) {
        const float kv = __ldg(kern + (size_t)p * D + d);
#pragma unroll
        for (int r = 0; r < kPnnRows; ++r)
          acc_op[r] += es[((size_t)r * F + i) * D + d] * es[((size_t)r * F + j) * D + d] * kv;
      }
    } else {
      const float kv = __ldg(kern + p);
#pragma unroll
      for (int r = 0; r < kPnnRows; ++r) acc_op[r] = acc_ip[r] * kv;
    }
  }
#pragma unroll
  for (int r = 0; r < kPnnRows; ++r) {
    if (row0 + r < B) {
      if (ip) ip[(size_t)(row0 + r) * P + p] = acc_ip[r];
      if (op) op[(size_t)(row0 + r) * P + p] = acc_op[r];
    }
  }
}

// dE: thread = pair, per-row contributions accumulated in shared memory, then one RED per element
__global__ void pnn_bwd_de_kernel(const int32_t* __restrict__ idx, const float* __restrict__ table,
                                  const int64_t* __restrict__ row_offsets, const float* __restrict__ kern,
                                  const float* __restrict__ d_ip, const float* __restrict__ d_op,
                                  float* __restrict__ grad_table, int B, int F, int D, int P, int ktype) {
  extern __shared__ float sm[];
  float* es = sm;                                   // [rows][F][D]
  float* des = sm + (size_t)kPnnRows * F * D;       // [rows][F][D]
  const int row0 = blockIdx.x * kPnnRows;
  stage_rows(idx, table, row_offsets, es, row0, kPnnRows, B, F, D, nullptr);
  for (int e = threadIdx.x; e < kPnnRows * F * D; e += blockDim.x) des[e] = 0.f;
  __syncthreads();
  const int p = threadIdx.x;
  if (p < P) {
    int i, j;
    pair_of(p, F, i, j);
    for (int r = 0; r < kPnnRows; ++r) {
      if (row0 + r >= B) break;
      const float gi = d_ip ? d_ip[(size_t)(row0 + r) * P + p] : 0.f;
      const float go = d_op ? d_op[(size_t)(row0 + r) * P + p] : 0.f;
      const float* ei = es + ((size_t)r * F + i) * D;
      const float* ej = es + ((size_t)r * F + j) * D;
      float* di = des + ((size_t)r * F + i) * D;
      float* dj = des + ((size_t)r * F + j) * D;
      if (d_ip)
        for (int d = 0; d < D; ++d) {
          atomicAdd(di + d, gi * ej[d]);
          atomicAdd(dj + d, gi * ei[d]);
        }
      if (d_op) {
        if (ktype == 0) {
          // op = sum_k (sum_d ei[d] K[k,p,d]) ej[k]
          float gi_acc[kMaxD];
          for (int d = 0; d < D; ++d) gi_acc[d] = 0.f;
          for (int k = 0; k < D; ++k) {
            const float* kr = kern + ((size_t)k * P + p) * D;
            float t = 0.f;
            const float ejk = ej[k];
            for (int d = 0; d < D; ++d) {
              const float kv = __ldg(kr + d);
              t += ei[d] * kv;
              gi_acc[d] += kv * ejk;
            }
            atomicAdd(dj + k, go * t);
          }
          for (int d = 0; d < D; ++d) atomicAdd(di + d, go * gi_acc[d]);
        } else if (ktype == 1) {
          for (int d = 0; d < D; ++d) {
            const float kv = __ldg(kern + (size_t)p * D + d);
            atomicAdd(di + d, go * ej[d] * kv);
            atomicAdd(dj + d, go * ei[d] * kv);
          }
        } else {
          const float kv = __ldg(kern + p);
          for (int d = 0; d < D; ++d) {
            atomicAdd(di + d, go * kv * ej[d]);
            atomicAdd(dj + d, go * kv * ei[d]);
          }
        }
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < kPnnRows * F * D; e += blockDim.x) {
    const int r = e / (F * D);
    if (row0 + r >= B) continue;
    const int rem = e - r * F * D;
    const int f = rem / D, d = rem - f * D;
    const int64_t rb = table_row(row_offsets, f, __ldg(idx + (int64_t)(row0 + r) * F + f), D, nullptr);
    if (rb >= 0 && des[e] != 0.f) atomicAdd(grad_table + rb + d, des[e]);
  }
}

// dK: persistent CTAs, thread = pair, k outermost so the accumulator is D registers; one atomic per
// kernel element per CTA
__global__ void pnn_bwd_dk_kernel(const int32_t* __restrict__ idx, const float* __restrict__ table,
                                  const int64_t* __restrict__ row_offsets, const float* __restrict__ d_op,
                                  float* __restrict__ d_kern, int B, int F, int D, int P, int ktype,
                                  int rows_per_cta) {
  const int p = threadIdx.x;
  if (p >= P) return;
  int i, j;
  pair_of(p, F, i, j);
  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(B, r_begin + rows_per_cta);
  if (ktype == 0) {
    for (int k = 0; k < D; ++k) {
      float acc[kMaxD];
      for (int d = 0; d < D; ++d) acc[d] = 0.f;
      for (int r = r_begin; r < r_end; ++r) {
        const int64_t bi = table_row(row_offsets, i, __ldg(idx + (int64_t)r * F + i), D, nullptr);
        const int64_t bj = table_row(row_offsets, j, __ldg(idx + (int64_t)r * F + j), D, nullptr);
        if (bi < 0 || bj < 0) continue;
        const float w = d_op[(size_t)r * P + p] * __ldg(table + bj + k);
        for (int d = 0; d < D; ++d) acc[d] += w * __ldg(table + bi + d);
      }
      for (int d = 0; d < D; ++d)
        if (acc[d] != 0.f) atomicAdd(d_kern + ((size_t)k * P + p) * D + d, acc[d]);
    }
  } else {
    float acc[kMaxD];
    const int n = ktype == 1 ? D : 1;
    for (int d = 0; d < n; ++d) acc[d] = 0.f;
    for (int r = r_begin; r < r_end; ++r) {
      const int64_t bi = table_row(row_offsets, i, __ldg(idx + (int64_t)r * F + i), D, nullptr);
      const int64_t bj = table_row(row_offsets, j, __ldg(idx + (int64_t)r * F + j), D, nullptr);
      if (bi < 0 || bj < 0) continue;
      const float go = d_op[(size_t)r * P + p];
      if (ktype == 1) {
        for (int d = 0; d < D; ++d) acc[d] += go * __ldg(table + bi + d) * __ldg(table + bj + d);
      } else {
        float s = 0.f;
        for (int d = 0; d < D; ++d) s += __ldg(table + bi + d) * __ldg(table + bj + d);
        acc[0] += go * s;
      }
    }
    for (int d = 0; d < n; ++d)
      if (acc[d] != 0.f) atomicAdd(d_kern + (size_t)p * n + d, acc[d]);
  }
}

// ------------------------------------------------------------------------------------------
// attention core.  qkvr: [B, F, 4*D] = relu projections [Q | K | V | R] of each field row.
// One CTA per batch row; thread = (head, field).
// ------------------------------------------------------------------------------------------
constexpr int kMaxDh = 64;

__global__ void attention_core_fwd_kernel(const float* __restrict__ qkvr, float* __restrict__ Y, int B, int F, int D,
                                          int heads, int use_res) {
  extern __shared__ float sm[];     // [F][4D]
  const int dh = D / heads;
  const float scale = rsqrtf((float)dh);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    const float* src = qkvr + (size_t)b * F * 4 * D;
    for (int e = threadIdx.x; e < F * 4 * D; e += blockDim.x) sm[e] = src[e];
    __syncthreads();
    const int t = threadIdx.x;
    if (t < heads * F) {
      const int h = t / F, i = t - h * F;
      const float* q = sm + (size_t)i * 4 * D + h * dh;
      float m = -INFINITY;
      for (int j = 0; j < F; ++j) {
        const float* k = sm + (size_t)j * 4 * D + D + h * dh;
        float s = 0.f;
        for (int c = 0; c < dh; ++c) s += q[c] * k[c];
        m = fmaxf(m, s * scale);
      }
      float sum = 0.f;
      float acc[kMaxDh];
      for (int c = 0; c < dh; ++c) acc[c] = 0.f;
      for (int j = 0; j < F; ++j) {
        const float* k = sm + (size_t)j * 4 * D + D + h * dh;
        const float* v = sm + (size_t)j * 4 * D + 2 * D + h * dh;
        float s = 0.f;
        for (int c = 0; c < dh; ++c) s += q[c] * k[c];
        const float pj = __expf(s * scale - m);
        sum += pj;
        for (int c = 0; c < dh; ++c) acc[c] += pj * v[c];
      }
      const float inv = 1.f / sum;
      float* y = Y + ((size_t)b * F + i) * D + h * dh;
      const float* res = sm + (size_t)i * 4 * D + 3 * D + h * dh;
      for (int c = 0; c < dh; ++c) {
        float o = acc[c] * inv;
        if (use_res) o += res[c];
        y[c] = fmaxf(o, 0.f);
      }
    }
  }
}

__global__ void attention_core_bwd_kernel(const float* __restrict__ qkvr, const float* __restrict__ Y,
                                          const float* __restrict__ dY, float* __restrict__ d_qkvr, int B, int F,
                                          int D, int heads, int use_res) {
  extern __shared__ float sm[];
  float* blk = sm;                          // [F][4D] inputs
  float* dz = sm + (size_t)F * 4 * D;       // [F][D]   dLoss / d(pre-relu output)
  float* st = dz + (size_t)F * D;           // [heads*F][3]: max, 1/sum, delta
  const int dh = D / heads;
  const float scale = rsqrtf((float)dh);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    const float* src = qkvr + (size_t)b * F * 4 * D;
    for (int e = threadIdx.x; e < F * 4 * D; e += blockDim.x) blk[e] = src[e];
    for (int e = threadIdx.x; e < F * D; e += blockDim.x) {
      const size_t o = (size_t)b * F * D + e;
      dz[e] = Y[o] > 0.f ? dY[o] : 0.f;
    }
    __syncthreads();
    const int t = threadIdx.x;
    float* dst = d_qkvr + (size_t)b * F * 4 * D;
    // phase 1: thread (h, i): softmax statistics, delta = dout . out, dQ row, dRes row
    if (t < heads * F) {
      const int h = t / F, i = t - h * F;
      const float* q = blk + (size_t)i * 4 * D + h * dh;
      const float* dout = dz + (size_t)i * D + h * dh;
      float m = -INFINITY;
      for (int j = 0; j < F; ++j) {
        const float* k = blk + (size_t)j * 4 * D + D + h * dh;
        float s = 0.f;
        for (int c = 0; c < dh; ++c) s += q[c] * k[c];
        m = fmaxf(m, s * scale);
      }
      float sum = 0.f, dsum = 0.f;       // dsum = sum_j p_j (dout . v_j)  (un-normalised)
      for (int j = 0; j < F; ++j) {
        const float* k = blk + (size_t)j * 4 * D + D + h * dh;
        const float* v = blk + (size_t)j * 4 * D + 2 * D + h * dh;
        float s = 0.f, dv = 0.f;
        for (int c = 0; c < dh; ++c) {
          s += q[c] * k[c];
          dv += dout[c] * v[c];
        }
        const float pj = __expf(s * scale - m);
        sum += pj;
        dsum += pj * dv;
      }
      const float inv = 1.f / sum;
      const float delta = dsum * inv;
      st[t * 3 + 0] = m;
      st[t * 3 + 1] = inv;
      st[t * 3 + 2] = delta;
      float dq[kMaxDh];
      for (int c = 0; c < dh; ++c) dq[c] = 0.f;
      for (int j = 0; j < F; ++j) {
        const float* k = blk + (size_t)j * 4 * D + D + h * dh;
        const float* v = blk + (size_t)j * 4 * D + 2 * D + h * dh;
        float s = 0.f, dv = 0.f;
        for (int c = 0; c < dh; ++c) {
          s += q[c] * k[c];
          dv += dout[c] * v[c];
        }
        const float ds = __expf(s * scale - m) * inv * (dv - delta);
        for (int c = 0; c < dh; ++c) dq[c] += ds * k[c];
      }
      float* o = dst + (size_t)i * 4 * D + h * dh;
      for (int c = 0; c < dh; ++c) {
        o[c] = dq[c] * scale;
        o[3 * D + c] = use_res ? dout[c] : 0.f;
      }
    }
    __syncthreads();
    // phase 2: thread (h, j): dK row, dV row
    if (t < heads * F) {
      const int h = t / F, j = t - h * F;
      const float* k = blk + (size_t)j * 4 * D + D + h * dh;
      const float* v = blk + (size_t)j * 4 * D + 2 * D + h * dh;
      float dk[kMaxDh], dvv[kMaxDh];
      for (int c = 0; c < dh; ++c) {
        dk[c] = 0.f;
        dvv[c] = 0.f;
      }
      for (int i = 0; i < F; ++i) {
        const float* q = blk + (size_t)i * 4 * D + h * dh;
        const float* dout = dz + (size_t)i * D + h * dh;
        const float* s3 = st + (h * F + i) * 3;
        float s = 0.f, dv = 0.f;
        for (int c = 0; c < dh; ++c) {
          s += q[c] * k[c];
          dv += dout[c] * v[c];
        }
        const float pij = __expf(s * scale - s3[0]) * s3[1];
        const float ds = pij * (dv - s3[2]);
        for (int c = 0; c < dh; ++c) {
          dk[c] += ds * q[c];
          dvv[c] += pij * dout[c];
        }
      }
      float* o = dst + (size_t)j * 4 * D + h * dh;
      for (int c = 0; c < dh; ++c) {
        o[D + c] = dk[c] * scale;
        o[2 * D + c] = dvv[c];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// The same two kernels with the head width as a template parameter (DH = 1, 2, 4, ... 64): q / dout / k / v rows and
// the per-thread accumulators live in REGISTERS (the generic kernels above index `acc[kMaxDh]` with a run-time bound, i.e.
// from local memory, and every lane re-read its q row from shared memory at a stride of 4D floats = one bank: a 32-way
// conflict on the innermost loop -- 8.6 ms per launch at 65 536 rows x 26 fields x D = 32, 2 % of the HBM rate).  Shared
// rows are padded by 4 floats; key / value reads are warp broadcasts.
// ------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(256) attention_core_fwd_t_kernel(const float* __restrict__ qkvr, float* __restrict__ Y, int B, int F, int D,
                                                                    int heads, int use_res) {
  extern __shared__ float sm[];     // [F][4D + 4]
  const int RS = 4 * D + 4;
  const float scale = rsqrtf((float)DH);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    const float4* src = reinterpret_cast<const float4*>(qkvr + (size_t)b * F * 4 * D);
    for (int e = threadIdx.x; e < F * D; e += blockDim.x) {           // D float4 per field row
      const int f = e / D, c4 = e - f * D;
      *reinterpret_cast<float4*>(sm + (size_t)f * RS + 4 * c4) = __ldg(src + e);
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t < heads * F) {
      const int h = t / F, i = t - h * F;
      float q[DH], acc[DH];
#pragma unroll
      for (int c = 0; c < DH; ++c) {
        q[c] = sm[(size_t)i * RS + h * DH + c] * scale;
        acc[c] = 0.f;
      }
      float m = -INFINITY;
      for (int j = 0; j < F; ++j) {
        const float* k = sm + (size_t)j * RS + D + h * DH;
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < DH; ++c) s = fmaf(q[c], k[c], s);
        m = fmaxf(m, s);
      }
      float sum = 0.f;
      for (int j = 0; j < F; ++j) {
        const float* k = sm + (size_t)j * RS + D + h * DH;
        const float* v = k + D;
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < DH; ++c) s = fmaf(q[c], k[c], s);
        const float pj = __expf(s - m);
        sum += pj;
#pragma unroll
        for (int c = 0; c < DH; ++c) acc[c] = fmaf(pj, v[c], acc[c]);
      }
      const float inv = 1.f / sum;
      float* y = Y + ((size_t)b * F + i) * D + h * DH;
      const float* res = sm + (size_t)i * RS + 3 * D + h * DH;
#pragma unroll
      for (int c = 0; c < DH; ++c) {
        float o = acc[c] * inv;
        if (use_res) o += res[c];
        y[c] = fmaxf(o, 0.f);
      }
    }
  }
}

template <int DH>
__global__ void __launch_bounds__(256) attention_core_bwd_t_kernel(const float* __restrict__ qkvr, const float* __restrict__ Y,
                                                                    const float* __restrict__ dY, float* __restrict__ d_qkvr, int B,
                                                                    int F, int D, int heads, int use_res) {
  extern __shared__ float sm[];
  const int RS = 4 * D + 4, RZ = D + 4;
  float* blk = sm;                          // [F][4D + 4] inputs
  float* dz = sm + (size_t)F * RS;          // [F][D + 4]  dLoss / d(pre-relu output)
  float* st = dz + (size_t)F * RZ;          // [heads*F][3]: max, 1/sum, delta
  const float scale = rsqrtf((float)DH);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    const float4* src = reinterpret_cast<const float4*>(qkvr + (size_t)b * F * 4 * D);
    for (int e = threadIdx.x; e < F * D; e += blockDim.x) {
      const int f = e / D, c4 = e - f * D;
      *reinterpret_cast<float4*>(blk + (size_t)f * RS + 4 * c4) = __ldg(src + e);
    }
    for (int e = threadIdx.x; e < F * D; e += blockDim.x) {
      const size_t o = (size_t)b * F * D + e;
      const int f = e / D, c = e - f * D;
      dz[(size_t)f * RZ + c] = Y[o] > 0.f ? dY[o] : 0.f;
    }
    __syncthreads();
    const int t = threadIdx.x;
    float* dst = d_qkvr + (size_t)b * F * 4 * D;
    // phase 1: thread (h, i): softmax statistics, delta = dout . out, dQ row, dRes row
    if (t < heads * F) {
      const int h = t / F, i = t - h * F;
      float q[DH], dout[DH], dq[DH];
#pragma unroll
      for (int c = 0; c < DH; ++c) {
        q[c] = blk[(size_t)i * RS + h * DH + c] * scale;
        dout[c] = dz[(size_t)i * RZ + h * DH + c];
        dq[c] = 0.f;
      }
      float m = -INFINITY;
      for (int j = 0; j < F; ++j) {
        const float* k = blk + (size_t)j * RS + D + h * DH;
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < DH; ++c) s = fmaf(q[c], k[c], s);
        m = fmaxf(m, s);
      }
      float sum = 0.f, dsum = 0.f;       // dsum = sum_j p_j (dout . v_j)  (un-normalised)
      for (int j = 0; j < F; ++j) {
        const float* k = blk + (size_t)j * RS + D + h * DH;
        const float* v = k + D;
        float s = 0.f, dv = 0.f;
#pragma unroll
        for (int c = 0; c < DH; ++c) {
          s = fmaf(q[c], k[c], s);
          dv = fmaf(dout[c], v[c], dv);
        }
        const float pj = __expf(s - m);
        sum += pj;
        dsum = fmaf(pj, dv, dsum);
      }
      const float inv = 1.f / sum;
      const float delta = dsum * inv;
      st[t * 3 + 0] = m;
      st[t * 3 + 1] = inv;
      st[t * 3 + 2] = delta;
      for (int j = 0; j < F; ++j) {
        const float* k = blk + (size_t)j * RS + D + h * DH;
        const float* v = k + D;
        float s = 0.f, dv = 0.f;
#pragma unroll
        for (int c = 0; c < DH; ++c) {
          s = fmaf(q[c], k[c], s);
          dv = fmaf(dout[c], v[c], dv);
        }
        const float ds = __expf(s - m) * inv * (dv - delta);
#pragma unroll
        for (int c = 0; c < DH; ++c) dq[c] = fmaf(ds, k[c], dq[c]);
      }
      float* o = dst + (size_t)i * 4 * D + h * DH;
#pragma unroll
      for (int c = 0; c < DH; ++c) {
        o[c] = dq[c] * scale;
        o[3 * D + c] = use_res ? dout[c] : 0.f;
      }
    }
    __syncthreads();
    // phase 2: thread (h, j): dK row, dV row
    if (t < heads * F) {
      const int h = t / F, j = t - h * F;
      float k[DH], v[DH], dk[DH], dvv[DH];
#pragma unroll
      for (int c = 0; c < DH; ++c) {
        k[c] = blk[(size_t)j * RS + D + h * DH + c];
        v[c] = blk[(size_t)j * RS + 2 * D + h * DH + c];
        dk[c] = 0.f;
        dvv[c] = 0.f;
      }
      for (int i = 0; i < F; ++i) {
        const float* q = blk + (size_t)i * RS + h * DH;
        const float* dout = dz + (size_t)i * RZ + h * DH;
        const float* s3 = st + (h * F + i) * 3;
        float s = 0.f, dv = 0.f;
#pragma unroll
        for (int c = 0; c < DH; ++c) {
          s = fmaf(q[c], k[c], s);
          dv = fmaf(dout[c], v[c], dv);
        }
        const float pij = __expf(s * scale - s3[0]) * s3[1];
        const float ds = pij * (dv - s3[2]);
#pragma unroll
        for (int c = 0; c < DH; ++c) {
          dk[c] = fmaf(ds, q[c], dk[c]);
          dvv[c] = fmaf(pij, dout[c], dvv[c]);
        }
      }
      float* o = dst + (size_t)j * 4 * D + h * DH;
#pragma unroll
      for (int c = 0; c < DH; ++c) {
        o[D + c] = dk[c] * scale;
        o[2 * D + c] = dvv[c];
      }
    }
  }
}

}  // namespace dtb

using namespace dtb;

extern "C" {

int dtb_pnn_fwd(const int32_t* idx, const float* table, const int64_t* row_offsets, const float* op_kernel,
                float* ip, float* op, int B, int F, int D, int kernel_type, int* status, void* stream) {
  DTB_CHECK_ARG(idx && table && row_offsets, "NULL argument");
  DTB_CHECK_ARG(F >= 2 && D >= 1 && D <= kMaxD && B >= 0, "need F >= 2, 1 <= D <= 64");
  DTB_CHECK_ARG(kernel_type >= 0 && kernel_type <= 2, "kernel_type must be 0 (mat), 1 (vec) or 2 (num)");
  DTB_CHECK_ARG(!op || op_kernel, "outer product requested without a kernel");
  if (B == 0 || (!ip && !op)) return DTB_OK;
  const int P = F * (F - 1) / 2;
  if (P > 1024) {
    set_error("dtb_pnn_fwd: %d pairs exceed one CTA", P);
    return DTB_ERR_UNSUPPORTED;
  }
  const int threads = (P + 31) / 32 * 32;
  const size_t smem = (size_t)kPnnRows * F * D * sizeof(float);
  DTB_CUDA_OK(cudaFuncSetAttribute(pnn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pnn_fwd_kernel<<<ceil_div(B, kPnnRows), threads, smem, (cudaStream_t)stream>>>(idx, table, row_offsets, op_kernel,
                                                                                 ip, op, B, F, D, P, kernel_type, status);
  DTB_LAUNCH_OK();
  return DTB_OK;
}

int dtb_pnn_bwd(const int32_t* idx, const float* table, const int64_t* row_offsets, const float* op_kernel,
                const float* d_ip, const float* d_op, float* grad_table, float* d_op_kernel, int B, int F, int D,
                int kernel_type, void* stream) {
  DTB_CHECK_ARG(idx && table && row_offsets && grad_table, "NULL argument");
  DTB_CHECK_ARG(F >= 2 && D >= 1 && D <= kMaxD && B >= 0, "need F >= 2, 1 <= D <= 64");
  DTB_CHECK_ARG(!d_op || (op_kernel && d_op_kernel), "outer product gradient without kernel buffers");
  if (B == 0 || (!d_ip && !d_op)) return DTB_OK;
  const int P = F * (F - 1) / 2;
  const int threads = (P + 31) / 32 * 32;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)2 * kPnnRows * F * D * sizeof(float);
  DTB_CUDA_OK(cudaFuncSetAttribute(pnn_bwd_de_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pnn_bwd_de_kernel<<<ceil_div(B, kPnnRows), threads, smem, st>>>(idx, table, row_offsets, op_kernel, d_ip, d_op,
                                                                  grad_table, B, F, D, P, kernel_type);
  DTB_LAUNCH_OK();
  if (d_op) {
    int ctas = sm_count() * 2;
    if (ctas > B) ctas = B;
    const int rows_per_cta = ceil_div(B, ctas);
    ctas = ceil_div(B, rows_per_cta);
    pnn_bwd_dk_kernel<<<ctas, threads, 0, st>>>(idx, table, row_offsets, d_op, d_op_kernel, B, F, D, P, kernel_type,
                                                rows_per_cta);
    DTB_LAUNCH_OK();
  }
  return DTB_OK;
}

int dtb_attention_core_fwd(const float* qkvr, float* Y, int B, int F, int D, int heads, int use_residual,
                           void* stream) {
  DTB_CHECK_ARG(qkvr && Y, "NULL argument");
  DTB_CHECK_ARG(B >= 0 && F >= 1 && D >= 1 && heads >= 1 && D % heads == 0, "bad shape (num_heads must divide D)");
  DTB_CHECK_ARG(D / heads <= kMaxDh && heads * F <= 1024, "head width <= 64 and heads*fields <= 1024");
  if (B == 0) return DTB_OK;
  const int threads = (heads * F + 31) / 32 * 32;
  int grid = sm_count() * 8;
  if (grid > B) grid = B;
  const int dh = D / heads;
  const int nthr = threads < 64 ? 64 : threads;
  if (D % 4 == 0 && nthr <= 256 && (reinterpret_cast<uintptr_t>(qkvr) & 15) == 0) {
    const size_t smem_t = (size_t)F * (4 * D + 4) * sizeof(float);
#define DTB_ATT_FWD(DHV)                                                                                                   \
  case DHV:                                                                                                                \
    DTB_CUDA_OK(cudaFuncSetAttribute(attention_core_fwd_t_kernel<DHV>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                     (int)smem_t));                                                                         \
    attention_core_fwd_t_kernel<DHV><<<grid, nthr, smem_t, (cudaStream_t)stream>>>(qkvr, Y, B, F, D, heads, use_residual); \
    DTB_LAUNCH_OK();                                                                                                       \
    return DTB_OK;
    switch (dh) {
      DTB_ATT_FWD(1) DTB_ATT_FWD(2) DTB_ATT_FWD(4) DTB_ATT_FWD(8) DTB_ATT_FWD(16) DTB_ATT_FWD(32) DTB_ATT_FWD(64)
      default: break;
    }
#undef DTB_ATT_FWD
  }
  const size_t smem = (size_t)F * 4 * D * sizeof(float);
  DTB_CUDA_OK(cudaFuncSetAttribute(attention_core_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_core_fwd_kernel<<<grid, nthr, smem, (cudaStream_t)stream>>>(qkvr, Y, B, F, D, heads, use_residual);
  DTB_LAUNCH_OK();
  return DTB_OK;
}

int dtb_attention_core_bwd(const float* qkvr, const float* Y, const float* dY, float* d_qkvr, int B, int F, int D,
                           int heads, int use_residual, void* stream) {
  DTB_CHECK_ARG(qkvr && Y && dY && d_qkvr, "NULL argument");
  DTB_CHECK_ARG(B >= 0 && F >= 1 && D >= 1 && heads >= 1 && D % heads == 0, "bad shape (num_heads must divide D)");
  DTB_CHECK_ARG(D / heads <= kMaxDh && heads * F <= 1024, "head width <= 64 and heads*fields <= 1024");
  if (B == 0) return DTB_OK;
  const int threads = (heads * F + 31) / 32 * 32;
  int grid = sm_count() * 8;
  if (grid > B) grid = B;
  const int dh = D / heads;
  const int nthr = threads < 64 ? 64 : threads;
  if (D % 4 == 0 && nthr <= 256 && (reinterpret_cast<uintptr_t>(qkvr) & 15) == 0) {
    const size_t smem_t = ((size_t)F * (4 * D + 4) + (size_t)F * (D + 4) + (size_t)heads * F * 3) * sizeof(float);
#define DTB_ATT_BWD(DHV)                                                                                                \
  case DHV:                                                                                                             \
    DTB_CUDA_OK(cudaFuncSetAttribute(attention_core_bwd_t_kernel<DHV>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                     (int)smem_t));                                                                      \
    attention_core_bwd_t_kernel<DHV><<<grid, nthr, smem_t, (cudaStream_t)stream>>>(qkvr, Y, dY, d_qkvr, B, F, D, heads, \
                                                                                   use_residual);                       \
    DTB_LAUNCH_OK();                                                                                                    \
    return DTB_OK;
    switch (dh) {
      DTB_ATT_BWD(1) DTB_ATT_BWD(2) DTB_ATT_BWD(4) DTB_ATT_BWD(8) DTB_ATT_BWD(16) DTB_ATT_BWD(32) DTB_ATT_BWD(64)
      default: break;
    }
#undef DTB_ATT_BWD
  }
  const size_t smem = ((size_t)F * 5 * D + (size_t)heads * F * 3) * sizeof(float);
  DTB_CUDA_OK(cudaFuncSetAttribute(attention_core_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_core_bwd_kernel<<<grid, nthr, smem, (cudaStream_t)stream>>>(qkvr, Y, dY, d_qkvr, B, F, D, heads, use_residual);
  DTB_LAUNCH_OK();
  return DTB_OK;
}

}  // extern "C"
