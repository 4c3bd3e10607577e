// Micro-probe: where does grad_rows_pack spend its time?  (development tool, not part of the library)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__global__ void gen_idx(int* idx, int n, int vocab, uint32_t seed) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = mix(i * 2654435761u + seed) % vocab;
}

template <int V>
__global__ void pack_variant(const int* __restrict__ idx, int64_t vocab, float* __restrict__ grad,
                             int* __restrict__ claim, float* __restrict__ packed, int step, int B, int F) {
  const int D = 16, Q = 4;
  const int64_t total = (int64_t)B * F * Q;
  const int lane = threadIdx.x & 31;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ref = i / Q;
    const int q = (int)(i - ref * Q);
    const int f = (int)(ref % F);
    const int64_t row = (int64_t)f * vocab + idx[ref];
    bool own = true;
    if (V == 0 || V == 1) {          // returning atomicMax
      int old = 0x7fffffff;
      if (q == 0) old = atomicMax(claim + row, step);
      old = __shfl_sync(0xffffffffu, old, lane - q);
      own = old < step;
    } else if (V == 4) {             // phase B of the two-phase claim
      own = claim[row] == (int)ref;
    } else if (V == 5) {             // plain load of the claim word + compare (no atomic)
      own = claim[row] < step;
    }
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (V == 6) {                    // V0 with the zero store ordered behind the dependent packed store
      int old = 0x7fffffff;
      if (q == 0) old = atomicMax(claim + row, step);
      old = __shfl_sync(0xffffffffu, old, lane - q);
      own = old < step;
      float4* src = reinterpret_cast<float4*>(grad + row * D + (q << 2));
      if (own) v = *src;
      *reinterpret_cast<float4*>(packed + ref * D + (q << 2)) = v;
      asm volatile("" ::: "memory");
      if (own) *src = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    if (V != 1 && own) {
      float4* src = reinterpret_cast<float4*>(grad + row * D + (q << 2));
      v = *src;
      *src = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    *reinterpret_cast<float4*>(packed + ref * D + (q << 2)) = v;
  }
}
__global__ void claim_store(const int* __restrict__ idx, int64_t vocab, int* __restrict__ claim, int B, int F) {
  const int64_t total = (int64_t)B * F;
  for (int64_t ref = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ref < total; ref += (int64_t)gridDim.x * blockDim.x)
    claim[(int64_t)(ref % F) * vocab + idx[ref]] = (int)ref;
}

int main() {
  const int B = 65536, F = 26, D = 16, vocab = 1000000;
  int *idx, *claim; float *grad, *packed;
  cudaMalloc(&idx, (size_t)B * F * 4);
  cudaMalloc(&claim, (size_t)F * vocab * 4);
  cudaMalloc(&grad, (size_t)F * vocab * D * 4);
  cudaMalloc(&packed, (size_t)B * F * D * 4);
  cudaMemset(claim, 0, (size_t)F * vocab * 4);
  cudaMemset(grad, 0, (size_t)F * vocab * D * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * 16;
  int step = 0; uint32_t seed = 1;
  auto fresh = [&]() { gen_idx<<<(B * F + 255) / 256, 256>>>(idx, B * F, vocab, seed++); cudaDeviceSynchronize(); };
  for (int rep = 0; rep < 3; ++rep) {
    float ms[8] = {0};
    for (int v = 0; v < 7; ++v) {
      fresh(); ++step;
      cudaEventRecord(e0);
      switch (v) {
        case 0: pack_variant<0><<<grid, 256>>>(idx, vocab, grad, claim, packed, step, B, F); break;
        case 1: pack_variant<1><<<grid, 256>>>(idx, vocab, grad, claim, packed, step, B, F); break;
        case 2: pack_variant<2><<<grid, 256>>>(idx, vocab, grad, claim, packed, step, B, F); break;
        case 3: claim_store<<<grid, 256>>>(idx, vocab, claim, B, F); break;
        case 4: claim_store<<<grid, 256>>>(idx, vocab, claim, B, F);
                pack_variant<4><<<grid, 256>>>(idx, vocab, grad, claim, packed, step, B, F); break;
        case 6: pack_variant<6><<<grid, 256>>>(idx, vocab, grad, claim, packed, step, B, F); break;
        case 5: pack_variant<5><<<grid, 256>>>(idx, vocab, grad, claim, packed, step, B, F); break;
      }
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms[v], e0, e1);
      if (v == 3 || v == 4) cudaMemset(claim, 0, (size_t)F * vocab * 4);
    }
    printf("rep %d: V0 current=%.0fus  V1 atomic-only(miss)=%.0fus  V2 grad-rmw-only=%.0fus  V3 claim-store=%.0fus  "
           "V4 two-phase total=%.0fus  V5 plain-claim-load+rmw=%.0fus  V6 ordered-zero=%.0fus\n", rep, ms[0] * 1e3, ms[1] * 1e3, ms[2] * 1e3,
           ms[3] * 1e3, ms[4] * 1e3, ms[5] * 1e3, ms[6] * 1e3);
  }
  cudaError_t err = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(err));
  return 0;
}
