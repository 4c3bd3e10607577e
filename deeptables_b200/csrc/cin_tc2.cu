// CIN on tcgen05, second organisation: ONE tensor pass on power-of-two-scaled fp16 operands (DTB_CIN_TC_F16X1) with
// TWO threads per GEMM row.
//
// Why a second organisation.  With a single tensor pass the MMA work of cin_tc_fwd_kernel<D, true> drops to a third,
// but its time only fell from 2.54 to 1.90 ms (B200, 65 536 rows): the kernel was never waiting for the tensor pipe any
// more, it was waiting for ITSELF -- 130 pipeline granules per super tile, each a serial chain
// (products -> wait empty -> tcgen05.st -> wait::st -> arrive -> MMA issuer: wait tile 0, issue, wait tile 1, issue, 2 commits)
// of roughly a thousand cycles for 256 cycles of tensor work.  Here:
//   * a granule is one x0 field x ALL hidden fields (K = Hp, up to 64): 78 granules per super tile instead of 130, and
//     twice the tensor work behind every handshake;
//   * both M = 128 tiles of the CTA share one "operand ready" barrier per stage (16 warp arrivals): the issuing warp
//     waits once per granule, issues 2 x Hp/16 MMAs, commits twice;
//   * each GEMM row m = (batch row, dim) is owned by a PAIR of threads, each holding half of h_k[b, :, d] in registers
//     (32 values instead of 64): 16 producer warps instead of 8 hide the tcgen05.st / mbarrier latencies of one
//     another, each writes half of the operand columns, and the accumulator read-out of a layer is split between the
//     two (halves the epilogue, which nothing overlaps because the next layer's operand depends on it);
//   * the fp16 operand of a granule is 32 TMEM columns per tile: 4 stages x 2 tiles = 256 columns next to the two
//     128-column accumulators (TMEM 100 % allocated, as before).
// Arithmetic, saved-activation format and the weight images are exactly those of cin_tc_fwd_kernel<D, true>
// (cin_tc.cu): per-row scale 2^e chosen from max|x0 row| * max|h row| (the two halves exchange their maxima through
// shared memory), per-layer weight scale from max|W_k|, both undone on the fp32 accumulator.
#include "cin_tc_common.cuh"
#include <cuda_fp16.h>

namespace dtb {

constexpr int kT2Threads = 576;        // warps 0-15 producer + epilogue, 16 MMA issue + TMEM owner, 17 weight loader
constexpr int kT2StagesA = 4;
constexpr int kT2StagesB = 6;
constexpr int kT2ACols = 32;           // TMEM columns of one (stage, tile) operand block: fp16 [128 x 64]

struct T2Smem {
  int b_off, x0_off, mx_off, bar_off, total;
};
__host__ __device__ inline T2Smem tc2_layout(int b_stage_bytes, int F) {
  T2Smem l;
  l.b_off = 0;
  l.x0_off = kT2StagesB * b_stage_bytes;
  l.mx_off = l.x0_off + 2 * 128 * F * 4;          // x0s[tile][r][i][d]
  l.bar_off = l.mx_off + 2 * 2 * 2 * 128 * 4;     // row maxima [tile][parity][half][t]
  l.bar_off = (l.bar_off + 15) / 16 * 16;
  l.total = l.bar_off + 256;
  return l;
}

template <int D>
__global__ void __launch_bounds__(kT2Threads, 1) cin_tc2_fwd_kernel(const __grid_constant__ CinTcParams p) {
  constexpr int R = 128 / D;                 // batch rows per M=128 tile
  extern __shared__ __align__(1024) uint8_t smem[];
  const T2Smem lay = tc2_layout(p.b_stage_bytes, p.F);
  uint8_t* smem_b = smem + lay.b_off;
  float* x0s = reinterpret_cast<float*>(smem + lay.x0_off);
  float* mxs = reinterpret_cast<float*>(smem + lay.mx_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bar_off);
  uint64_t* full_a = bars;                        // [stage] 16 producer warps (both tiles)
  uint64_t* empty_a = bars + 4;                   // [stage] commit
  uint64_t* full_b = bars + 8;                    // [stage] bulk copy (tx)
  uint64_t* empty_b = bars + 14;                  // [stage] commit
  uint64_t* acc_full = bars + 20;                 // commit after the last granule of a layer (both tiles)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int F = p.F;
  const int n_super = (p.B + 2 * R - 1) / (2 * R);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kT2StagesA; ++s) {
      tc::mbar_init(&full_a[s], 16);
      tc::mbar_init(&empty_a[s], 1);
    }
    for (int s = 0; s < kT2StagesB; ++s) {
      tc::mbar_init(&full_b[s], 1);
      tc::mbar_init(&empty_b[s], 1);
    }
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 16) tc::tmem_alloc(tmem_slot, kTmemCols);
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 16) {
    // =================== producer + epilogue: tile g, half q of GEMM row t ============================
    const int g = warp >> 3, q = (warp >> 2) & 1;
    const int t = (warp & 3) * 32 + lane;
    const int tt = q * 128 + t;                     // index inside the tile's 256-thread group
    const int r = t / D, d = t % D;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    float* x0g = x0s + (size_t)g * 128 * F;         // [r][i][d]
    float* mxg = mxs + (size_t)g * 2 * 2 * 128;     // [parity][half][t]
    uint32_t gran = 0, layer_cnt = 0;
    float h[32];
    for (int st = blockIdx.x; st < n_super; st += gridDim.x) {
      const int row0 = (st * 2 + g) * R;
      const int b = row0 + r;
      // ---- gather this tile's x0 block: R rows x F fields x D floats, 16-byte pieces, 256 threads ------------
      {
        constexpr int Q = D / 4;
        for (int e = tt; e < R * F * Q; e += 256) {
          const int rr = e / (F * Q);
          const int rem = e - rr * F * Q;
          const int i = rem / Q, qq = rem - i * Q;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row0 + rr < p.B) {
            const int64_t rb = table_row(p.row_offsets, i, __ldg(p.idx + (int64_t)(row0 + rr) * F + i), D, p.status);
            if (rb >= 0) v = ldg_stream_f4(p.table + rb + (qq << 2));
          }
          *reinterpret_cast<float4*>(x0g + ((size_t)rr * F + i) * D + (qq << 2)) = v;
        }
      }
      asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
      // ---- h_0 = this thread's half of the x0 row (zero padded to Hp[0]) ; max|x0 row| --------------------------
      {
        const int nh0 = p.Hp[0] >> 1;
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const int j = q * nh0 + jj;
          h[jj] = (jj < nh0 && j < F) ? x0g[((size_t)r * F + j) * D + d] : 0.f;
        }
      }
      float xmax = 0.f;
      for (int i = 0; i < F; ++i) xmax = fmaxf(xmax, fabsf(x0g[((size_t)r * F + i) * D + d]));
      float hmax = xmax;                               // max|h_k row| over BOTH halves
      if (p.saved) {
        // block-transposed copy for the wgrad kernel: [m / 64][field][68]; the two halves write alternate fields
        const size_t m_pad = (size_t)(st * 2 + g) * 128 + t;
        float* xb = p.saved + p.xb_off + (m_pad >> 6) * (size_t)(F * kWgPad) + (m_pad & 63);
        for (int j = q; j < F; j += 2) xb[j * kWgPad] = x0g[((size_t)r * F + j) * D + d];
      }
      for (int k = 0; k < p.n_layers; ++k) {
        const int Hp = p.Hp[k], L = p.L[k];
        const int nh = Hp >> 1;                        // operand values of this thread per granule: 16 or 32
        float srow, inv_row, sw, inv_w;
        tc::pow2_scale_to_1024(xmax * hmax, srow, inv_row);
        tc::pow2_scale_to_1024(__int_as_float(__ldg(p.wmax + k)), sw, inv_w);
        const float inv_acc = inv_row * inv_w;
        for (int i = 0; i < F; ++i) {
          const float xi = x0g[((size_t)r * F + i) * D + d] * srow;
          const uint32_t sa = gran % kT2StagesA, pa = (gran / kT2StagesA) & 1;
          ++gran;
          uint32_t zh[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) zh[c] = tc::pack_f16x2(xi * h[2 * c], xi * h[2 * c + 1]);
          tc::mbar_wait(&empty_a[sa], pa ^ 1);
          tc::fence_after_thread_sync();
          const uint32_t a_col = tmem_base + lane_base + 2 * kAccCols + (sa * 2 + g) * kT2ACols + q * (nh >> 1);
          tc::tmem_st8v(a_col, zh[0], zh[1], zh[2], zh[3], zh[4], zh[5], zh[6], zh[7]);
          if (nh == 32) tc::tmem_st8v(a_col + 8, zh[8], zh[9], zh[10], zh[11], zh[12], zh[13], zh[14], zh[15]);
          tc::tmem_wait_st();
          tc::fence_before_thread_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&full_a[sa]);
        }
        // ---- epilogue of layer k: this thread's share of its accumulator row ------------------------------------
        tc::mbar_wait(acc_full, layer_cnt & 1);
        ++layer_cnt;
        tc::fence_after_thread_sync();
        const int hid_n = p.hid_n[k], pool_lo = p.pool_lo[k], pool_n = p.pool_n[k];
        const int nhn = (k + 1 < p.n_layers) ? (p.Hp[k + 1] >> 1) : 0;      // next layer: operand values per thread
        const float* bias = p.bias ? p.bias + p.bias_off[k] : nullptr;
        uint16_t* mrow = (p.saved && p.act == DTB_ACT_RELU && b < p.B)
                             ? reinterpret_cast<uint16_t*>(reinterpret_cast<uint32_t*>(p.saved + p.saved_off[k]) +
                                                           ((size_t)b * D + d) * ((L + 31) >> 5))
                             : nullptr;
        float* hb = nullptr;
        if (p.saved && hid_n > 0) {
          const size_t m_pad = (size_t)(st * 2 + g) * 128 + t;
          hb = p.saved + p.hb_off[k] + (m_pad >> 6) * (size_t)(hid_n * kWgPad) + (m_pad & 63);
        }
        // every granule of this layer has been handed over: h is dead and becomes the next layer's operand in place
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) h[jj] = 0.f;
        // block ownership.  Hidden 16-column blocks (columns < hid_n = the next layer's h): the owner is the half whose
        // operand range [q*nhn, (q+1)*nhn) holds them -- slots 0 and 1 of this thread.  Pooled-only blocks (direct=False:
        // columns [pool_lo, pool_lo + pool_n)): first half of them to q = 0, the rest to q = 1.
        const int first_pb = (pool_lo >= hid_n) ? (pool_lo >> 4) : (hid_n >> 4);     // first block that is pooled but not hidden
        const int last_pb = (pool_lo + pool_n) >> 4;
        const int n_pb = last_pb > first_pb ? last_pb - first_pb : 0;
        const int pb_split = (n_pb + 1) >> 1;
#pragma unroll
        for (int slot = 0; slot < 6; ++slot) {
          // slots 0-1: hidden blocks of this half; slots 2-5: pooled-only blocks of this half
          int cb;
          bool live;
          if (slot < 2) {
            cb = ((q * nhn) >> 4) + slot;
            live = (slot * 16 < nhn) && (cb * 16 < hid_n);
          } else {
            const int s2 = slot - 2;
            cb = first_pb + q * pb_split + s2;
            live = s2 < (q == 0 ? pb_split : n_pb - pb_split);
          }
          if (live) {                                   // warp-uniform
            uint32_t v[16];
            tc::tmem_ld16(tmem_base + lane_base + g * kAccCols + cb * 16, v);
            tc::tmem_wait_ld();
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float val = __uint_as_float(v[j]) * inv_acc;
              if (bias) val += __ldg(bias + cb * 16 + j);
              if (p.act == DTB_ACT_RELU) val = fmaxf(val, 0.f);
              o[j] = val;
            }
            if (slot < 2) {
#pragma unroll
              for (int j = 0; j < 16; ++j) h[(slot & 1) * 16 + j] = o[j];
              if (hb) {
#pragma unroll
                for (int j = 0; j < 16; ++j) hb[(cb * 16 + j) * kWgPad] = o[j];
              }
            }
            if (mrow) {
              uint32_t bits = 0u;
#pragma unroll
              for (int j = 0; j < 16; ++j) bits |= (o[j] > 0.f ? 1u : 0u) << j;
              mrow[cb] = (uint16_t)bits;
            }
            // sum over the D lanes that share a batch row (reduce-scatter butterfly, see cin_tc_fwd_kernel)
            if constexpr (D <= 16) {
#pragma unroll
              for (int blk = 0; blk < 16 / D; ++blk) {
                const int col0 = cb * 16 + blk * D;
                if (col0 + D > pool_lo && col0 < pool_lo + pool_n) {     // warp-uniform
                  float w[D];
#pragma unroll
                  for (int j = 0; j < D; ++j) w[j] = o[blk * D + j];
#pragma unroll
                  for (int off = D / 2; off >= 1; off >>= 1) {
                    const bool up = (d & off) != 0;
#pragma unroll
                    for (int j = 0; j < off; ++j) {
                      const float send = up ? w[j] : w[j + off];
                      const float keep = up ? w[j + off] : w[j];
                      w[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                    }
                  }
                  const int col = col0 + d;
                  if (b < p.B && col >= pool_lo && col < pool_lo + pool_n)
                    p.pooled[(size_t)b * p.P + p.pcol0[k] + (col - pool_lo)] = w[0];
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int col = cb * 16 + j;
                if (col >= pool_lo && col < pool_lo + pool_n) {      // warp-uniform
                  float sum = o[j];
#pragma unroll
                  for (int off = 1; off < 32; off <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
                  if (d == 0 && b < p.B) p.pooled[(size_t)b * p.P + p.pcol0[k] + (col - pool_lo)] = sum;
                }
              }
            }
          }
        }
        tc::fence_before_thread_sync();
        // ---- next layer's operand: own half of h_{k+1}; row maximum over both halves -----------------------------
        if (k + 1 < p.n_layers) {
          float own = 0.f;
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) own = fmaxf(own, fabsf(h[jj]));
          float* mx = mxg + (size_t)(k & 1) * 2 * 128;
          mx[q * 128 + t] = own;
          asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
          hmax = fmaxf(mx[t], mx[128 + t]);
        }
      }
      asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");   // x0 block free for the next super tile
    }
  } else if (warp == 16) {
    // ================================ MMA issuer ===============================================
    const bool leader = elect_one_sync();
    const uint32_t smem_b_u32 = tc::smem_u32(smem_b);
    uint32_t gran = 0, chunk = 0;
    for (int st = blockIdx.x; st < n_super; st += gridDim.x) {
      for (int k = 0; k < p.n_layers; ++k) {
        const int Hp = p.Hp[k], L = p.L[k];
        const uint32_t idesc = tc::make_idesc_f16(128, (uint32_t)L);
        const uint32_t lbo_b = (uint32_t)(L >> 3) * 128;       // K-direction core stride of the W image
        const uint64_t desc_hi = ((uint64_t)((lbo_b >> 4) & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
        for (int i = 0; i < F; ++i, ++chunk, ++gran) {
          const uint32_t sb = chunk % kT2StagesB, pb = (chunk / kT2StagesB) & 1;
          const uint32_t sa = gran % kT2StagesA, pa = (gran / kT2StagesA) & 1;
          tc::mbar_wait(&full_b[sb], pb);
          tc::mbar_wait(&full_a[sa], pa);
          tc::fence_after_thread_sync();
          if (leader) {
            const uint32_t b_addr = smem_b_u32 + sb * (uint32_t)p.b_stage_bytes;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const uint32_t d_tmem = tmem_base + g * kAccCols;
              const uint32_t a_base = tmem_base + 2 * kAccCols + (sa * 2 + g) * kT2ACols;
#pragma unroll
              for (int ks = 0; ks < kMaxHp / 16; ++ks) {
                if (ks * 16 < Hp) {
                  const uint64_t desc_b = desc_hi | (uint64_t)(((b_addr + ks * 2 * lbo_b) >> 4) & 0x3FFF);
                  tc::mma_ts(d_tmem, a_base + ks * 8, desc_b, idesc, (uint32_t)((i | ks) != 0));
                }
              }
            }
            tc::mma_commit(&empty_a[sa]);
            tc::mma_commit(&empty_b[sb]);
            if (i == F - 1) tc::mma_commit(acc_full);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================================ weight loader ============================================
    if (lane == 0) {
      uint32_t chunk = 0;
      for (int st = blockIdx.x; st < n_super; st += gridDim.x) {
        for (int k = 0; k < p.n_layers; ++k) {
          const uint32_t bytes = (uint32_t)p.L[k] * p.Hp[k] * 2;       // one fp16 image
          const uint32_t stride = (uint32_t)p.L[k] * p.Hp[k] * 4;      // the pack keeps room for a lo image
          const uint8_t* src = p.wpack + p.wpack_off[k];
          for (int i = 0; i < F; ++i, ++chunk) {
            const uint32_t sb = chunk % kT2StagesB, pb = (chunk / kT2StagesB) & 1;
            tc::mbar_wait(&empty_b[sb], pb ^ 1);
            tc::mbar_arrive_expect_tx(&full_b[sb], bytes);
            tc::bulk_g2s(smem_b + (size_t)sb * p.b_stage_bytes, src + (size_t)i * stride, bytes, &full_b[sb]);
          }
        }
      }
    }
    __syncwarp();
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 16) {
    tc::fence_after_thread_sync();
    tc::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int tc2_b_stage(const CinTcParams& p) {
  int b = 0;
  for (int k = 0; k < p.n_layers; ++k) {
    const int bytes = p.L[k] * p.Hp[k] * 2;
    if (bytes > b) b = bytes;
  }
  return b;
}

bool cin_tc2_fwd_supported(const CinTcParams& p, int D) {
  if (D != 16 && D != 32) return false;
  if (p.n_pass != 1) return false;
  if (p.saved && !p.compact) return false;             // the full (fp32 T_k) saved format is written by cin_tc_fwd_kernel only
  for (int k = 0; k < p.n_layers; ++k) {
    if (p.Hp[k] != 32 && p.Hp[k] != 64) return false;
    if (p.L[k] % 16 || p.L[k] > kMaxL) return false;
    if (p.hid_n[k] % 16 || p.pool_lo[k] % 16 || p.pool_n[k] % 16) return false;
    if (p.hid_n[k] > 0 && p.pool_lo[k] != 0 && p.pool_lo[k] != p.hid_n[k]) return false;
  }
  return tc2_layout(tc2_b_stage(p), p.F).total <= 227 * 1024;
}

template <int D>
static int tc2_launch(const CinTcParams& p_in, cudaStream_t st) {
  CinTcParams p = p_in;
  p.b_stage_bytes = tc2_b_stage(p);
  const T2Smem lay = tc2_layout(p.b_stage_bytes, p.F);
  auto kern = cin_tc2_fwd_kernel<D>;
  DTB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.total));
  const int R = 128 / D;
  const int n_super = (p.B + 2 * R - 1) / (2 * R);
  int grid = sm_count();
  if (grid > n_super) grid = n_super;
  kern<<<grid, kT2Threads, lay.total, st>>>(p);
  DTB_LAUNCH_OK();
  return DTB_OK;
}

int cin_tc2_launch_fwd(const CinTcParams& p, int D, cudaStream_t st) {
  if (D == 16) return tc2_launch<16>(p, st);
  if (D == 32) return tc2_launch<32>(p, st);
  set_error("cin_tc2: embedding dim %d unsupported", D);
  return DTB_ERR_UNSUPPORTED;
}

}  // namespace dtb
